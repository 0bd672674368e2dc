#!/bin/bash
# round-2 call 3: batched LOBPCG bring-up (GPU suite), C2 SCF timing, launch list of the INT8 GEMM (backend 3)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "lobpcg" 2>&1 | tail -30
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
ORACLE=0 timeout 300 python scripts/small_scf_probe.py 2>&1 | tail -20
ONLY_NONLOCAL=1 BACKENDS=3 REPS=1 K=264859 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/i8_launches_r2.csv python scripts/i8_perf_probe.py > gpurun_out/i8_ncu.log 2>&1
tail -5 gpurun_out/i8_ncu.log
