#!/bin/bash
# round-2 call 12: kernel breakdown of the nonlocal apply on the INT8 tensor-core path
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "i8_emulated" 2>&1 | tail -5
ONLY_NONLOCAL=1 BACKENDS=4 REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 60 --csv --log-file gpurun_out/i8_nonlocal_launches_r2.csv python scripts/i8_perf_probe.py > gpurun_out/i8_ncu2.log 2>&1
tail -4 gpurun_out/i8_ncu2.log
