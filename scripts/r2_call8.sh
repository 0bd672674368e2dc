#!/bin/bash
# round-2 call 8: TMA-fed INT8 tensor-core kernel (gemm_backend 4): parity test, then timing against backends 0 / 1 / 3
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "i8_emulated and -4]" 2>&1 | tail -15
ONLY_NONLOCAL=1 BACKENDS=0,3,4 timeout 300 python scripts/i8_perf_probe.py 2>&1 | tail -6
BACKENDS=0,1,4 timeout 300 python scripts/i8_perf_probe.py 2>&1 | tail -8
