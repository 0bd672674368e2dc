#!/bin/bash
# First GPU call of the next round: bring-up of the INT8 tensor-core path (see DESIGN.md "Groundwork").
#   gpurun --timeout 900 -- 'bash scripts/next_round_first_call.sh > gpurun_out/i8_bringup.log 2>&1'
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/i8probe scripts/tcgen05_i8_probe.cu && timeout 120 /tmp/i8probe
# reference pipeline (backend 2) first, then the tcgen05 kernel (backend 3), each against the FP64 DMMA GEMM
DFTK_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "i8_emulated and -2]" -x
DFTK_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "i8_emulated and -3]"
