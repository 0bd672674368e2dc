#!/bin/bash
# First GPU call of the next round: bring-up of the INT8 tensor-core path (see DESIGN.md "Groundwork").
#   gpurun --timeout 1200 -- 'bash scripts/next_round_first_call.sh > gpurun_out/i8_bringup.log 2>&1'
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/i8probe scripts/tcgen05_i8_probe.cu
for v in 0 1 2 3 mn0 mn1 rate; do timeout 60 /tmp/i8probe $v; done
export DFTK_B200_EXPERIMENTAL=1
# reference pipeline (backend 2: integer products on CUDA cores) against the FP64 DMMA GEMM
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "i8_emulated and -2]"
# tcgen05 kernel (backend 3): both descriptor conventions, pipelined and unpipelined; each in a fresh process because a
# faulting kernel poisons the CUDA context
for swap in 0 1; do for simple in 1 0; do
  echo "=== backend 3: DFTK_B200_I8TC_SWAP=$swap DFTK_B200_I8TC_SIMPLE=$simple"
  DFTK_B200_I8TC_SWAP=$swap DFTK_B200_I8TC_SIMPLE=$simple timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "i8_emulated and -3]" 2>&1 | tail -5
done; done
# product against ABINIT directly (PBE silicon, spin-polarised iron) and the forces tests
timeout 600 python -m pytest tests/test_gpu_scf.py -q -k "pbe_vs_abinit or collinear_vs_abinit"
# if backend 3 failed above: first-tile dump + diagnosis (add DFTK_B200_I8TC_SWAP / _SIMPLE as needed)
DFTK_B200_I8TC_DUMP=/tmp/i8tc DFTK_B200_I8TC_SIMPLE=1 timeout 120 python scripts/i8tc_debug.py 512 40 24
