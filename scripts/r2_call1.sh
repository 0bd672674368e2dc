#!/bin/bash
# round-2 first call: INT8 bring-up (one shot), opt-in ABINIT tests, host-side profile of the C2 SCF
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/i8probe scripts/tcgen05_i8_probe.cu
for v in 0 1 2 3 mn0 mn1 rate; do timeout 40 /tmp/i8probe $v; done
export DFTK_B200_EXPERIMENTAL=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "i8_emulated and -2]" 2>&1 | tail -15
for swap in 0 1; do for simple in 1 0; do
  echo "=== backend 3: DFTK_B200_I8TC_SWAP=$swap DFTK_B200_I8TC_SIMPLE=$simple"
  DFTK_B200_I8TC_SWAP=$swap DFTK_B200_I8TC_SIMPLE=$simple timeout 120 python -m pytest tests/test_gpu_kernels.py -q -k "i8_emulated and -3]" 2>&1 | tail -5
done; done
timeout 500 python -m pytest tests/test_gpu_scf.py -q -k "pbe_vs_abinit or collinear_vs_abinit" 2>&1 | tail -15
unset DFTK_B200_EXPERIMENTAL
ORACLE=0 timeout 300 python -m cProfile -s tottime scripts/small_scf_probe.py 2>&1 | head -70
