#!/bin/bash
# round-2 call 23 (1 GPU): pipelined two-group scheduler of the batched solves -- tests, then C2/C4/C5 SCF times with and without it
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_scf.py -x -q -k "lobpcg or baseline_config or random_orbitals or svd or silicon" > gpurun_out/pipe_tests.log 2>&1
tail -5 gpurun_out/pipe_tests.log
for c in C2 C4 C5; do
  for p in 1 0; do
    CONFIG=$c PIPE=$p timeout 300 python scripts/small_profile.py 2>&1 | grep -v Warn | tail -2
  done
done
