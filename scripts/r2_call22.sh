#!/bin/bash
# round-2 call 22 (1 GPU): kernel-time distribution of the batched small-k-block path (C4 and C5), launch list under ncu
set -x
mkdir -p gpurun_out
for c in C4 C5; do
  CONFIG=$c timeout 300 python scripts/small_profile.py > gpurun_out/small_${c}.log 2>&1
  tail -3 gpurun_out/small_${c}.log
done
CONFIG=C4 MAXITER=2 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_c4_r2.csv python scripts/small_profile.py > gpurun_out/small_c4_ncu.log 2>&1
tail -3 gpurun_out/small_c4_ncu.log
ls -la gpurun_out/launches_c4_r2.csv
