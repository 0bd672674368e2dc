#!/bin/bash
# round-2 call 14: ncu --set full on the update-type INT8 kernels
set -x
mkdir -p gpurun_out
NONLOCAL_APPLY=1 ONLY_NONLOCAL=1 BACKENDS=4 REPS=1 timeout 900 ncu --set full --clock-control none -k regex:'k_i8_gemm_tc2_nn|k_i8_crt_nn|k_i8_gemm_tc2$' -c 4 -f -o /tmp/prof_i8 python scripts/i8_perf_probe.py > gpurun_out/i8_ncu3.log 2>&1
ncu -i /tmp/prof_i8.ncu-rep --page raw --csv > gpurun_out/prof_i8_raw.csv 2>> gpurun_out/i8_ncu3.log
tail -3 gpurun_out/i8_ncu3.log; ls -la gpurun_out/prof_i8_raw.csv
