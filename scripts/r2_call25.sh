#!/bin/bash
# round-2 call 25 (1 GPU): full GPU suite, then the default bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1
tail -6 gpurun_out/gpu_tests_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_r2.json 2> gpurun_out/bench_n1_r2.err
tail -c 600 gpurun_out/bench_n1_r2.err
head -c 600 gpurun_out/bench_n1_r2.json
