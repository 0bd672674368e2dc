#!/bin/bash
# round-2 call 17: full GPU suite and the N=1 bench with the default options (INT8 tensor-core path on)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_r2.json 2> gpurun_out/bench_n1_r2.err
tail -c 1500 gpurun_out/bench_n1_r2.err
head -c 2500 gpurun_out/bench_n1_r2.json
