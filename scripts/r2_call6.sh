#!/bin/bash
# round-2 call 6: full GPU suite with batched band energies / density / shared potentials / setup kernels; C2 + Al probes; bench N=1
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
ORACLE=0 timeout 300 python scripts/small_scf_probe.py 2>&1 | tail -4
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_r2.json 2> gpurun_out/bench_n1_r2.err
tail -c 3000 gpurun_out/bench_n1_r2.err
cat gpurun_out/bench_n1_r2.json | head -c 6000
