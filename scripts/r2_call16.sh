#!/bin/bash
# round-2 call 16: INT8 tensor-core path (both product types): parity tests, clean timings of nonlocal apply and LOBPCG at C3
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or i8_emulated or lobpcg" 2>&1 | tail -6
ONLY_NONLOCAL=1 BACKENDS=0,4 timeout 300 python scripts/i8_perf_probe.py 2>&1 | tail -5
BACKENDS=0,4 MAXITER=6 timeout 900 python scripts/lobpcg_probe.py 2>&1 | grep -v "^  \|profile" | tail -6
