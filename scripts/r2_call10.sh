#!/bin/bash
# round-2 call 10: INT8 tensor-core path inside LOBPCG / SCF (parity), LOBPCG at C3 backend 0 vs 4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or i8_emulated" 2>&1 | tail -8
BACKENDS=0,4 MAXITER=6 timeout 900 python scripts/lobpcg_probe.py 2>&1 | grep -v "^  \|profile" | tail -8
