#!/bin/bash
# round-2 call 4: new tests (Ewald kernels, LdosMixing, mixing helpers, random orbitals, multi-k FFT), full suite, small-SCF timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ewald or ldos or mixing_helpers or random_orbitals or lobpcg or c_program" 2>&1 | tail -30
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
ORACLE=0 timeout 300 python scripts/small_scf_probe.py 2>&1 | tail -22
ORACLE=0 timeout 300 python -m cProfile -s tottime scripts/small_scf_probe.py 2>&1 | grep -A45 "Ordered by" | head -60
