#!/bin/bash
# round-2 call 2 (2 GPUs): full GPU suite incl. the sharded-SCF parity tests, INT8 perf probe
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/i8_perf_probe.py 2>&1 | tail -12
