#!/bin/bash
# round-2 call 5: latency diagnosis, C2 probe, z-pipeline A/B at C3
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,utilization.gpu,memory.used --format=csv
timeout 300 python scripts/sync_probe.py 2>&1 | tail -20
ORACLE=0 timeout 300 python scripts/small_scf_probe.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "apply_h or full_size or engines or sphere" 2>&1 | tail -5
timeout 600 python scripts/zpipe_probe.py 2>&1 | tail -12
