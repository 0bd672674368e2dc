#!/bin/bash
# round-2 call 20 (2 GPUs): un-profiled slab LOBPCG timing (repeated) vs the bench's slab section
set -x
mkdir -p gpurun_out
PROFILE=0 TAG=_noprof REPEATS=3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 scripts/slab_probe.py > gpurun_out/slab_probe2.out 2>&1
cat gpurun_out/slab_probe_rank0_noprof.log
