#!/bin/bash
# round-2 call 19 (2 GPUs): section profile of the slab LOBPCG vs one GPU at C3
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 scripts/slab_probe.py > gpurun_out/slab_probe.out 2>&1
tail -5 gpurun_out/slab_probe.out
cat gpurun_out/slab_probe_rank0.log
grep -A3 "slab rank" gpurun_out/slab_probe_rank1.log
