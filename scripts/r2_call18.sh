#!/bin/bash
# round-2 call 18 (2 GPUs): single-k multi-GPU (plane-wave slab LOBPCG) parity tests, then the slab section of the bench at C3
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -k slab > gpurun_out/slab_tests.log 2>&1
tail -30 gpurun_out/slab_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-library --no-small --no-sharded --no-e2e --scf-steps 0 > gpurun_out/bench_slab_n2.json 2> gpurun_out/bench_slab_n2.err
tail -c 2000 gpurun_out/bench_slab_n2.err
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/bench_slab_n2.json"))
    print(json.dumps(d.get("single_k_slab"), indent=1))
    print(json.dumps(d.get("lobpcg"), indent=1))
except Exception as e:
    print("no json", e)
P
