#!/bin/bash
# round-2 call 28 (N GPUs): SCF iterations of the C3 cell with all ranks on its one k-point (comm_slab)
set -x
N=${N:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-library --no-small --no-e2e --no-sharded --scf-steps 0 --slab-scf-steps 3 > gpurun_out/bench_slabscf_n${N}.json 2> gpurun_out/bench_slabscf_n${N}.err
tail -c 1500 gpurun_out/bench_slabscf_n${N}.err
python - <<P
import json
d = json.load(open("gpurun_out/bench_slabscf_n${N}.json"))
print(json.dumps(d.get("single_k_slab"), indent=1))
P
