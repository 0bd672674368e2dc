#!/bin/bash
# round-2 call 26 (N GPUs): bench with the sharded SCFs (C5, C4) and the single-k slab LOBPCG section at N ranks
set -x
N=${N:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-library --no-small --no-e2e --scf-steps 0 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err
tail -c 1500 gpurun_out/bench_n${N}_r2.err
python - <<P
import json
d = json.load(open("gpurun_out/bench_n${N}_r2.json"))
print(d["value"], d["ms_per_step"])
print(json.dumps(d.get("single_k_slab"), indent=1))
for k, v in d.get("sharded_scf", {}).items():
    print(k, {q: v.get(q) for q in ("total_s", "n_iter", "s_per_iter", "blocks_this_rank", "collectives_per_step", "dE_per_atom_vs_oracle", "max_d_eigenvalue_vs_oracle", "parity_ok", "error")})
P
