#!/bin/bash
# round-2 call 7 (2 GPUs): ncu launch lists + full captures (GPU 0 only; reports converted to CSV on the box), then the 2-rank bench
set -x
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
M=102 REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|kr_|kb_' -c 400 --csv --log-file gpurun_out/launches_r2.csv python scripts/profile_probe.py > gpurun_out/ncu_a.log 2>&1
M=51 REPS=1 timeout 900 ncu --set full --clock-control none -k regex:'kr_|k_zgemm|k_i8_gemm|k_i8_crt_nn|k_i8_residues_ld4' -c 16 -f -o /tmp/prof_r2 python scripts/profile_probe.py > gpurun_out/ncu_b.log 2>&1
ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv 2>> gpurun_out/ncu_b.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|kr_|kb_' -c 3000 --csv --log-file gpurun_out/launches_small_r2.csv python scripts/sync_probe.py > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log
ls -la gpurun_out
unset CUDA_VISIBLE_DEVICES
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err
tail -c 2500 gpurun_out/bench_n2_r2.err
head -c 3000 gpurun_out/bench_n2_r2.json
