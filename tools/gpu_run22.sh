#!/bin/bash
# N = 8: configs[4] (32768^3 bf16 row-sharded, fused all-gather) after the lean issue loops.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_fused.json 2> gpurun_out/bench_n8_fused.err; echo "rc=$?"; cut -c1-1200 gpurun_out/bench_n8_fused.json; tail -2 gpurun_out/bench_n8_fused.err
