#!/bin/bash
# GPU call 8 (4 GPUs): N = 4 row-sharded parity + bench (16384^3).
mkdir -p gpurun_out
echo "=== N=4 sharded check ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tools/sharded_check.py > gpurun_out/sharded_check4.log 2>&1; echo "rc=$?"; grep -E "rank 0|SHARDED|rror" gpurun_out/sharded_check4.log | tail -8
echo "=== bench N=4 fused (16384^3) ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4_fused.json 2> gpurun_out/bench_n4_fused.err; echo "rc=$?"; tail -c 1800 gpurun_out/bench_n4_fused.json; tail -3 gpurun_out/bench_n4_fused.err
echo "=== bench N=4 reference arm ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 4 --steps 3 --warmup 3 2>/dev/null | cut -c1-400
