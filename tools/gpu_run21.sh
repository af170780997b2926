#!/bin/bash
# N = 2 regression after the context-lifetime change: multi-GPU tests, sharded check, fused bench.
mkdir -p gpurun_out
echo "=== pytest multi ==="
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
echo "=== N=2 sharded check ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/sharded_check.py > gpurun_out/sharded_check2.log 2>&1; echo "rc=$?"; grep -E "rank 0|SHARDED|rror" gpurun_out/sharded_check2.log | tail -6
echo "=== bench N=2 fused (8192^3) ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_n2_fused.json; tail -3 gpurun_out/bench_n2_fused.err
