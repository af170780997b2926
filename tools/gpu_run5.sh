#!/bin/bash
# GPU call 5 (8 GPUs): the N = 8 row-sharded path: parity check, then bench at BASELINE configs[4] (32768^3).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "=== N=8 sharded check ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/sharded_check.py > gpurun_out/sharded_check8.log 2>&1; echo "rc=$?"; grep -E "rank 0|SHARDED|Error|error|rror" gpurun_out/sharded_check8.log | tail -12
echo "=== bench N=8 fused (32768^3) ==="
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8_fused.json 2> gpurun_out/bench_n8_fused.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_n8_fused.json; tail -4 gpurun_out/bench_n8_fused.err
