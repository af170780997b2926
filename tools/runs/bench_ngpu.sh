#!/bin/bash
# bench.py at N GPUs the way the driver launches it
mkdir -p gpurun_out
N=${1:-2}; STEPS=${2:-20}; WARM=${3:-5}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup $WARM ) > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?"; python tools/show_bench.py gpurun_out/bench_n$N.log | cut -c1-5000; grep -h "Error\|error\|PARITY" gpurun_out/bench_n$N.log | head -5 | cut -c1-300
