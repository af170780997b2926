#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): sharded check (all fused forms, host-enqueue with whole / sliced B), bench N = 2.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run c6_sharded_check 600 $TR --master-port 29511 tools/sharded_check.py
run c6_bench_n2 600 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5
grep -h "rank 0\|SHARDED" $O/c6_sharded_check.log | cut -c1-400
python tools/show_bench.py $O/c6_bench_n2.log
tail -5 $O/c6_bench_n2.log | cut -c1-600
