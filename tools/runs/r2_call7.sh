#!/bin/bash
# Round 2, GPU call 7 (8 GPUs): bench N = 8 (lock step + pipelined + parity + e2e), sharded check.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run c7_bench_n8 600 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5
run c7_sharded_check 300 $TR --master-port 29511 tools/sharded_check.py quick
python tools/show_bench.py $O/c7_bench_n8.log
grep -h "rank 0\|SHARDED" $O/c7_sharded_check.log | cut -c1-400
