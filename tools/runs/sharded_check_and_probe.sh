#!/bin/bash
# N GPUs: sharded check (incl. multicast forms) + sharded probe
mkdir -p gpurun_out
N=${1:-2}
export PROBE_STEPS=${2:-30} PROBE_REPS=${3:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29511 tools/sharded_check.py $4 ) > gpurun_out/c13_check_n$N.log 2>&1
echo "check rc=$?"; grep -h "rank 0\]\|SHARDED\|multicast" gpurun_out/c13_check_n$N.log | cut -c1-420; grep -h "Error\|error" gpurun_out/c13_check_n$N.log | head -5 | cut -c1-300
( time timeout 400 $TR --master-port 29513 tools/sharded_probe.py ) > gpurun_out/c13_probe_n$N.log 2>&1
echo "probe rc=$?"; grep -h "SHARDPROBE" gpurun_out/c13_probe_n$N.log | cut -c1-250; grep -h "Error\|error" gpurun_out/c13_probe_n$N.log | head -5 | cut -c1-300
