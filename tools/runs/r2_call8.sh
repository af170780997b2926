#!/bin/bash
# sharded probe at N ranks (argument), one process per GPU
mkdir -p gpurun_out
N=${1:-2}
export PROBE_STEPS=${2:-20} PROBE_REPS=${3:-2}
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/sharded_probe.py ) > gpurun_out/c8_probe_n$N.log 2>&1
echo "rc=$?"; grep -h "SHARDPROBE\|Error\|error" gpurun_out/c8_probe_n$N.log | cut -c1-250; tail -4 gpurun_out/c8_probe_n$N.log | cut -c1-300
