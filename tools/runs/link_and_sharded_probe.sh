#!/bin/bash
# 8 GPUs: link micro-benchmark, then the sharded probe (all fused forms on the same box)
mkdir -p gpurun_out
N=${1:-8}
export PROBE_STEPS=${2:-30} PROBE_REPS=${3:-2}
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/link_probe.py ) > gpurun_out/c12_link_n$N.log 2>&1
echo "link rc=$?"; grep -h "^LINK" gpurun_out/c12_link_n$N.log | cut -c1-250
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/sharded_probe.py ) > gpurun_out/c12_probe_n$N.log 2>&1
echo "probe rc=$?"; grep -h "SHARDPROBE" gpurun_out/c12_probe_n$N.log | cut -c1-250; grep -v "SHARDPROBE" gpurun_out/c12_probe_n$N.log | tail -5 | cut -c1-300
