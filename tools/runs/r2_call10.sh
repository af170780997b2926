#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/link_probe.py ) > gpurun_out/c10_link_n$N.log 2>&1
echo "rc=$?"; grep -h "^LINK" gpurun_out/c10_link_n$N.log | cut -c1-250; grep -v "^LINK" gpurun_out/c10_link_n$N.log | tail -12 | cut -c1-300
