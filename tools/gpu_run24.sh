#!/bin/bash
mkdir -p gpurun_out
echo "=== N=2 sharded check (incl. host-enqueue path) ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/sharded_check.py > gpurun_out/sharded_check2.log 2>&1; echo "rc=$?"; grep -E "rank 0|SHARDED|rror|Traceback" gpurun_out/sharded_check2.log | tail -8
echo "=== bench N=2 fused (8192^3) ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_fused.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"]); print(json.dumps(d["e2e"])[:1500])
PY
tail -3 gpurun_out/bench_n2_fused.err
