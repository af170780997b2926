#!/bin/bash
# N = 2, experimental sliced B upload (WGB_SHARD_B_UPLOAD=1): bench e2e only.
mkdir -p gpurun_out
WGB_SHARD_B_UPLOAD=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_splitb.json 2> gpurun_out/bench_n2_splitb.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2_splitb.json').read().strip().splitlines()[-1])
    print(d["value"]); print(json.dumps(d["e2e"])[:900])
except Exception as e:
    print("no line", e)
PY
tail -5 gpurun_out/bench_n2_splitb.err
