#!/bin/bash
# Enqueued host GEMMs (two device slots, three streams): parity + bench line (no extras).
mkdir -p gpurun_out
echo "=== pytest gpu (host gemm) ==="
timeout 900 python -m pytest tests -m gpu -q -x -k "host" 2>&1 | tail -3
echo "=== bench N=1 (no extras) ==="
timeout 900 python bench.py --steps 50 --warmup 5 --no-extras > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -3 gpurun_out/bench_r1c.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1c.json')); x=d.pop('extra',{})
print(json.dumps(d["e2e"], indent=1)); print(d["value"], d["roofline"]["frac"])
PY
