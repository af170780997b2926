#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan_sort.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ss.json 2> gpurun_out/bench_ss.err; tail -3 gpurun_out/bench_ss.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ss.json')); x=d.pop('extra',{})
for r in x.get("scan_sort",[]): print(r)
PY
