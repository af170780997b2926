#!/bin/bash
# Three-stream host GEMM + lean issue loops: parity, host-link probe, full bench line.
mkdir -p gpurun_out
echo "=== pytest gpu ==="
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== pcie probe ==="
timeout 300 python tools/pcie_probe.py 2>&1 | tee gpurun_out/pcie_probe2.txt | grep -E "PCIE|HOSTGEMM|CUBLAS|Error|error"
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -3 gpurun_out/bench_r1b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1b.json')); x=d.pop('extra',{})
print(json.dumps(d)[:2500])
for k,v in x.items():
    for r in v: print(r)
PY
