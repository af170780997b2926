#!/bin/bash
# Final single-GPU validation: what the driver runs at round end (tests, smoke, both bench arms) + the HBM-kernel captures.
mkdir -p gpurun_out
echo "=== pytest gpu ==="
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench reference arm ==="
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-700 gpurun_out/bench_ref.json
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json')); x=d.pop('extra',{})
print(json.dumps(d)[:2500])
for k,v in x.items():
    for r in v: print(r)
PY
echo "=== ncu full: gemv + level1 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_n_kernel|gemv_t_kernel|reduce_kernel|op_assign_kernel|reduce_columns_kernel" -c 12 -o gpurun_out/prof_level12 python tools/l2_probe.py > gpurun_out/ncu_l12.log 2>&1; tail -1 gpurun_out/ncu_l12.log
