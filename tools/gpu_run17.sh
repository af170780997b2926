#!/bin/bash
# Round-1 validation after the geometry widening: what the driver runs at round end + captures of the newer kernels.
mkdir -p gpurun_out
echo "=== pytest gpu (all) ==="
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench reference arm ==="
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-600 gpurun_out/bench_ref.json
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json')); x=d.pop('extra',{})
print(json.dumps(d)[:2500])
for k,v in x.items():
    for r in v: print(r)
PY
echo "=== geom probe large ==="
timeout 300 python tools/geom_probe.py 24 2>&1 | tail -20
echo "=== ncu scan / sort ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_reduce_kernel|scan_apply_kernel|radix_count_kernel|radix_scatter_kernel" -c 12 -o gpurun_out/prof_scan_sort2 env SS_NCU=1 python tools/ss_probe.py 26 > gpurun_out/ncu_ss2.log 2>&1; tail -2 gpurun_out/ncu_ss2.log
echo "=== launch list of bench ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu2.log 2>&1; tail -1 gpurun_out/bench_under_ncu2.log | cut -c1-200
