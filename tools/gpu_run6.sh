#!/bin/bash
# GPU call 6: tail N-split validation + perf, graph API test, full tests, bench, profiles.
mkdir -p gpurun_out
for cg in 2 1; do
echo "=== quick probe cg=$cg ==="
timeout 200 python tools/tc_probe.py quick $cg > gpurun_out/probe6_quick$cg.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe6_quick$cg.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe6_quick$cg.log | head -20
done
echo "=== perf cg=2 nsplit on ==="
timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe6_perf_nsplit.log 2>&1; cat gpurun_out/probe6_perf_nsplit.log
echo "=== perf cg=2 nsplit off ==="
WGB_TC_NSPLIT=0 timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe6_perf_nonsplit.log 2>&1; cat gpurun_out/probe6_perf_nonsplit.log
echo "=== pytest gpu ==="
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench4.json')); x=d.pop('extra',{})
print(json.dumps(d))
for k,v in x.items():
    for r in v: print(r)
PY
echo "=== ncu launch list (bench --no-extras) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
echo "=== ncu full: gemm_tc (bf16 4096^3) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -o gpurun_out/prof_gemm_tc python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log | cut -c1-200
