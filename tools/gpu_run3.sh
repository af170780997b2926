#!/bin/bash
# GPU call 3: validate tf32 MN-major (direct + transposed), tail split-K, then tests / bench / profiles.
mkdir -p gpurun_out
echo "=== quick probe cg=2 (transpose prep for f32 non-tr) ==="
timeout 200 python tools/tc_probe.py quick 2 > gpurun_out/probe3_quick2.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe3_quick2.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe3_quick2.log | head -20
echo "=== quick probe cg=2, WGB_TF32_MN_DIRECT=1 ==="
WGB_TF32_MN_DIRECT=1 timeout 200 python tools/tc_probe.py quick 2 > gpurun_out/probe3_quick2_direct.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe3_quick2_direct.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe3_quick2_direct.log | head -20
echo "=== quick probe cg=1 ==="
timeout 200 python tools/tc_probe.py quick 1 > gpurun_out/probe3_quick1.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe3_quick1.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe3_quick1.log | head -10
echo "=== perf cg=2 split-K on ==="
timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe3_perf_split.log 2>&1; cat gpurun_out/probe3_perf_split.log
echo "=== perf cg=2 split-K off ==="
WGB_TC_SPLITK=0 timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe3_perf_nosplit.log 2>&1; cat gpurun_out/probe3_perf_nosplit.log
echo "=== acc ==="
timeout 200 python tools/tc_probe.py acc 2 > gpurun_out/probe3_acc.log 2>&1; cat gpurun_out/probe3_acc.log
echo "=== pytest gpu ==="
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json')); x=d.pop('extra',{})
print(json.dumps(d))
for k,v in x.items():
    for r in v: print(r)
PY
echo "=== ncu launch list ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "=== ncu full: gemm_tc ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -o gpurun_out/prof_gemm_tc python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log | cut -c1-200
echo "=== ncu full: gemv + level1 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_n_kernel|gemv_t_kernel|reduce_kernel|op_assign_kernel|reduce_columns_kernel" -c 12 -o gpurun_out/prof_level12 python tools/l2_probe.py > gpurun_out/ncu_l12.log 2>&1; tail -2 gpurun_out/ncu_l12.log
