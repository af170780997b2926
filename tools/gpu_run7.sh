#!/bin/bash
# GPU call 7: PDL + fused epilogue validation, load-skip diagnostics, compute-sanitizer memcheck, bench.
mkdir -p gpurun_out
echo "=== quick probe cg=2 (PDL on) ==="
timeout 200 python tools/tc_probe.py quick 2 > gpurun_out/probe7_quick2.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe7_quick2.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe7_quick2.log | head -20
echo "=== perf PDL on ==="
timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe7_perf_pdl.log 2>&1; cat gpurun_out/probe7_perf_pdl.log
echo "=== perf PDL off ==="
WGB_TC_PDL=0 timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe7_perf_nopdl.log 2>&1; cat gpurun_out/probe7_perf_nopdl.log
for sk in 1 2 3; do
echo "=== diagnostic: WGB_TC_DEBUG_SKIP=$sk (1 = no A loads, 2 = no B loads, 3 = no loads; results invalid, timing only) ==="
WGB_TC_DEBUG_SKIP=$sk timeout 200 python tools/tc_probe.py perf 2 2>&1 | grep -E "bf16 tr=0 n=(4096|8192)|tf32" | tee gpurun_out/probe7_skip$sk.log
done
echo "=== pytest gpu ==="
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== compute-sanitizer memcheck (subset) ==="
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "reference_replay or simt_any_shape or views_offsets or gemv_multi_column or op_assign_lengths or reduce_all_ops or reduce_columns or golden or bf16_tcgen05 or f32_3xtf32 or fused_op" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -3 gpurun_out/bench5.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5.json')); x=d.pop('extra',{})
print(json.dumps(d)[:1800])
PY
