#!/bin/bash
# Checkpoint after the lean issue loops + enqueued host GEMM: tests, smoke, both bench arms, ncu launch list + full GEMM capture.
mkdir -p gpurun_out
echo "=== pytest gpu ==="
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench reference arm ==="
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-600 gpurun_out/bench_ref.json
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); x=d.pop('extra',{})
print(json.dumps(d)[:3000])
PY
echo "=== ncu launch list (bench --no-extras) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
echo "=== ncu full: gemm_tc (bf16 4096^3) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -o gpurun_out/prof_gemm_tc python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log | cut -c1-200
