#!/bin/bash
# Lean issue loops (converged warp + elect.sync): parity, timeline trace, GEMM sweep.
mkdir -p gpurun_out
echo "=== pytest gpu ==="
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== tc_trace ==="
timeout 200 python tools/tc_trace.py 4096 bf16 2>&1 | tee gpurun_out/tc_trace_4096_lean.txt | tail -24
echo "=== probe perf ==="
timeout 300 python tools/tc_probe.py perf 2>&1 | tee gpurun_out/tc_probe_perf_lean.txt | grep PERF
