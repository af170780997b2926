#!/bin/bash
# Round 2, GPU call 3 (1 GPU): full loopback check (also with lazy loading now that the flag kernels are preloaded), A/B timing.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c3_loopback 600 python tools/loopback_check.py
CUDA_MODULE_LOADING=LAZY WGB_PEER_TIMEOUT_MS=12000 LOOPBACK_CASES=0,4 run c3_loopback_lazy 120 python tools/loopback_check.py
for i in 1 2 3; do
  run c3_ab_r1_$i 120 python tools/ab_probe.py wgmath_b200/libwgebra_b200_r1.so 4096 8192
  run c3_ab_r2_$i 120 python tools/ab_probe.py wgmath_b200/libwgebra_b200.so 4096 8192
done
grep -h "OK \|FAIL\|LOOPBACK" $O/c3_loopback.log $O/c3_loopback_lazy.log
grep -h "^AB" $O/c3_ab_*.log | sort
