#!/bin/bash
# Round 2, GPU call 2 (1 GPU): loopback diagnosis, variants re-run, A/B of the round-1 and round-2 libraries.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
export WGB_PEER_TIMEOUT_MS=12000
LOOPBACK_DEBUG=5 LOOPBACK_CASES=6 run c2_lb_d3_deferred_tma 120 python tools/loopback_check.py
LOOPBACK_DEBUG=5 LOOPBACK_CASES=2 run c2_lb_d3_deferred_stg 120 python tools/loopback_check.py
LOOPBACK_DEBUG=5 LOOPBACK_CASES=4 run c2_lb_d1_lock_stg 120 python tools/loopback_check.py
LOOPBACK_DEBUG=5 LOOPBACK_CASES=0 run c2_lb_d1_lock_tma 120 python tools/loopback_check.py
CUDA_MODULE_LOADING=EAGER LOOPBACK_DEBUG=5 LOOPBACK_CASES=0 run c2_lb_d1_lock_tma_eager 120 python tools/loopback_check.py
unset WGB_PEER_TIMEOUT_MS
run c2_variants 900 python -m pytest tests/test_gpu_tc_variants.py -q
run c2_ab 300 python tools/ab_probe.py wgmath_b200/libwgebra_b200_r1.so wgmath_b200/libwgebra_b200.so 4096 8192
cat $O/c2_lb_*.log | grep -v "^$\|^real\|^user\|^sys"
tail -15 $O/c2_variants.log
grep -h "^AB" $O/c2_ab.log
