#!/bin/bash
# Round 2, GPU call 5 (1 GPU): pair-loop epilogue: correctness (variants, loopback, parity suite) and A/B against the round-1 library.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c5_variants 900 python -m pytest tests/test_gpu_tc_variants.py -q -x
run c5_loopback 600 python tools/loopback_check.py
for i in 1 2 3; do
  run c5_ab_r1_$i 120 python tools/ab_probe.py wgmath_b200/libwgebra_b200_r1.so 4096 8192
  run c5_ab_r2_$i 120 python tools/ab_probe.py wgmath_b200/libwgebra_b200.so 4096 8192
done
run c5_pytest_gpu 1500 python -m pytest tests -m gpu -q
run c5_peerstore8k 200 python tools/peer_store_probe.py 8192
grep -h "FAIL\|LOOPBACK" $O/c5_loopback.log
grep -h "^AB" $O/c5_ab_*.log | awk '{print $2, $4, $5}' | sort | awk '{k=$1" "$2; a[k]=a[k]" "$3} END {for (k in a) print k, a[k]}' | sort
grep -h PEERSTORE $O/c5_peerstore8k.log
