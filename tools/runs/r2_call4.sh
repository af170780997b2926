#!/bin/bash
# Round 2, GPU call 4 (1 GPU): where do the 0.6 us per launch go?  r1 / r2 / X (one dst map in the parameters) / Y (compile-time ring length, no TMA epilogue code)
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3; do
  for v in r1 expX expY; do
    timeout 120 python tools/ab_probe.py wgmath_b200/libwgebra_b200_$v.so 4096 > $O/c4_ab_${v}_$i.log 2>&1
  done
  timeout 120 python tools/ab_probe.py wgmath_b200/libwgebra_b200.so 4096 > $O/c4_ab_r2_$i.log 2>&1
done
grep -h "^AB" $O/c4_ab_*.log | sort | awk '{print $4, $5}' | sort | awk '{a[$1]=a[$1]" "$2} END {for (k in a) print k, a[k]}'
