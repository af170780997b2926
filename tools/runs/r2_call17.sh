#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c17_variants 900 python -m pytest tests/test_gpu_tc_variants.py -q -x -k "passes=3"
run c17_acc 300 python tools/tc_probe.py acc 2
run c17_f32probe 600 python tools/f32_probe.py
grep -h "^ACC" $O/c17_acc.log | grep 3xtf32
grep -h "^F32PROBE" $O/c17_f32probe.log | grep -v "auto"
tail -30 $O/c17_variants.log | grep -v "^$" | tail -12
