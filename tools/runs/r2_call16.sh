#!/bin/bash
# 1 GPU: fused-split 3xTF32: variants test, 3xTF32 parity tests, accuracy at long K, timing sweep
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c16_variants 900 python -m pytest tests/test_gpu_tc_variants.py -q -x -k "passes=3"
run c16_parity_f32 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_orderings.py -q -x -k "3xtf32 or f32 or tf32 or cfg2 or fused_op or golden or reference_replay"
run c16_acc 300 python tools/tc_probe.py acc
run c16_f32probe 600 python tools/f32_probe.py
grep -h "^ACC" $O/c16_acc.log
grep -h "^F32PROBE" $O/c16_f32probe.log
