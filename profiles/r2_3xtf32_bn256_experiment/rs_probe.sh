#!/bin/bash
# 1 GPU: the 3xTF32 BLOCK_N 256 (RS) kernels: forced-variant parity, accuracy at long K, timing sweep
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run rs_variants 600 python -m pytest tests/test_gpu_tc_variants.py -q -x -k "tf32 and 256 and 3"
WGB_3X_BN256=1 WGB_TF32_FUSED_SPLIT=0 run rs_acc 300 python tools/tc_probe.py acc 2
run rs_f32probe 600 python tools/f32_probe.py 1024 2048 4096 8192
grep -v "^\.\|^$" $O/rs_variants.log | tail -12 | cut -c1-250
grep -h "^ACC" $O/rs_acc.log | grep 3xtf32
grep -h "^F32PROBE" $O/rs_f32probe.log | grep -v "auto\|tf32 single"
