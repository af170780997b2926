#!/bin/bash
# 1 GPU: the streaming prefix sum — forced-path parity tests, then the tuning sweep
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run scan_tests 600 python -m pytest tests/test_gpu_scan_sort.py tests/test_gpu_reference_vectors.py -q -x -k "prefix_sum"
run scan_probe 300 python tools/scan_probe.py 26 24 22
grep -h "^SCANPROBE" $O/scan_probe.log
grep -v "^\.\|^$" $O/scan_tests.log | tail -15 | cut -c1-250
