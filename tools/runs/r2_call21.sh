#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c21_new 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_reduce"
run c21_pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -v "^\.\|^$" $O/c21_new.log | tail -25 | cut -c1-300
grep -v "^\.\|^$" $O/c21_pytest_gpu.log | tail -12 | cut -c1-300
