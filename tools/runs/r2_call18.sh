#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c18_new 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_reduce or gemv"
run c18_pytest_gpu 1500 python -m pytest tests -m gpu -q
run c18_bench 900 python bench.py --steps 20 --warmup 5
grep -v "^\.\|^$" $O/c18_pytest_gpu.log | tail -15 | cut -c1-300
python tools/show_bench.py $O/c18_bench.log | grep -v "geometry\|scan_sort" | cut -c1-2500
