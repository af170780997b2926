#!/bin/bash
# What the driver runs at round end on one GPU: pytest -m gpu, smoke(), both bench arms.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run final_pytest_gpu 1800 python -m pytest tests -x -q -m gpu
run final_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run final_bench_ref 600 python bench.py --impl reference --steps 5 --warmup 3
run final_bench 900 python bench.py
python tools/show_bench.py $O/final_bench_ref.log | cut -c1-900
python tools/show_bench.py $O/final_bench.log | grep -v "geometry\|scan_sort\|level12\|gemm_f32" | cut -c1-3000
