#!/bin/bash
# 1 GPU: gemv_tr after the wave-aware split, graph test, then the ncu evidence of the round (launch list + full captures)
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
run c19_tests 600 python -m pytest tests/test_gpu_parity.py -q -x -k "gemv or graph or reallocation"
run c19_bench 900 python bench.py --steps 20 --warmup 5
python tools/show_bench.py $O/c19_bench.log | grep "level12\|^{" | cut -c1-400
# launch list of the bench command (shares only: cold caches, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/c19_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
# full captures: the headline kernel, the level-2 kernels, the 3xTF32 kernels (in-kernel split at 1024, split kernels + GEMM at 4096)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -o $O/r2_prof_gemm_tc python bench.py --steps 2 --warmup 3 --no-extras > $O/c19_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_n_kernel|gemv_t_kernel|reduce_kernel|op_assign_kernel|reduce_columns" -c 8 -o $O/r2_prof_level12 python tools/l2_probe.py > $O/c19_ncu_l12.log 2>&1; echo "ncu l12 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|split_tf32" -s 3 -c 6 -o $O/r2_prof_3xtf32 python tools/f32_probe.py 1024 4096 > $O/c19_ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
ls -la $O/*.ncu-rep
