#!/bin/bash
# GPU call 2: bring-up of the tcgen05 GEMM + profiles of the level-1 kernels.
mkdir -p gpurun_out
for cg in 1 2; do
echo "=== quick probe cg=$cg ==="
timeout 200 python tools/tc_probe.py quick $cg > gpurun_out/probe_quick$cg.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe_quick$cg.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe_quick$cg.log | head -30
echo "=== perf cg=$cg ==="
timeout 200 python tools/tc_probe.py perf $cg > gpurun_out/probe_perf$cg.log 2>&1; tail -20 gpurun_out/probe_perf$cg.log
echo "=== acc cg=$cg ==="
timeout 200 python tools/tc_probe.py acc $cg > gpurun_out/probe_acc$cg.log 2>&1; tail -10 gpurun_out/probe_acc$cg.log
done
echo "=== pytest gpu (CG=1) ==="
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "=== ncu level-1 ==="
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"reduce_kernel|op_assign_kernel" -c 6 -o gpurun_out/prof_l1 python tools/l1_probe.py > gpurun_out/ncu_l1.log 2>&1; tail -3 gpurun_out/ncu_l1.log
