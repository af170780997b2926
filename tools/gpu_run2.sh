#!/bin/bash
# GPU call 2: bring-up of the tcgen05 GEMM + profiles of the level-1 kernels.
mkdir -p gpurun_out
echo "=== quick probe (all configs) ==="
timeout 300 python tools/tc_probe.py quick > gpurun_out/probe_quick.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe_quick.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe_quick.log | head -40
echo "=== acc ==="
timeout 200 python tools/tc_probe.py acc > gpurun_out/probe_acc.log 2>&1; cat gpurun_out/probe_acc.log | tail -20
echo "=== perf ==="
timeout 300 python tools/tc_probe.py perf > gpurun_out/probe_perf.log 2>&1; cat gpurun_out/probe_perf.log | tail -40
echo "=== pytest gpu (CG=1) ==="
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "=== ncu level-1 ==="
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"reduce_kernel|op_assign_kernel" -c 6 -o gpurun_out/prof_l1 python tools/l1_probe.py > gpurun_out/ncu_l1.log 2>&1; tail -3 gpurun_out/ncu_l1.log
