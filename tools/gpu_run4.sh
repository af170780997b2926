#!/bin/bash
# GPU call 4 (2 GPUs): re-validate GEMM after the split-K fixup + chunked 3xTF32 changes, full tests, bench, profiles,
# then the N = 2 row-sharded paths (NCCL and fused peer stores).
mkdir -p gpurun_out
nvidia-smi -L
echo "=== quick probe cg=2 ==="
timeout 200 python tools/tc_probe.py quick 2 > gpurun_out/probe4_quick2.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe4_quick2.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe4_quick2.log | head -20
echo "=== quick probe cg=1 ==="
timeout 200 python tools/tc_probe.py quick 1 > gpurun_out/probe4_quick1.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/probe4_quick1.log; grep -E "^FAIL|^EXC|QUICK" gpurun_out/probe4_quick1.log | head -10
echo "=== perf cg=2 split-K on ==="
timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe4_perf_split.log 2>&1; cat gpurun_out/probe4_perf_split.log
echo "=== perf cg=2 split-K off ==="
WGB_TC_SPLITK=0 timeout 200 python tools/tc_probe.py perf 2 > gpurun_out/probe4_perf_nosplit.log 2>&1; cat gpurun_out/probe4_perf_nosplit.log
echo "=== acc ==="
timeout 200 python tools/tc_probe.py acc 2 > gpurun_out/probe4_acc.log 2>&1; cat gpurun_out/probe4_acc.log
echo "=== pytest gpu ==="
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench3.json')); x=d.pop('extra',{})
print(json.dumps(d))
for k,v in x.items():
    for r in v: print(r)
PY
echo "=== bench reference arm ==="
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 | cut -c1-600
echo "=== ncu launch list ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
echo "=== ncu full: gemm_tc (bf16 4096^3) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -o gpurun_out/prof_gemm_tc python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log | cut -c1-200
echo "=== ncu full: gemv + level1 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_n_kernel|gemv_t_kernel|reduce_kernel|op_assign_kernel|reduce_columns_kernel" -c 12 -o gpurun_out/prof_level12 python tools/l2_probe.py > gpurun_out/ncu_l12.log 2>&1; tail -1 gpurun_out/ncu_l12.log
echo "=== N=2 sharded check ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/sharded_check.log 2>&1; echo "rc=$?"; grep -E "rank|SHARDED|Error|error" gpurun_out/sharded_check.log | tail -20
echo "=== bench N=2 fused ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_n2_fused.json; tail -3 gpurun_out/bench_n2_fused.err
echo "=== bench N=2 nccl ==="
WGB_SHARD_MODE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_n2_nccl.json; tail -3 gpurun_out/bench_n2_nccl.err
