#!/bin/bash
# Round 2, GPU call 1 (1 GPU): every new code path in its own process, so one sticky CUDA error cannot hide the rest.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(tail -n 4 $O/$name.log | cut -c1-300 | tr '\n' '|')"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
run c1_tc_quick 300 python tools/tc_probe.py quick
run c1_tf32_input 120 python tools/tf32_input_probe.py
run c1_variants 900 python -m pytest tests/test_gpu_tc_variants.py -x -q
run c1_loopback 600 python tools/loopback_check.py
run c1_lookback 300 python -m pytest tests/test_gpu_scan_sort.py -x -q -k lookback
run c1_advice 300 python -m pytest tests/test_gpu_parity.py -x -q -k "sizes_change or reallocation or cfg3"
run c1_pytest_gpu 1500 python -m pytest tests -m gpu -q
run c1_peerstore32k 300 python tools/peer_store_probe.py 32768
run c1_peerstore8k 200 python tools/peer_store_probe.py 8192
run c1_peerstore4k 200 python tools/peer_store_probe.py 4096
run c1_scan_default 120 python tools/ss_probe.py 26
WGB_SCAN_LOOKBACK=1 run c1_scan_lookback 120 python tools/ss_probe.py 26
run c1_bench 900 python bench.py --steps 20 --warmup 5
python tools/show_bench.py $O/c1_bench.log
grep -h "PEERSTORE\|TF32INPUT\|SCAN\|SORT" $O/c1_*.log
