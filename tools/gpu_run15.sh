#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ss_probe.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"prefix_sum_kernel|radix_onesweep_kernel|radix_hist_kernel" -s 4 -c 4 -o gpurun_out/prof_scan_sort python tools/ss_probe.py > gpurun_out/ncu_ss.log 2>&1; tail -2 gpurun_out/ncu_ss.log
