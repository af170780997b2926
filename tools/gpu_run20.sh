#!/bin/bash
for m in all; do python -X faulthandler tools/exit_probe.py $m 2>&1 | tail -12; echo "rc=${PIPESTATUS[0]}"; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py tests/test_gpu_scan_sort.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python tools/geom_probe.py 20 | tail -2; echo "probe rc=${PIPESTATUS[0]}"
