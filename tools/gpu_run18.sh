#!/bin/bash
echo "=== ballots ==="
timeout 200 python tools/ss_probe.py 26 2>&1 | grep SORT
echo "=== match.any ==="
WGB_RS_MATCH=1 timeout 200 python tools/ss_probe.py 26 2>&1 | grep SORT
WGB_RS_MATCH=1 timeout 600 python -m pytest tests/test_gpu_scan_sort.py -m gpu -q 2>&1 | tail -2
