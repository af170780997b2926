#!/bin/bash
# compute-sanitizer racecheck + synccheck over the shared-memory kernels added since the memcheck run (geometry, scan / sort),
# plus the level-1 / gemv families; small problem sizes only (racecheck slows kernels by one to two orders of magnitude).
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  echo "=== $tool: geometry ==="
  timeout 500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_geometry.py -m gpu -q -x -k "ragged or sub_range or special or in_place" 2>&1 | tail -3
  echo "=== $tool: scan / sort ==="
  timeout 500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_scan_sort.py -m gpu -q -x -k "reference_replay or lengths or sub_view or few_distinct or (matches_the_oracle and not 300_000 and not 70001)" 2>&1 | tail -3
done
echo "=== nvtx smoke ==="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
