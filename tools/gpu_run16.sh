#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu geometry ==="
timeout 900 python -m pytest tests/test_gpu_geometry.py -m gpu -q -x 2>&1 | tail -15
echo "=== memcheck ==="
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_geometry.py -m gpu -q -x -k "ragged or sub_range or special or in_place" 2>&1 | tail -4
echo "=== geom probe ==="
timeout 300 python tools/geom_probe.py 22 2>&1 | tail -20
echo "=== ncu geometry ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"geom_batch_kernel" -c 17 -o gpurun_out/prof_geometry env GEOM_NCU=1 python tools/geom_probe.py 22 > gpurun_out/ncu_geom.log 2>&1; tail -2 gpurun_out/ncu_geom.log
