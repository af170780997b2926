#!/bin/bash
# Re-entry validation: GPU tests on the restored tree, GEMM timeline trace, host-link probe.
mkdir -p gpurun_out
echo "=== pytest gpu ==="
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== tc_trace ==="
timeout 200 python tools/tc_trace.py 4096 bf16 2>&1 | tee gpurun_out/tc_trace_4096.txt | tail -30
echo "=== pcie probe ==="
timeout 300 python tools/pcie_probe.py 2>&1 | tee gpurun_out/pcie_probe.txt | grep -E "PCIE|HOSTGEMM|CUBLAS|Error|error"
