"""GPU test (-m gpu) of the N > 1 path on real devices: skipped unless the box exposes >= 2 GPUs.
Launches tools/sharded_check.py with one rank per GPU (2 ranks) over 127.0.0.1."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.gpu
def test_row_sharded_gemm_two_ranks():
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "SHARDED CHECK ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
