"""GPU test (-m gpu) of the fused GEMM + all-gather protocol on a single-GPU box: tools/loopback_check.py runs P ranks (one
context each) on device 0 with their gathered buffers connected in-process, so the peer stores (per-lane and TMA bulk), the
ready / done flags, the epochs, the buffer rotation (depth 1 / 2 / 3) and the deferred wait are all exercised where
tests/test_gpu_multi.py has to be skipped.  Separate process: CUDA_DEVICE_MAX_CONNECTIONS must be set before CUDA starts."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fused_gather_protocol_on_one_gpu():
    # (EAGER: no kernel is loaded for the first time while another rank's GEMM spins on a flag, see comm.cu)
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", WGB_PEER_TIMEOUT_MS="20000", CUDA_MODULE_LOADING="EAGER")
    for k in ("WGB_TC_EPI", "WGB_TC_BN", "WGB_TC_CG", "WGB_TC_NSPLIT", "WGB_TC_SPLITK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "loopback_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0 and "LOOPBACK CHECK ALL OK" in r.stdout, r.stdout[-4000:] + r.stderr[-3000:]
