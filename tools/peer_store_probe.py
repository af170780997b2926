"""Single-GPU diagnostic: cost of the fused all-gather's extra epilogue stores on the SM side (no NVLink): the per-rank GEMM of
BASELINE configs[4] (4096 x 32768 x 32768, bf16) with the epilogue storing each output block 1, 2, 4 and 8 times locally, in both
epilogue forms (WGB_TC_EPI: 0 = per-lane global stores, 1 = shared-memory staged TMA bulk stores).
    python tools/peer_store_probe.py [N = K] [M]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from wgmath_b200._lib import check, lib  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
L = lib()
N = K = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
a = w.TensorBuilder.matrix(M, K, ST).build(dev, "bf16")
b = w.TensorBuilder.matrix(K, N, ST).build(dev, "bf16")
c = w.TensorBuilder.matrix(M, N, ST).build(dev, "bf16")
enc = dev.create_command_encoder()
with enc.compute_pass("init", None) as p:
    w.fill_uniform(dev, p, a, 1)
    w.fill_uniform(dev, p, b, 2)
gemm = w.Gemm.from_device(dev)
# long warm-up first and x1 repeated at the end: a cold box runs the first ~100 ms in the burst power regime
for fake, epi in ((1, 0), (1, 0), (1, 1), (8, 0), (8, 1), (1, 0), (1, 1), (8, 1), (8, 0), (2, 1), (4, 1), (1, 0)):
    os.environ["WGB_TC_DEBUG_FAKE_PEERS"] = str(fake)
    os.environ["WGB_TC_EPI"] = str(epi)
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    steps = 30 if N >= 32768 else 300
    with enc.compute_pass("t", None) as p:
        for _ in range(steps // 2):
            gemm.dispatch(dev, shapes, p, c, a, b)
        check(L.wgb_event_record(e0, p._h))
        for _ in range(steps):
            gemm.dispatch(dev, shapes, p, c, a, b)
        check(L.wgb_event_record(e1, p._h))
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    t = ms.value / steps
    print(f"PEERSTORE {M}x{N}x{K} bf16, {'TMA bulk' if epi else 'per-lane'} epilogue stores x{fake}: {t:.3f} ms  {2.0 * M * N * K / t / 1e9:.1f} TFLOP/s", flush=True)
