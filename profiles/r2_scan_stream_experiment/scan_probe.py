"""Exclusive prefix sum at n = 2^22 .. 2^26: the three-step form against the chunk-pipelined persistent kernel
(scan_stream_kernel) over its two tuning knobs, CTAs per SM and steps of slack between the two passes of a chunk.
    python tools/scan_probe.py [log2 n ...]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from wgmath_b200._lib import check, lib  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
L = lib()
ps = w.WgPrefixSum.from_device(dev)
ws = w.PrefixSumWorkspace.new()


def timed(fn, steps=20, warm=5):
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("t", None) as p:
        for _ in range(warm):
            fn(p)
        check(L.wgb_event_record(e0, p._h))
        for _ in range(steps):
            fn(p)
        check(L.wgb_event_record(e1, p._h))
    dev.poll_wait()
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    return ms.value / steps


for lg in [int(a) for a in sys.argv[1:]] or [26, 24, 22]:
    n = 1 << lg
    host = np.random.default_rng(lg).integers(0, 256, n, dtype=np.uint32)
    d = w.TensorBuilder.vector(n, ST).build_init(dev, host, "u32")
    forms = [("three-step", {"WGB_SCAN_STREAM": "0"})]
    for ctas in (2, 3, 4):
        for depth in (1, 2, 3, 4):
            forms.append((f"stream ctas/SM={ctas} depth={depth}", {"WGB_SCAN_STREAM": "1", "WGB_SCAN_STREAM_CTAS": str(ctas), "WGB_SCAN_STREAM_DEPTH": str(depth)}))
    forms.append(("default", {}))
    for name, env in forms:
        for k in ("WGB_SCAN_STREAM", "WGB_SCAN_STREAM_CTAS", "WGB_SCAN_STREAM_DEPTH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t = timed(lambda p: ps.dispatch(dev, p, ws, d))
        print(f"SCANPROBE n=2^{lg} {name:28s}: {t * 1e3:8.1f} us  {8 * n / t / 1e6:7.0f} GB/s of algorithmic bytes", flush=True)
