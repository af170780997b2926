"""Per-cluster timeline of one tcgen05 GEMM launch in steady state (wgb_debug_tc_trace): where the ~100 us of a bf16 4096^3
launch go - launch / PDL wait, first-operand latency, the MMA issue loop (SM cycles per k-block = tensor-pipe rate at the
actual clock), the epilogue tail.  Usage: python tools/tc_trace.py [N] [dtype]"""
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
shapes = w.ViewShapeBuffers.new()
L = lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
sets = 4
a = [w.TensorBuilder.matrix(N, N, ST).build(dev, dtype) for _ in range(sets)]
b = [w.TensorBuilder.matrix(N, N, ST).build(dev, dtype) for _ in range(sets)]
c = [w.TensorBuilder.matrix(N, N, ST).build(dev, dtype) for _ in range(sets)]
enc = dev.create_command_encoder()
with enc.compute_pass("init", None) as p:
    for i in range(sets):
        w.fill_uniform(dev, p, a[i], 1 + i)
        w.fill_uniform(dev, p, b[i], 11 + i)
gemm = w.Gemm.from_device(dev)
check(L.wgb_debug_tc_trace(dev._h, 1, None, 0, None))
for label, skip in (("normal", "0"), ("no operand loads", "3")):
    os.environ["WGB_TC_DEBUG_SKIP"] = skip
    steps = 40
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("t", None) as p:
        for it in range(steps + 10):
            if it == 10:
                check(L.wgb_event_record(e0, p._h))
            gemm.dispatch(dev, shapes, p, c[it % sets], a[it % sets], b[it % sets])
        check(L.wgb_event_record(e1, p._h))
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    buf = (ctypes.c_ulonglong * (256 * 8))()
    n = ctypes.c_size_t(0)
    check(L.wgb_debug_tc_trace(dev._h, 1, buf, 256, ctypes.byref(n)))
    t = np.frombuffer(buf, dtype=np.uint64).reshape(256, 8).astype(np.float64)
    t = t[t[:, 5] > 0]
    t0 = t[:, 0].min()
    q = lambda x: f"min {np.min(x):9.2f}  med {np.median(x):9.2f}  max {np.max(x):9.2f}"
    issue_ns = t[:, 3] - t[:, 2]
    print(f"TRACE {dtype} {N}^3 [{label}]  event time per launch {ms.value / steps * 1e3:.1f} us "
          f"({2.0 * N ** 3 / (ms.value / steps) / 1e9:.1f} TFLOP/s), {len(t)} clusters, last launch:")
    print(f"  entry skew across clusters (us)      {q((t[:, 0] - t0) / 1e3)}")
    print(f"  entry -> PDL wait passed (us)        {q((t[:, 1] - t[:, 0]) / 1e3)}")
    print(f"  PDL wait -> first operands (us)      {q((t[:, 2] - t[:, 1]) / 1e3)}")
    print(f"  MMA issue loop (us)                  {q(issue_ns / 1e3)}")
    print(f"  k-blocks / units per cluster         {q(t[:, 5])} / {q(t[:, 7])}")
    print(f"  SM cycles per k-block                {q(t[:, 4] / t[:, 5])}")
    print(f"  SM clock during the loop (MHz)       {q(t[:, 4] / issue_ns * 1e3)}")
    print(f"  last issue -> epilogue done (us)     {q((t[:, 6] - t[:, 3]) / 1e3)}")
    print(f"  whole launch: first entry -> last epilogue done {(t[:, 6].max() - t0) / 1e3:.2f} us", flush=True)
