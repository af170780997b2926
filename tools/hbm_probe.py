"""Device-timed GB/s of the GEMV family at BASELINE configs[3] sizes (quick A/B probe; bench.py is the reference number)."""
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
M, K = 65536, 4096
m = w.TensorBuilder.matrix(M, K, ST).build(dev)
x, y = w.TensorBuilder.vector(K, ST).build(dev), w.TensorBuilder.vector(M, ST).build(dev)
xm, yk = w.TensorBuilder.vector(M, ST).build(dev), w.TensorBuilder.vector(K, ST).build(dev)
x4, y4 = w.TensorBuilder.matrix(K, 4, ST).build(dev), w.TensorBuilder.matrix(M, 4, ST).build(dev)
xm4, yk4 = w.TensorBuilder.matrix(M, 4, ST).build(dev), w.TensorBuilder.matrix(K, 4, ST).build(dev)
enc = dev.create_command_encoder()
with enc.compute_pass("init", None) as p:
    for t, s in ((m, 1), (x, 3), (xm, 3), (x4, 5), (xm4, 6)):
        w.fill_uniform(dev, p, t, s)
gemv = w.Gemv.from_device(dev)
mt = m.reshape((K, M))


def timed(name, nbytes, fn, steps=20):
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("t", None) as p:
        for _ in range(3):
            fn(p)
        check(L.wgb_event_record(e0, p._h))
        for _ in range(steps):
            fn(p)
        check(L.wgb_event_record(e1, p._h))
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    print(f"HBM {name:34s} {ms.value / steps * 1e3:8.1f} us  {nbytes * steps / ms.value / 1e6:8.1f} GB/s", flush=True)


B1 = 4 * (M * K + K + M)
B4 = 4 * (M * K + 4 * K + 4 * M)
timed("gemv 65536x4096", B1, lambda p: gemv.dispatch(dev, shapes, p, y, m, x))
timed("gemv_tr 65536x4096", B1, lambda p: gemv.dispatch_tr(dev, shapes, p, yk, m, xm))
timed("gemv_tr 4096x65536", B1, lambda p: gemv.dispatch_tr(dev, shapes, p, y, mt, x))
timed("gemv 4096x65536", B1, lambda p: gemv.dispatch(dev, shapes, p, yk, mt, xm))
timed("gemv 65536x4096, 4 rhs columns", B4, lambda p: gemv.dispatch(dev, shapes, p, y4, m, x4))
timed("gemv_tr 65536x4096, 4 rhs columns", B4, lambda p: gemv.dispatch_tr(dev, shapes, p, yk4, m, xm4))
