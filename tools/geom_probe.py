"""Device-timed sweep of the batched small-matrix factorizations (wgb_geometry_batch): GB/s of algorithmic bytes
(input matrix + output struct per element) and matrices/s per (op, dim).  python tools/geom_probe.py [log2_n]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wgmath_b200 as w  # noqa: E402
from wgmath_b200 import geometry as G  # noqa: E402
from wgmath_b200._lib import check  # noqa: E402

LG = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
L = w.lib()


NCU = os.environ.get("GEOM_NCU") == "1"   # one launch per (op, dim): the capture list is then one kernel per variant


def timed(fn, steps=10, warmup=3):
    if NCU:
        steps, warmup = 1, 0
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    p = enc.compute_pass("probe", None)
    for _ in range(warmup):
        fn(p)
    dev.poll_wait()
    check(L.wgb_event_record(e0, p._h))
    for _ in range(steps):
        fn(p)
    check(L.wgb_event_record(e1, p._h))
    p.end()
    gpu.queue().submit(enc.finish())
    dev.poll_wait()
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    return ms.value / steps


for dim in (2, 3, 4):
    n = 1 << (LG + 2 if dim == 2 else LG)
    base = np.random.default_rng(dim).random((1 << 16, dim, dim)).astype(np.float32)
    sym = ((base + np.transpose(base, (0, 2, 1))) * 0.5 + dim * np.eye(dim, dtype=np.float32)).astype(np.float32)
    gen_in = w.TensorBuilder.vector(n, ST).build_init(dev, np.tile(G.pack(base), n >> 16), f"mat{dim}")
    sym_in = w.TensorBuilder.vector(n, ST).build_init(dev, np.tile(G.pack(sym), n >> 16), f"mat{dim}")
    ops = [("cholesky", getattr(w, f"WgCholesky{dim}"), sym_in, f"mat{dim}"), ("lu", getattr(w, f"WgLU{dim}"), gen_in, f"lu{dim}"),
           ("qr", getattr(w, f"WgQR{dim}"), gen_in, f"qr{dim}"), ("eig", getattr(w, f"WgSymmetricEigen{dim}"), sym_in, f"eig{dim}")]
    if dim < 4:
        ops.append(("svd", getattr(w, f"WgSvd{dim}"), gen_in, f"svd{dim}"))
    for name, cls, src, odt in ops:
        dst = w.TensorBuilder.vector(n, ST).build(dev, odt)
        sh = cls.from_device(dev)
        ms = timed(lambda p: sh.dispatch(dev, p, src, dst))
        nbytes = n * (G.Matrix[dim].itemsize + cls.OUT_TYPE.itemsize)
        print(f"GEOM {name}{dim} n=2^{int(np.log2(n))}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s  {n / ms / 1e6:7.2f} Gmat/s", flush=True)
        del dst
    dst = w.TensorBuilder.vector(n, ST).build(dev, f"mat{dim}")
    winv = w.WgInv.from_device(dev)
    ms = timed(lambda p: winv.dispatch(dev, p, dim, sym_in, dst))
    print(f"GEOM inv{dim} n=2^{int(np.log2(n))}: {ms * 1e3:8.1f} us  {n * 2 * G.Matrix[dim].itemsize / ms / 1e6:8.1f} GB/s  {n / ms / 1e6:7.2f} Gmat/s", flush=True)
    del gen_in, sym_in, dst
