"""Scan / sort probe (for ncu captures and quick timings): exclusive scan and key/value radix sort at n = 2^26.
Usage: python tools/ss_probe.py [log2 n]"""
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
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
rng = np.random.default_rng(1)
kh = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
keys = w.TensorBuilder.vector(n, ST).build_init(dev, kh, "u32")
vals = w.TensorBuilder.vector(n, ST).build_init(dev, np.arange(n, dtype=np.uint32), "u32")
ok, ov = w.TensorBuilder.vector(n, ST).build(dev, "u32"), w.TensorBuilder.vector(n, ST).build(dev, "u32")
sd = w.TensorBuilder.vector(n, ST).build_init(dev, kh & np.uint32(0xFF), "u32")
ns = w.TensorBuilder.scalar(ST).build_init(dev, np.array([n], np.uint32), "u32")
ps, rs = w.WgPrefixSum.from_device(dev), w.RadixSort.from_device(dev)


NCU = os.environ.get("SS_NCU") == "1"   # one dispatch per case, so an ncu capture holds every kernel once


def timed(fn, steps=10):
    if NCU:
        steps = 1
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("t", None) as p:
        for _ in range(0 if NCU else 3):
            fn(p)
        check(L.wgb_event_record(e0, p._h))
        for _ in range(steps):
            fn(p)
        check(L.wgb_event_record(e1, p._h))
    dev.poll_wait()
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    return ms.value / steps


t = timed(lambda p: ps.dispatch(dev, p, w.PrefixSumWorkspace.new(), sd))
print(f"SCAN n=2^{int(np.log2(n))}: {t * 1e3:.1f} us  {8 * n / t / 1e6:.0f} GB/s")
for bits in (8, 16, 32):
    t = timed(lambda p: rs.dispatch(dev, p, w.RadixSortWorkspace.new(dev), keys, vals, ns, bits, ok, ov))
    print(f"SORT n=2^{int(np.log2(n))} bits={bits}: {t * 1e3:.1f} us  {n / t / 1e6:.2f} Gpair/s")
