"""Host-link probe for the e2e leg (wgb_gemm_host): what the PCIe link of the box delivers for the byte counts of one bf16
4096^3 step (H2D 64 MiB, D2H 32 MiB) alone and concurrently, the host GEMM call at several panel counts, and the library
GEMM of the same shape (torch.matmul -> cuBLAS) with the same operand rotation as bench.py for comparison.
Usage: python tools/pcie_probe.py"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from wgmath_b200._lib import check, lib  # noqa: E402

torch.cuda.init()
MiB = 1 << 20
h_in = torch.empty(64 * MiB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(32 * MiB, dtype=torch.uint8).pin_memory()
d_in = torch.empty(64 * MiB, dtype=torch.uint8, device="cuda")
d_out = torch.empty(32 * MiB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def wall(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def chunked(nchunks):
    def f():
        n = d_in.numel() // nchunks
        m = d_out.numel() // nchunks
        for i in range(nchunks):
            with torch.cuda.stream(s1):
                d_in[i * n:(i + 1) * n].copy_(h_in[i * n:(i + 1) * n], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i * m:(i + 1) * m].copy_(d_out[i * m:(i + 1) * m], non_blocking=True)
    return f


t = wall(h2d)
print(f"PCIE h2d 64 MiB alone          {t:.3f} ms  {64 * MiB / t / 1e6:.1f} GB/s")
t = wall(d2h)
print(f"PCIE d2h 32 MiB alone          {t:.3f} ms  {32 * MiB / t / 1e6:.1f} GB/s")
t = wall(both)
print(f"PCIE h2d 64 + d2h 32 concurrent {t:.3f} ms")
for nc in (8, 16):
    t = wall(chunked(nc))
    print(f"PCIE concurrent in {nc} chunks    {t:.3f} ms")

# host GEMM through the C ABI at several panel counts
gpu = w.GpuInstance.new(0)
dev = gpu.device()
L = lib()
n = 4096
hb = n * n * 2
ha, hbb, hc = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
for h in (ha, hbb, hc):
    check(L.wgb_host_alloc(hb, ctypes.byref(h)))
ctypes.memset(ha, 0, hb)
ctypes.memset(hbb, 0, hb)
gemm = w.Gemm.from_device(dev)
for npan in (1, 2, 4, 8, 16):
    def f():
        gemm.dispatch_host(dev, n, n, n, hc, ha, hbb, in_dtype="bf16", out_dtype="bf16", n_panels=npan)
    t = wall(f, 8)
    print(f"HOSTGEMM bf16 4096^3 n_panels={npan:2d}: {t:.3f} ms  {2.0 * n ** 3 / t / 1e9:.1f} TFLOP/s")
hc2 = ctypes.c_void_p()
check(L.wgb_host_alloc(hb, ctypes.byref(hc2)))
for npan in (1, 2, 4, 8, 16):
    steps = 12
    def f():
        for i in range(steps):
            gemm.enqueue_host(dev, n, n, n, hc if i % 2 == 0 else hc2, ha, hbb, in_dtype="bf16", out_dtype="bf16", n_panels=npan)
        dev.poll_wait()
    t = wall(f, 4) / steps
    print(f"HOSTGEMM enqueue x{steps} n_panels={npan:2d}: {t:.3f} ms/product  {2.0 * n ** 3 / t / 1e9:.1f} TFLOP/s")

# library GEMM (cuBLAS through torch.matmul), operands rotated through 4 sets like bench.py
sets = [(torch.rand(n, n, device="cuda").bfloat16(), torch.rand(n, n, device="cuda").bfloat16(),
         torch.empty(n, n, device="cuda", dtype=torch.bfloat16)) for _ in range(4)]
for i in range(10):
    a, b, c = sets[i % 4]
    torch.matmul(a, b, out=c)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(50):
    a, b, c = sets[i % 4]
    torch.matmul(a, b, out=c)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"CUBLAS bf16 4096^3 rotated operands: {ms * 1e3:.1f} us  {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s")
