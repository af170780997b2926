"""Device-timed GEMM for every ordering combination of out / m1 / m2 (wgb_gemm_ord): bf16 4096^3 (MN-major B read in place)
and f32 3xTF32 4096^3 (an N-contiguous m2 goes through the transposing split).  Usage: python tools/ord_probe.py [N]"""
import ctypes
import itertools
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ORD = {0: w.ColumnMajor, 1: w.RowMajor}
gemm = w.Gemm.from_device(dev)
for dtype, steps in (("bf16", 60), ("f32", 12)):
    sets = 4   # rotate operand sets so the inputs do not sit in L2
    a = [w.TensorBuilder.matrix(N, N, ST).build(dev, dtype) for _ in range(sets)]
    b = [w.TensorBuilder.matrix(N, N, ST).build(dev, dtype) for _ in range(sets)]
    c = [w.TensorBuilder.matrix(N, N, ST).build(dev, dtype) for _ in range(sets)]
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        for i in range(sets):
            w.fill_uniform(dev, p, a[i], 1 + i)
            w.fill_uniform(dev, p, b[i], 11 + i)
    for tr, ro, r1, r2 in itertools.product((False, True), (0, 1), (0, 1), (0, 1)):
        var = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
        e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
        check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
        check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
        enc = dev.create_command_encoder()
        with enc.compute_pass("t", None) as p:
            for it in range(steps + 5):
                if it == 5:
                    check(L.wgb_event_record(e0, p._h))
                i = it % sets
                gemm.dispatch_generic(dev, shapes, p, c[i].as_view(ORD[ro]), a[i].as_view(ORD[r1]), b[i].as_view(ORD[r2]), var)
            check(L.wgb_event_record(e1, p._h))
            path = p.last_gemm_path()
        ms = ctypes.c_float()
        check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        t = ms.value / steps
        print(f"ORD {dtype} {N}^3 tr={int(tr)} row-major(out,m1,m2)={ro}{r1}{r2} path {path}: {t:.4f} ms  {2.0 * N ** 3 / t / 1e9:.1f} TFLOP/s", flush=True)
