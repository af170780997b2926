"""GPU bring-up probe for the tcgen05 GEMM (not part of the test suite): runs small and large cases over
every kernel configuration (dtype x tr x CTA-group x BLOCK_N), prints error structure for wrong results
and device timings for the big ones.  Usage: python tools/tc_probe.py [quick|perf|acc]"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from oracle import oracle as O  # noqa: E402
from wgmath_b200._lib import check, lib  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
gemm = w.Gemm.from_device(dev)
L = lib()


def cm(flat, r, c):
    return np.asarray(flat).reshape(c, r).T


def run(M, N, K, tr, dtype, mode, out_dtype="f32"):
    ar, ac = (K, M) if tr else (M, K)
    A = O.uniform(O.SEED_BASE + 1, ar, ac)
    B = O.uniform(O.SEED_BASE + 2, K, N)
    if dtype == "bf16":
        A, B = O.to_bf16_rne(A), O.to_bf16_rne(B)
        ta = w.TensorBuilder.matrix(ar, ac, ST).build_init(dev, O.bf16_bits(A), "bf16")
        tb = w.TensorBuilder.matrix(K, N, ST).build_init(dev, O.bf16_bits(B), "bf16")
    else:
        ta = w.TensorBuilder.matrix(ar, ac, ST).build_init(dev, A)
        tb = w.TensorBuilder.matrix(K, N, ST).build_init(dev, B)
    tc = w.TensorBuilder.matrix(M, N, ST).build_init(dev, np.full(M * N, -1.0, np.float32))
    enc = dev.create_command_encoder()
    with enc.compute_pass("probe", None) as p:
        gemm.dispatch_generic(dev, shapes, p, tc, ta, tb, w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm, f32_mode=mode)
        path = p.last_gemm_path()
    got = cm(tc.read(), M, N).astype(np.float64)
    a = cm(A, ar, ac).astype(np.float64)
    ref = (a.T if tr else a) @ cm(B, K, N).astype(np.float64)
    return got, ref, path


def report(tag, got, ref, tol):
    err = np.abs(got - ref) / np.abs(ref)
    worst = err.max()
    ok = worst < tol
    print(f"{'OK  ' if ok else 'FAIL'} {tag}: max rel err {worst:.3e} (tol {tol:g})", flush=True)
    if not ok:
        bad = err >= tol
        print(f"     bad fraction {bad.mean():.4f}; untouched(-1) fraction {(got == -1.0).mean():.4f}; "
              f"zeros {(got == 0).mean():.4f}; nan {np.isnan(got).mean():.4f}")
        M, N = got.shape
        rb, cb = min(M, 256), min(N, 256)
        blk = bad[:rb, :cb]
        rows = blk.reshape(rb // 8 if rb % 8 == 0 else 1, -1, cb).any(axis=(1, 2)) if rb % 8 == 0 else None
        print("     bad rows (by 8):", "".join("x" if r else "." for r in rows) if rows is not None else "n/a")
        cols = blk.T.reshape(cb // 8 if cb % 8 == 0 else 1, -1, rb).any(axis=(1, 2)) if cb % 8 == 0 else None
        print("     bad cols (by 8):", "".join("x" if c else "." for c in cols) if cols is not None else "n/a")
        print("     got[:4,:4]\n", np.array2string(got[:4, :4], precision=4))
        print("     ref[:4,:4]\n", np.array2string(ref[:4, :4], precision=4))
        print("     ratio[:4,:8]\n", np.array2string((got / ref)[:4, :8], precision=3))
    return ok


def set_cfg(cg, bn):
    os.environ["WGB_TC_CG"] = str(cg)
    os.environ["WGB_TC_BN"] = str(bn)


def quick(cgs=(1, 2)):
    allok = True
    for cg in cgs:
        for bn in (128, 256):
            set_cfg(cg, bn)
            for dtype, mode, tol, name in (("bf16", w.F32Mode.Auto, 1e-4, "bf16"), ("f32", w.F32Mode.Tf32, 3e-3, "tf32"),
                                           ("f32", w.F32Mode.X3Tf32, 1e-5, "3xtf32")):
                for tr in (False, True):
                    for (M, N, K) in ((128, 256, 64), (256, 256, 256), (512, 512, 128), (200, 136, 72), (1024, 768, 520), (512, 512, 2048)):
                        tag = f"cg{cg} bn{bn} {name} tr={int(tr)} {M}x{N}x{K}"
                        try:
                            got, ref, path = run(M, N, K, tr, dtype, mode)
                            ok = report(tag + f" path={path}", got, ref, tol)
                        except Exception as e:
                            print(f"EXC  {tag}: {e!r}", flush=True)
                            return False
                        allok &= ok
    return allok


def timed_gemm(n, dtype, mode, steps=10, out_dtype=None, tr=False):
    out_dtype = out_dtype or dtype
    a = w.TensorBuilder.matrix(n, n, ST).build(dev, dtype)
    b = w.TensorBuilder.matrix(n, n, ST).build(dev, dtype)
    c = w.TensorBuilder.matrix(n, n, ST).build(dev, out_dtype)
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, a, 1)
        w.fill_uniform(dev, p, b, 2)
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    var = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
    with enc.compute_pass("t", None) as p:
        for _ in range(3):
            gemm.dispatch_generic(dev, shapes, p, c, a, b, var, f32_mode=mode)
        check(L.wgb_event_record(e0, p._h))
        for _ in range(steps):
            gemm.dispatch_generic(dev, shapes, p, c, a, b, var, f32_mode=mode)
        check(L.wgb_event_record(e1, p._h))
        path = p.last_gemm_path()
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    return ms.value / steps, path


def perf(cgs=(1, 2)):
    for cg in cgs:
        for bn in (256,):
            set_cfg(cg, bn)
            for n in (1024, 2048, 4096, 8192):
                for tr in (False, True):
                    ms, path = timed_gemm(n, "bf16", w.F32Mode.Auto, steps=10 if n == 4096 else 4, tr=tr)
                    print(f"PERF cg{cg} bn{bn} bf16 tr={int(tr)} n={n}: {ms:.4f} ms  {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s path={path}", flush=True)
            for mode, name in ((w.F32Mode.X3Tf32, "3xtf32"), (w.F32Mode.Tf32, "tf32")):
                ms, path = timed_gemm(4096, "f32", mode, steps=5)
                print(f"PERF cg{cg} bn{bn} {name} n=4096: {ms:.4f} ms  {2 * 4096 ** 3 / ms / 1e9:.1f} TFLOP/s path={path}", flush=True)


def acc(cgs=(1, 2)):
    """Does TMEM accumulation round to nearest?  Long-K 3xTF32 against float64."""
    os.environ.pop("WGB_TC_BN", None)
    for cg in cgs:
        os.environ["WGB_TC_CG"] = str(cg)
        for K in (1024, 4096, 16384, 32768):
            got, ref, path = run(256, 256, K, True, "f32", w.F32Mode.X3Tf32)
            e = (got - ref) / ref
            print(f"ACC cg{cg} 3xtf32 K={K}: max|rel| {np.abs(e).max():.3e} mean rel {e.mean():+.3e} path={path}", flush=True)
            got, ref, path = run(256, 256, K, False, "f32", w.F32Mode.Simt)
            e = (got - ref) / ref
            print(f"ACC      simt   K={K}: max|rel| {np.abs(e).max():.3e} mean rel {e.mean():+.3e} path={path}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "quick"
    t0 = time.time()
    cgs = (int(sys.argv[2]),) if len(sys.argv) > 2 else (1, 2)
    if what == "quick":
        ok = quick(cgs)
        print("QUICK", "ALL OK" if ok else "FAILURES", f"{time.time() - t0:.1f}s")
        sys.exit(0 if ok else 1)
    elif what == "perf":
        perf(cgs)
    elif what == "acc":
        acc(cgs)
