"""Differential fuzzing of the C oracle against the reference's own shaders (run through tests/golden/wgsl_interp.py).

Random views — sizes, strides, offsets, batches, ragged dimensions, random (non-zero) padding — over random buffers, every
linalg variant, the factorizations on random and degenerate matrices, scan and sort at random lengths.  Each case is run by
the interpreter on the unmodified WGSL of /root/reference (same composition and dispatch code as make_reference_vectors.py)
and by oracle/*.c; the two output buffers must be identical bit for bit (NaNs: both NaN).  Build-container only.

    python tests/golden/fuzz_reference_vs_oracle.py [seed] [cases per family]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import make_reference_vectors as G  # noqa: E402
import reference_cases as C  # noqa: E402
from oracle import oracle as O  # noqa: E402


def c4(x):
    return (x + 3) // 4 * 4


def view(rng, rows, cols, mats, vec4=True):
    """a column-major view whose whole-vec4-block footprint (rows and columns rounded up to 4) stays inside its own matrix slot"""
    a = 4 if vec4 else 1
    stride = c4(rows) + a * int(rng.integers(0, 3)) if vec4 else rows + int(rng.integers(0, 4))
    stride_mat = stride * (c4(cols) if vec4 else cols) + a * int(rng.integers(0, 4))
    offset = a * int(rng.integers(0, 4))
    length = offset + stride_mat * mats + 8
    return (rows, cols, mats, stride, stride_mat, offset), length


def same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype == np.float32:
        nan = np.isnan(a) & np.isnan(b)
        return np.array_equal(np.where(nan, 0, a.view(np.uint32)), np.where(nan, 0, b.view(np.uint32)))
    return np.array_equal(a, b)


def fuzz_gemm(rng, n):
    bad = 0
    for k in range(n):
        variant = ["gemm", "gemm_tr", "gemm_fast", "gemm_tr_fast"][k % 4 if k % 10 < 2 else k % 2]
        fast = variant.endswith("fast")
        M, N, T = int(rng.integers(1, 24)), int(rng.integers(1, 24)), int(rng.integers(1, 4))
        K = int(rng.choice([256, 512])) if fast else int(rng.integers(1, 40))
        if fast:
            M, N, T = int(rng.integers(1, 9)), int(rng.integers(1, 9)), int(rng.integers(1, 3))
        tr = "tr" in variant
        s1, l1 = view(rng, K, M, T) if tr else view(rng, M, K, T)
        s2, l2 = view(rng, K, N, T)
        so, lo = view(rng, M, N, T)
        if fast:   # the workgroup variants read whole rounds of 64 invocations: K is a multiple of 256 here, nothing to pad
            pass
        case = dict(kind="gemm", name=f"fuzz{k}/{variant}", variant=variant, so=so, s1=s1, s2=s2, len_out=lo, len_1=l1, len_2=l2, zero_k_padding=None)
        want = G.run_gemm(case)
        b = C.inputs(case)
        rc = O.gemm(C.GEMM_VARIANTS[variant], b["out"], O.Shape(*so), b["m1"], O.Shape(*s1), b["m2"], O.Shape(*s2))
        if rc != O.ORC_OK or not same(b["out"], want):
            bad += 1
            print("MISMATCH", case, flush=True)
    return n, bad


def fuzz_gemv(rng, n):
    bad = 0
    for k in range(n):
        variant = ["gemv", "gemv_tr", "gemv_fast", "gemv_tr_fast"][k % 4 if k % 10 < 2 else k % 2]
        fast = variant.endswith("fast")
        M, Cn, T = int(rng.integers(1, 40)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
        K = int(rng.choice([128, 256, 384])) if fast else int(rng.integers(1, 60))
        if fast:
            M = 4 * int(rng.integers(1, 5))                      # gemv.rs:122 asserts out_nrows % 4 == 0
        tr = "tr" in variant
        sm, lm = view(rng, K, M, T) if tr else view(rng, M, K, T)
        sv, lv = view(rng, K, Cn, T)
        so, lo = view(rng, M, Cn, T)
        case = dict(kind="gemv", name=f"fuzz{k}/{variant}", variant=variant, so=so, sm=sm, sv=sv, len_out=lo, len_m=lm, len_v=lv)
        want = G.run_gemv(case)
        b = C.inputs(case)
        rc, ran = O.gemv(C.GEMV_VARIANTS[variant], b["out"], O.Shape(*so), b["m"], O.Shape(*sm), b["v"], O.Shape(*sv))
        if rc != O.ORC_OK or ran != C.GEMV_VARIANTS[variant] or not same(b["out"], want):
            bad += 1
            print("MISMATCH", case, rc, ran, flush=True)
    return n, bad


def fuzz_level1(rng, n):
    bad = 0
    for k in range(n):
        nn = int(rng.integers(0, 400))
        oa, ob = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        op = list(C.OP_ASSIGN)[k % 5]
        case = dict(kind="op_assign", name=f"fuzz{k}/{op}", op=op, sa=(nn, 1, 1, nn, nn, oa), sb=(nn, 1, 1, nn, nn, ob), len_a=nn + oa + 3, len_b=nn + ob + 3, fixture=None)
        want = G.run_op_assign(case)
        b = C.inputs(case)
        rc = O.op_assign(C.OP_ASSIGN[op], b["a"], O.Shape(*case["sa"]), b["b"], O.Shape(*case["sb"]))
        if rc != O.ORC_OK or not same(b["a"], want):
            bad += 1
            print("MISMATCH", case, flush=True)
        nn, off = int(rng.integers(0, 900)), int(rng.integers(0, 9))
        rop = list(C.REDUCE)[k % 5]
        case = dict(kind="reduce", name=f"fuzz{k}/{rop}", op=rop, s=(nn, 1, 1, nn, nn, off), len=max(nn + off, 4))
        want = G.run_reduce(case)
        b = C.inputs(case)
        got = np.array([O.reduce(C.REDUCE[rop], b["x"], O.Shape(*case["s"]))], np.float32)
        if not same(got, want):
            bad += 1
            print("MISMATCH", case, got, want, flush=True)
    return 2 * n, bad


def fuzz_geometry(rng, n):
    """random matrices at several magnitudes plus degenerate ones (zero, rank one, repeated eigenvalues, huge / tiny entries)"""
    bad = 0
    total = 0
    for op, dim in C.geometry_cases():
        mats = rng.random((n, dim, dim), dtype=np.float32) * 2 - 1
        mats *= np.float32(10.0) ** rng.integers(-3, 4, (n, 1, 1)).astype(np.float32)
        if op in ("cholesky", "eig"):
            m64 = mats.astype(np.float64)
            mats = (np.einsum("nki,nkj->nij", m64, m64) + 1e-3 * np.eye(dim)).astype(np.float32)
            mats = (mats + np.transpose(mats, (0, 2, 1))) * np.float32(0.5)
        mats[0] = 0.0
        mats[1] = np.eye(dim, dtype=np.float32) * np.float32(3.0)
        mats[2] = np.outer(np.arange(1, dim + 1), np.arange(1, dim + 1)).astype(np.float32)      # rank one
        mats[3] = np.float32(1e18) * np.eye(dim, dtype=np.float32)
        mats[4] = np.float32(1e-20) * np.ones((dim, dim), np.float32)
        packed = O.geom_pack(mats)
        pr = G.geometry_program(G.geometry_source(op, dim))
        res = np.zeros((n, C.GEOMETRY_OUT_WORDS[(op, dim)]), np.float32)
        pr.dispatch("test", {(0, 0): packed.reshape(-1).view(np.uint8), (0, 1): res.reshape(-1).view(np.uint8)}, (n, 1, 1))
        ref = O.geom_batch(C.GEOMETRY_OPS[op], dim, packed)
        rows = [i for i in range(n) if not same(res[i], ref[i])]
        total += n
        if rows:
            bad += len(rows)
            print("MISMATCH", op, dim, "matrices", rows[:8], flush=True)
    return total, bad


def fuzz_scan_sort(rng, n):
    bad = 0
    for k in range(n):
        nn = int(rng.integers(1, 3000)) if k else 70000          # one length with three levels
        d = rng.integers(0, 2 ** 32, nn, dtype=np.uint64).astype(np.uint32)
        want = G.run_prefix_sum(d.copy())
        got = d.copy()
        O.prefix_sum(got)
        if not same(got, want):
            bad += 1
            print("MISMATCH scan", nn, flush=True)
    sorts = max(1, n // 4)
    for k in range(sorts):
        ln = int(rng.integers(1, 1300))
        n_sort = ln if k % 2 == 0 else int(rng.integers(0, ln + 1))
        bits = int(rng.integers(1, 17))
        keys = rng.integers(0, 2 ** 32, ln, dtype=np.uint64).astype(np.uint32)
        if k % 3 == 0:
            keys %= np.uint32(5)
        vals = rng.integers(0, 2 ** 32, ln, dtype=np.uint64).astype(np.uint32)
        wk, wv, _ = G.run_radix_sort(keys, vals, n_sort, bits)
        ok_, ov = keys.copy(), vals.copy()
        O.radix_sort(keys, vals, n_sort, bits, ok_, ov)
        if not (same(ok_[:n_sort], wk[:n_sort]) and same(ov[:n_sort], wv[:n_sort])):
            bad += 1
            print("MISMATCH sort", ln, n_sort, bits, flush=True)
    return n + sorts, bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rng = np.random.default_rng(seed)
    total = wrong = 0
    for name, fn, count in (("gemm", fuzz_gemm, per), ("gemv", fuzz_gemv, per), ("op_assign + reduce", fuzz_level1, per),
                            ("factorizations", fuzz_geometry, max(8, per // 2)), ("scan + sort", fuzz_scan_sort, max(4, per // 6))):
        t = time.time()
        n, bad = fn(rng, count)
        total, wrong = total + n, wrong + bad
        print(f"fuzz {name:20s}: {n - bad} of {n} cases bit-identical  ({time.time() - t:.0f} s)", flush=True)
    print(f"FUZZ seed {seed}: {total - wrong} of {total} cases bit-identical between the reference's shaders and the oracle", flush=True)
    sys.exit(1 if wrong else 0)
