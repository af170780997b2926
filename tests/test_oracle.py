"""CPU tests (-m "not gpu"): pin the oracle (oracle/wgsl_oracle.c).

The reference holds no golden vectors for this path.  What pins its results are
four unit tests that compare against nalgebra; they are replayed here against the oracle with an
independent float64 reference in nalgebra's place, at the reference's own tolerances:
  gemm.rs:144-202    256x256, all four variants, eps 1e-3 (absolute)
  gemv.rs:153-197    1024x1024, all four variants, out pre-filled with random data, eps 1e-3
  op_assign.rs:109-157  LEN=1757, v0[i]=i+0.1, v1[i]=10i+0.1, eps 1e-7
  reduce.rs:139-179  LEN=345, five ops, eps 1e-3
plus the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def colmajor(flat, nrows, ncols):
    return flat.reshape(ncols, nrows).T


def test_shape_struct_is_24_bytes():
    import ctypes
    assert ctypes.sizeof(O.Shape) == 24  # shapes.rs:9-21


@pytest.mark.parametrize("variant", [O.GEMM, O.GEMM_TR, O.GEMM_FAST, O.GEMM_TR_FAST])
def test_reference_gpu_gemm_replay(variant):
    n = 256
    m1 = O.uniform(O.SEED_BASE + 1, n, n)
    m2 = O.uniform(O.SEED_BASE + 2, n, n)
    out = np.zeros(n * n, np.float32)  # gemm.rs:154: pre-zeroed
    assert O.gemm(variant, out, O.shape(n, n), m1, O.shape(n, n), m2, O.shape(n, n)) == O.ORC_OK
    a, b = colmajor(m1, n, n).astype(np.float64), colmajor(m2, n, n).astype(np.float64)
    ref = (a.T if variant in (O.GEMM_TR, O.GEMM_TR_FAST) else a) @ b
    got = colmajor(out, n, n)
    assert np.max(np.abs(got - ref)) < 1e-3          # gemm.rs:200
    assert np.max(np.abs(got - ref) / ref) < 1e-5    # north_star tolerance


@pytest.mark.parametrize("variant", [O.GEMV, O.GEMV_TR, O.GEMV_FAST, O.GEMV_TR_FAST])
def test_reference_gpu_gemv_replay(variant):
    n = 1024
    m = O.uniform(O.SEED_BASE + 1, n, n)
    v = O.uniform(O.SEED_BASE + 3, n)
    out = O.uniform(O.SEED_BASE + 4, n).copy()        # gemv.rs:163: pre-randomised => overwrite semantics
    rc, ran = O.gemv(variant, out, O.shape(n), m, O.shape(n, n), v, O.shape(n))
    assert rc == O.ORC_OK and ran == variant
    a = colmajor(m, n, n).astype(np.float64)
    ref = (a.T if variant in (O.GEMV_TR, O.GEMV_TR_FAST) else a) @ v.astype(np.float64)
    assert np.max(np.abs(out - ref)) < 1e-3           # gemv.rs:195
    assert np.max(np.abs(out - ref) / ref) < 1e-5


@pytest.mark.parametrize("op", [O.OP_ADD, O.OP_SUB, O.OP_MUL, O.OP_DIV, O.OP_COPY])
def test_reference_gpu_op_assign_replay(op):
    n = 1757                                           # op_assign.rs:123 (not a multiple of 64)
    i = np.arange(n, dtype=np.float32)
    v0, v1 = i + np.float32(0.1), i * np.float32(10.0) + np.float32(0.1)
    a = v0.copy()
    assert O.op_assign(op, a, O.shape(n), v1, O.shape(n)) == O.ORC_OK
    ref = {O.OP_ADD: v0 + v1, O.OP_SUB: v0 - v1, O.OP_MUL: v0 * v1, O.OP_DIV: v0 / v1, O.OP_COPY: v1}[op]
    np.testing.assert_array_equal(a, ref)              # stronger than the reference's 1e-7


@pytest.mark.parametrize("op", [O.RED_MIN, O.RED_MAX, O.RED_SUM, O.RED_SQNORM, O.RED_PROD])
def test_reference_gpu_reduce_replay(op):
    n = 345
    x = O.uniform(O.SEED_BASE + 3, n)
    got = O.reduce(op, x, O.shape(n))
    x64 = x.astype(np.float64)
    ref = {O.RED_MIN: x64.min(), O.RED_MAX: x64.max(), O.RED_SUM: x64.sum(), O.RED_PROD: x64.prod(),
           O.RED_SQNORM: (x64 * x64).sum()}[op]
    assert abs(got - ref) < 1e-3                       # reduce.rs:176
    assert abs(got - ref) <= 1e-5 * abs(ref) + 1e-37


def test_reduce_init_constants_and_empty():
    empty = np.zeros(4, np.float32)
    s = O.shape(0)
    assert O.reduce(O.RED_MIN, empty, s) == np.float32(3.4e38)    # reduce.wgsl:40-42, not FLT_MAX / inf
    assert O.reduce(O.RED_MAX, empty, s) == np.float32(-3.4e38)
    assert O.reduce(O.RED_SUM, empty, s) == 0.0 and O.reduce(O.RED_PROD, empty, s) == 1.0


def test_dispatch_rules():
    z = np.zeros(64 * 64, np.float32)
    assert O.gemm(O.GEMM, z.copy(), O.shape(64, 64), z, O.shape(64, 32), z, O.shape(64, 64)) == O.ORC_DIM_MISMATCH  # gemm.rs:91
    assert O.gemm(O.GEMM, z.copy(), O.shape(32, 64), z, O.shape(64, 64), z, O.shape(64, 64)) == O.ORC_DIM_MISMATCH  # :92
    assert O.gemm(O.GEMM, z.copy(), O.shape(64, 64, 1), z, O.shape(32, 32, 2), z, O.shape(32, 64, 1)) == O.ORC_DIM_MISMATCH  # :94
    assert O.op_assign(O.OP_ADD, z.copy(), O.shape(8), z, O.shape(9)) == O.ORC_DIM_MISMATCH                        # op_assign.rs:82
    # gemv.rs:99-104: GemvTrFast falls back to GemvTr unless m.nrows % 128 == 0
    m = O.uniform(1, 64, 64); v = O.uniform(2, 64); out = np.zeros(64, np.float32)
    rc, ran = O.gemv(O.GEMV_TR_FAST, out, O.shape(64), m, O.shape(64, 64), v, O.shape(64))
    assert rc == O.ORC_OK and ran == O.GEMV_TR
    # gemv.rs:122: fast variants assert out rows % 4 == 0
    rc, _ = O.gemv(O.GEMV_FAST, np.zeros(6, np.float32), O.shape(6), np.zeros(6 * 128, np.float32), O.shape(6, 128),
                   np.zeros(128, np.float32), O.shape(128))
    assert rc == O.ORC_DIM_MISMATCH


def test_views_offsets_and_batches():
    """Sub-views the reference's tests never exercise: columns(), rows(), nmats > 1, non-zero offset."""
    rng_a = O.uniform(11, 64, 48 * 3)            # parent: 3 matrices of 64 x 48 back to back
    rng_b = O.uniform(12, 48, 40 * 3)
    M, K, N, T = 32, 48, 20, 3
    s1 = O.Shape(M, K, T, 64, 64 * 48, 8)        # rows 8..40 of each 64 x 48 matrix
    s2 = O.Shape(K, N, T, 48, 48 * 40, 48 * 4)   # columns 4..24 of each 48 x 40 matrix
    so = O.Shape(M, N, T, 36, 36 * 24, 4)        # padded output, offset 4
    out = np.full(4 + 36 * 24 * T, -7.0, np.float32)
    assert O.gemm(O.GEMM, out, so, rng_a, s1, rng_b, s2) == O.ORC_OK
    for t in range(T):
        a = colmajor(rng_a[t * 64 * 48:(t + 1) * 64 * 48], 64, 48)[8:8 + M, :].astype(np.float64)
        b = colmajor(rng_b[t * 48 * 40:(t + 1) * 48 * 40], 48, 40)[:, 4:4 + N].astype(np.float64)
        got = colmajor(out[4 + t * 36 * 24: 4 + (t + 1) * 36 * 24], 36, 24)
        assert np.max(np.abs(got[:M, :N] - a @ b) / (a @ b)) < 1e-5
        assert np.all(got[M:, :] == -7.0) and np.all(got[:, N:] == -7.0)   # nothing outside the view written
    assert np.all(out[:4] == -7.0)
    ref64 = O.gemm_f64(False, M, N, K, T, rng_a, s1, rng_b, s2)
    for t in range(T):
        a = colmajor(rng_a[t * 64 * 48:(t + 1) * 64 * 48], 64, 48)[8:8 + M, :].astype(np.float64)
        b = colmajor(rng_b[t * 48 * 40:(t + 1) * 48 * 40], 48, 40)[:, 4:4 + N].astype(np.float64)
        np.testing.assert_allclose(ref64[t].T, a @ b, rtol=1e-12)


def test_seeded_generator_properties():
    a = O.uniform(O.SEED_BASE + 1, 64, 32)
    assert a.dtype == np.float32 and 0.0 <= a.min() and a.max() < 1.0
    # value is a function of (seed, i, j) only: a shard equals the slice of the whole
    blk = O.uniform(O.SEED_BASE + 1, 16, 8, row0=32, col0=4)
    np.testing.assert_array_equal(colmajor(blk, 16, 8), colmajor(a, 64, 32)[32:48, 4:12])
    assert not np.array_equal(a, O.uniform(O.SEED_BASE + 2, 64, 32))
    b = O.to_bf16_rne(a)
    assert np.all((b.view(np.uint32) & 0xFFFF) == 0) and np.max(np.abs(b - a) / a) <= 2.0 ** -8
    np.testing.assert_array_equal(O.bf16_from_bits(O.bf16_bits(a)), b)


def test_golden_fixtures():
    g = np.load(os.path.join(GOLD, "cfg1_gemm64.npz"))
    for variant, key in [(O.GEMM, "gemm"), (O.GEMM_TR, "gemm_tr")]:
        out = np.zeros(64 * 64, np.float32)
        assert O.gemm(variant, out, O.shape(64, 64), g["m1"], O.shape(64, 64), g["m2"], O.shape(64, 64)) == O.ORC_OK
        assert np.max(np.abs(out - g[key]) / g[key]) < 1e-5
    l1 = np.load(os.path.join(GOLD, "level12.npz"))
    out = np.zeros(128, np.float32)
    rc, _ = O.gemv(O.GEMV, out, O.shape(128), l1["m"], O.shape(128, 64), l1["v"], O.shape(64))
    assert rc == O.ORC_OK and np.max(np.abs(out - l1["gemv"]) / l1["gemv"]) < 1e-5
    out = np.zeros(64, np.float32)
    rc, _ = O.gemv(O.GEMV_TR, out, O.shape(64), l1["m"], O.shape(128, 64), l1["x128"], O.shape(128))
    assert rc == O.ORC_OK and np.max(np.abs(out - l1["gemv_tr"]) / l1["gemv_tr"]) < 1e-5
    for op, key in [(O.RED_MIN, "min"), (O.RED_MAX, "max"), (O.RED_SUM, "sum"), (O.RED_PROD, "prod"), (O.RED_SQNORM, "sqnorm")]:
        got = O.reduce(op, l1["x345"], O.shape(345))
        assert abs(got - l1[key]) <= 1e-5 * abs(l1[key]) + 1e-37


@pytest.mark.parametrize("tr", [0, 2])
def test_row_major_addressing_restatement(tr):
    """orc_gemm_ord / orc_gemv_ord (shape.wgsl:49-57): every ordering combination equals the float64 product, equals the
    literal vec4 restatement when everything is column-major, and never writes outside the output view."""
    rng = np.random.default_rng(7)
    M, N, K, T = 12, 20, 16, 2
    for ro in (0, 1):
        for r1 in (0, 1):
            for r2 in (0, 1):
                a_dims = (K, M) if tr else (M, K)
                A, B = rng.random((T,) + a_dims, dtype=np.float32), rng.random((T, K, N), dtype=np.float32)

                def pack(X, rm, pad, off):
                    r, c = X.shape[1:]
                    ld = (c if rm else r) + pad
                    smat = ld * (r if rm else c) + 4
                    buf = np.full(off + smat * T, np.float32(-7.0))
                    for t in range(T):
                        for i in range(r):
                            for j in range(c):
                                buf[off + t * smat + (i * ld + j if rm else i + j * ld)] = X[t, i, j]
                    return buf, O.Shape(r, c, T, ld, smat, off)
                a, sa = pack(A, r1, 4, 8)
                b, sb = pack(B, r2, 0, 4)
                c0, sc = pack(np.zeros((T, M, N), np.float32), ro, 4, 4)
                out = np.full_like(c0, np.float32(-7.0))
                assert O.gemm_ord(tr, out, sc, ro, a, sa, r1, b, sb, r2) == O.ORC_OK
                ref = np.einsum("tkm,tkn->tmn" if tr else "tmk,tkn->tmn", A.astype(np.float64), B.astype(np.float64))
                want, _ = pack(ref.astype(np.float32), ro, 4, 4)
                live = c0 == 0
                np.testing.assert_allclose(out[live], want[live], rtol=2e-6)
                assert np.all(out[~live] == np.float32(-7.0))
                if (ro, r1, r2) == (0, 0, 0):
                    lit = np.full_like(c0, np.float32(-7.0))
                    assert O.gemm(tr, lit, sc, a, sa, b, sb) == O.ORC_OK
                    np.testing.assert_allclose(lit[live], out[live], rtol=2e-6)
    # gemv with a row-major matrix == gemv_tr / gemv on the transposed column-major memory
    R, C = 24, 40
    Mx, v = rng.random((R, C), dtype=np.float32), rng.random(R if tr else C, dtype=np.float32)
    out = np.zeros(C if tr else R, np.float32)
    sm = O.Shape(R, C, 1, C, R * C, 0)
    assert O.gemv_ord(tr, out, O.shape(out.size), np.ascontiguousarray(Mx).reshape(-1), sm, 1, v, O.shape(v.size)) == O.ORC_OK
    np.testing.assert_allclose(out, (Mx.T if tr else Mx).astype(np.float64) @ v.astype(np.float64), rtol=2e-6)
    bad = np.zeros(3, np.float32)
    assert O.gemv_ord(tr, bad, O.shape(3), np.ascontiguousarray(Mx).reshape(-1), sm, 1, v, O.shape(v.size)) == O.ORC_DIM_MISMATCH
