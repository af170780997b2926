"""GPU parity tests (-m gpu): the CUDA path, called through the reference-shaped host API and the
C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): f32 results within 1e-5 relative of the oracle, bf16 within
1e-2; op_assign is bit-exact (the reference's own bound there is 1e-7, op_assign.rs:154).
The first block replays the reference's four unit tests (gemm.rs:144-202, gemv.rs:153-197,
op_assign.rs:109-157, reduce.rs:139-179); the rest covers what those tests never exercise:
sub-views, offsets, batches, ragged sizes, and the full BASELINE.json sizes through
size-independent properties."""
import ctypes
import os

import numpy as np
import pytest

import wgmath_b200 as w
from oracle import oracle as O
from tests.helpers import SEED_A, SEED_B, SEED_OUT, SEED_V, STORAGE, oshape, rel_err, run_pass, upload

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5
BF16_TOL = 1e-2
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cm(flat, r, c):
    return np.asarray(flat).reshape(c, r).T


# ------------------------------------------------------------------ reference test replays
@pytest.mark.parametrize("variant", [w.GemmVariant.Gemm, w.GemmVariant.GemmTr, w.GemmVariant.GemmFast, w.GemmVariant.GemmTrFast])
def test_gpu_gemm_reference_replay(gpu, shapes, variant):
    n = 256
    m1c, m2c = O.uniform(SEED_A, n, n), O.uniform(SEED_B, n, n)
    gemm = w.Gemm.from_device(gpu.device())
    m1 = w.TensorBuilder.matrix(n, n, STORAGE).build_init(gpu.device(), m1c)
    m2 = w.TensorBuilder.matrix(n, n, STORAGE).build_init(gpu.device(), m2c)
    result = w.TensorBuilder.matrix(n, n, STORAGE).build_init(gpu.device(), np.zeros(n * n, np.float32))
    staging = w.TensorBuilder.matrix(n, n, w.BufferUsages.MAP_READ | w.BufferUsages.COPY_DST).build(gpu.device())
    enc = gpu.device().create_command_encoder()
    p = enc.compute_pass("test", None)
    gemm.dispatch_generic(gpu.device(), shapes, p, result.as_embedded_view(), m1.as_embedded_view(), m2.as_embedded_view(), variant)
    p.end()
    staging.copy_from(enc, result)
    gpu.queue().submit(enc.finish())
    got = staging.read(gpu.device())
    ref = np.zeros(n * n, np.float32)
    assert O.gemm(int(variant), ref, O.shape(n, n), m1c, O.shape(n, n), m2c, O.shape(n, n)) == O.ORC_OK
    assert rel_err(got, ref) < F32_TOL
    a, b = cm(m1c, n, n).astype(np.float64), cm(m2c, n, n).astype(np.float64)
    exact = (a.T if variant in (w.GemmVariant.GemmTr, w.GemmVariant.GemmTrFast) else a) @ b
    assert np.max(np.abs(cm(got, n, n) - exact)) < 1e-3          # the reference's own bound, gemm.rs:200


@pytest.mark.parametrize("variant", [w.GemvVariant.Gemv, w.GemvVariant.GemvTr, w.GemvVariant.GemvFast, w.GemvVariant.GemvTrFast])
def test_gpu_gemv_reference_replay(gpu, shapes, variant):
    n = 1024
    mc, vc, oc = O.uniform(SEED_A, n, n), O.uniform(SEED_V, n), O.uniform(SEED_OUT, n)
    gemv = w.Gemv.from_device(gpu.device())
    m = w.TensorBuilder.matrix(n, n, STORAGE).build_init(gpu.device(), mc)
    v = w.TensorBuilder.vector(n, STORAGE).build_init(gpu.device(), vc)
    result = w.TensorBuilder.vector(n, STORAGE).build_init(gpu.device(), oc)      # pre-randomised: overwrite semantics
    run_pass(gpu, lambda p: gemv.dispatch_generic(gpu.device(), shapes, p, result, m, v, variant))
    got = result.read()
    ref = oc.copy()
    rc, _ = O.gemv(int(variant), ref, O.shape(n), mc, O.shape(n, n), vc, O.shape(n))
    assert rc == O.ORC_OK and rel_err(got, ref) < F32_TOL


@pytest.mark.parametrize("op", list(w.OpAssignVariant))
def test_gpu_op_assign_reference_replay(gpu, shapes, op):
    n = 1757
    i = np.arange(n, dtype=np.float32)
    v0, v1 = i + np.float32(0.1), i * np.float32(10.0) + np.float32(0.1)
    a = w.TensorBuilder.vector(n, STORAGE).build_init(gpu.device(), v0)
    b = w.TensorBuilder.vector(n, STORAGE).build_init(gpu.device(), v1)
    opk = w.OpAssign.new(gpu.device(), op)
    run_pass(gpu, lambda p: opk.dispatch(gpu.device(), shapes, p, a, b))
    ref = v0.copy()
    assert O.op_assign(int(op), ref, O.shape(n), v1, O.shape(n)) == O.ORC_OK
    np.testing.assert_array_equal(a.read(), ref)                  # bit-exact
    np.testing.assert_array_equal(b.read(), v1)                   # b untouched


@pytest.mark.parametrize("op", list(w.ReduceOp))
def test_gpu_reduce_reference_replay(gpu, shapes, op):
    n = 345
    x = O.uniform(SEED_V, n)
    vec = w.TensorBuilder.vector(n, STORAGE).build_init(gpu.device(), x)
    res = w.TensorBuilder.scalar(STORAGE).build(gpu.device())
    red = w.Reduce.new(gpu.device(), op)
    run_pass(gpu, lambda p: red.dispatch(gpu.device(), shapes, p, vec, res))
    got = float(res.read()[0])
    ref = O.reduce(int(op), x, O.shape(n))
    assert abs(got - ref) <= F32_TOL * abs(ref) + 1e-37
    assert abs(got - red.eval_cpu(x)) < 1e-3                      # reduce.rs:176


# ------------------------------------------------------------------ golden fixtures (BASELINE configs[0])
def test_cfg1_gemm_64_golden(gpu, shapes):
    g = np.load(os.path.join(GOLD, "cfg1_gemm64.npz"))
    gemm = w.Gemm.from_device(gpu.device())
    m1, m2 = upload(gpu, g["m1"], (64, 64)), upload(gpu, g["m2"], (64, 64))
    for variant, key in [(w.GemmVariant.Gemm, "gemm"), (w.GemmVariant.GemmTr, "gemm_tr")]:
        out = upload(gpu, np.full(64 * 64, -1.0, np.float32), (64, 64))
        run_pass(gpu, lambda p: gemm.dispatch_generic(gpu.device(), shapes, p, out, m1, m2, variant))
        assert rel_err(out.read(), g[key]) < F32_TOL


def test_level12_golden(gpu, shapes):
    l1 = np.load(os.path.join(GOLD, "level12.npz"))
    gemv = w.Gemv.from_device(gpu.device())
    m, v, x128 = upload(gpu, l1["m"], (128, 64)), upload(gpu, l1["v"], (64,)), upload(gpu, l1["x128"], (128,))
    o1, o2 = upload(gpu, np.zeros(128, np.float32), (128,)), upload(gpu, np.zeros(64, np.float32), (64,))
    run_pass(gpu, lambda p: (gemv.dispatch(gpu.device(), shapes, p, o1, m, v), gemv.dispatch_tr(gpu.device(), shapes, p, o2, m, x128)))
    assert rel_err(o1.read(), l1["gemv"]) < F32_TOL and rel_err(o2.read(), l1["gemv_tr"]) < F32_TOL
    x = upload(gpu, l1["x345"], (345,))
    res = w.TensorBuilder.scalar(STORAGE).build(gpu.device())
    for op, key in [(w.ReduceOp.Min, "min"), (w.ReduceOp.Max, "max"), (w.ReduceOp.Sum, "sum"), (w.ReduceOp.Prod, "prod"), (w.ReduceOp.SqNorm, "sqnorm")]:
        red = w.Reduce.new(gpu.device(), op)
        run_pass(gpu, lambda p: red.dispatch(gpu.device(), shapes, p, x, res))
        assert abs(float(res.read()[0]) - l1[key]) <= F32_TOL * abs(l1[key]) + 1e-37


# ------------------------------------------------------------------ seeded fill == oracle generator
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_fill_uniform_matches_oracle_bitwise(gpu, dtype):
    r, c = 100, 37
    t = w.TensorBuilder.matrix(r, c, STORAGE).build(gpu.device(), dtype)
    run_pass(gpu, lambda p: w.fill_uniform(gpu.device(), p, t, SEED_A, row0=5, col0=9))
    ref = O.uniform(SEED_A, r, c, row0=5, col0=9)
    if dtype == "f32":
        np.testing.assert_array_equal(t.read(), ref)
    else:
        np.testing.assert_array_equal(t.read(), O.bf16_bits(ref))


# ------------------------------------------------------------------ GEMM: views, batches, ragged sizes
def gemm_case(gpu, shapes, M, N, K, T=1, tr=False, pad=(0, 0, 0), off=(0, 0, 0), mode=None, tol=F32_TOL, dtype="f32", out_dtype="f32"):
    """m1 / m2 / out live in padded parents with non-zero offsets; result checked against the float64 reference
    (the oracle's literal WGSL restatement is only defined for multiples of 4)."""
    ar, ac = (K, M) if tr else (M, K)
    lda, ldb, ldc = ar + pad[0], K + pad[1], M + pad[2]
    sa, sb, sc = lda * ac + 8, ldb * N + 8, ldc * N + 12
    A = O.uniform(SEED_A, off[0] + sa * T)
    B = O.uniform(SEED_B, off[1] + sb * T)
    if dtype == "bf16":
        A, B = O.to_bf16_rne(A), O.to_bf16_rne(B)
    s1 = w.ViewShape((ar, ac, T), lda, sa, off[0])
    s2 = w.ViewShape((K, N, T), ldb, sb, off[1])
    so = w.ViewShape((M, N, T), ldc, sc, off[2])
    sentinel = np.float32(-3.0)
    C0 = np.full(off[2] + sc * T, sentinel, np.float32)
    if dtype == "bf16":
        ta, tb = upload(gpu, O.bf16_bits(A), (A.size,), "bf16"), upload(gpu, O.bf16_bits(B), (B.size,), "bf16")
    else:
        ta, tb = upload(gpu, A, (A.size,)), upload(gpu, B, (B.size,))
    tc = upload(gpu, O.bf16_bits(C0) if out_dtype == "bf16" else C0, (C0.size,), out_dtype)
    gemm = w.Gemm.from_device(gpu.device())
    va = w.GpuTensorView(s1, ta.buffer(), dtype, 3)
    vb = w.GpuTensorView(s2, tb.buffer(), dtype, 3)
    vc = w.GpuTensorView(so, tc.buffer(), out_dtype, 3)
    path = []

    def go(p):
        gemm.dispatch_generic(gpu.device(), shapes, p, vc, va, vb, w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm, f32_mode=mode)
        path.append(p.last_gemm_path())
    run_pass(gpu, go)
    got = tc.read()
    if out_dtype == "bf16":
        got = O.bf16_from_bits(got)
    ref = O.gemm_f64(tr, M, N, K, T, A, oshape(s1), B, oshape(s2))      # [t][n][m]
    mask = np.ones(C0.size, bool)
    worst = 0.0
    for t in range(T):
        for n in range(N):
            lo = off[2] + t * sc + n * ldc
            worst = max(worst, rel_err(got[lo:lo + M], ref[t, n]))
            mask[lo:lo + M] = False
    assert worst < tol, f"rel err {worst:.3e} (path {path})"
    assert np.all(got[mask] == sentinel), "wrote outside the output view"
    return path[0]


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (4, 4, 4), (1, 1, 1), (5, 3, 7), (130, 70, 33), (257, 129, 65), (128, 256, 512)])
@pytest.mark.parametrize("tr", [False, True])
def test_gemm_simt_any_shape(gpu, shapes, M, N, K, tr):
    assert gemm_case(gpu, shapes, M, N, K, tr=tr, mode=w.F32Mode.Simt) == 1


@pytest.mark.parametrize("tile", ["32", "64", "128"])
@pytest.mark.parametrize("tr", [False, True])
def test_gemm_simt_every_tile_size(gpu, shapes, tile, tr, monkeypatch):
    """The FFMA family has three tile sizes (128 / 64 / 32) picked by a cost model; force each one over ragged, batched, padded
    views, f32 and bf16."""
    monkeypatch.setenv("WGB_SIMT_TILE", tile)
    for (M, N, K, T) in [(1, 1, 1, 1), (33, 31, 9, 5), (70, 130, 65, 2), (16, 16, 16, 64)]:
        assert gemm_case(gpu, shapes, M, N, K, T=T, tr=tr, pad=(3, 5, 1), off=(7, 9, 3), mode=w.F32Mode.Simt) == 1
    assert gemm_case(gpu, shapes, 45, 52, 24, T=3, tr=tr, pad=(1, 3, 5), off=(1, 1, 1), dtype="bf16", out_dtype="bf16", tol=BF16_TOL) == 1


def test_gemm_batched_small_matrices(gpu, shapes):
    # many small matrices: the cost model must not pad them into 128 x 128 tiles; bf16 batches of <= 128 rows run one CTA per tile
    assert gemm_case(gpu, shapes, 16, 16, 16, T=512) == 1
    assert gemm_case(gpu, shapes, 64, 64, 64, T=96) == 1
    assert gemm_case(gpu, shapes, 128, 128, 64, T=40, dtype="bf16", tol=1e-4) == 2
    assert gemm_case(gpu, shapes, 128, 96, 128, T=7, mode=w.F32Mode.X3Tf32) == 4
    assert gemm_case(gpu, shapes, 128, 96, 128, T=7) == 1          # default f32 mode: small matrices stay on the FFMA tiles
    assert gemm_case(gpu, shapes, 256, 96, 128, T=3) == 4


@pytest.mark.parametrize("tr", [False, True])
def test_gemm_views_offsets_batches_unaligned(gpu, shapes, tr):
    # odd leading dimensions and offsets: not TMA-eligible -> must still be correct (FFMA path)
    assert gemm_case(gpu, shapes, 100, 60, 52, T=3, tr=tr, pad=(3, 5, 1), off=(7, 9, 3)) == 1
    # aligned sub-views of larger parents, batched: TMA-eligible
    gemm_case(gpu, shapes, 192, 136, 96, T=3, tr=tr, pad=(64, 32, 128), off=(16, 32, 64))


def test_gemm_k_zero_writes_zeros(gpu, shapes):
    out = upload(gpu, np.full(8 * 8, 5.0, np.float32), (8, 8))
    a = w.GpuTensorView(w.ViewShape((8, 0, 1), 8, 0, 0), upload(gpu, np.zeros(8, np.float32), (8,)).buffer(), "f32", 3)
    b = w.GpuTensorView(w.ViewShape((0, 8, 1), 0, 0, 0), upload(gpu, np.zeros(8, np.float32), (8,)).buffer(), "f32", 3)
    gemm = w.Gemm.from_device(gpu.device())
    run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, out, a, b))
    np.testing.assert_array_equal(out.read(), np.zeros(64, np.float32))      # gemm.wgsl:89,106-110: sum = 0 is stored


def test_gemm_dimension_mismatch_panics(gpu, shapes):
    gemm = w.Gemm.from_device(gpu.device())
    z = lambda r, c: upload(gpu, np.zeros(r * c, np.float32), (r, c))
    with pytest.raises(w.DimensionMismatch, match="Gemm: dimension mismatch"):
        run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, z(8, 8), z(8, 4), z(8, 8)))
    with pytest.raises(w.DimensionMismatch):
        run_pass(gpu, lambda p: gemm.dispatch_tr(gpu.device(), shapes, p, z(8, 8), z(8, 4), z(8, 8)))
    with pytest.raises(w.WgbError, match="reaches"):
        bad = w.GpuTensorView(w.ViewShape((8, 8, 1), 8, 64, 1), z(8, 8).buffer(), "f32", 3)
        run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, bad, z(8, 8), z(8, 8)))


def test_empty_dispatch_is_skipped(gpu, shapes):
    # kernel.rs:111-113,144: zero-sized binding or empty grid => nothing queued, no error
    e = w.TensorBuilder.vector(0, STORAGE).build(gpu.device())
    opk = w.OpAssign.new(gpu.device(), w.OpAssignVariant.Add)
    n0 = gpu.device().launch_count()
    run_pass(gpu, lambda p: opk.dispatch(gpu.device(), shapes, p, e, e))
    assert gpu.device().launch_count() == n0


# ------------------------------------------------------------------ GEMM on the tensor cores
TC_SHAPES = [(128, 256, 64), (256, 256, 256), (384, 512, 192), (200, 136, 72), (1024, 768, 520), (136, 264, 40)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("tr", [False, True])
def test_gemm_bf16_tcgen05(gpu, shapes, M, N, K, tr):
    path = gemm_case(gpu, shapes, M, N, K, tr=tr, dtype="bf16", out_dtype="f32", tol=1e-4, pad=(8, 8, 4))
    assert path == 2, f"expected the tcgen05 bf16 path, got {path}"
    gemm_case(gpu, shapes, M, N, K, tr=tr, dtype="bf16", out_dtype="bf16", tol=BF16_TOL)


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("tr", [False, True])
def test_gemm_f32_3xtf32(gpu, shapes, M, N, K, tr):
    path = gemm_case(gpu, shapes, M, N, K, tr=tr, mode=w.F32Mode.X3Tf32, tol=F32_TOL, pad=(4, 8, 4))
    assert path == 4, f"expected the 3xTF32 path, got {path}"


@pytest.mark.parametrize("tr", [False, True])
def test_gemm_f32_single_tf32_is_tf32_accurate(gpu, shapes, tr):
    path = gemm_case(gpu, shapes, 256, 256, 512, tr=tr, mode=w.F32Mode.Tf32, tol=2e-3)
    assert path == 3


def test_gemm_f32_default_is_parity_gated(gpu, shapes):
    # WGB_F32_AUTO must meet 1e-5 whichever path it picks
    for (M, N, K) in [(512, 512, 2048), (256, 128, 8192)]:
        gemm_case(gpu, shapes, M, N, K, tol=F32_TOL)


def test_gemm_batched_tc(gpu, shapes):
    gemm_case(gpu, shapes, 256, 192, 128, T=5, dtype="bf16", tol=1e-4)
    gemm_case(gpu, shapes, 256, 192, 128, T=3, mode=w.F32Mode.X3Tf32)


# ------------------------------------------------------------------ GEMV
def gemv_case(gpu, shapes, R, Cc, ncol=1, T=1, tr=False, pad=0, off=(0, 0, 0), variant=None):
    """m is R x Cc (+pad rows in the parent).  out = m v (len R) or tr(m) v (len Cc)."""
    ldm = R + pad
    sm = ldm * Cc + (4 if pad % 4 == 0 else 3)
    klen, olen = (R, Cc) if tr else (Cc, R)
    ldv, ldo = klen + pad, olen + pad
    sv, so_ = ldv * ncol + 8, ldo * ncol + 4
    Mbuf = O.uniform(SEED_A, off[0] + sm * T)
    Vbuf = O.uniform(SEED_V, off[1] + sv * T)
    sentinel = np.float32(-9.0)
    Obuf = np.full(off[2] + so_ * T, sentinel, np.float32)
    vm = w.ViewShape((R, Cc, T), ldm, sm, off[0])
    vv = w.ViewShape((klen, ncol, T), ldv, sv, off[1])
    vo = w.ViewShape((olen, ncol, T), ldo, so_, off[2])
    tm, tv, to = upload(gpu, Mbuf, (Mbuf.size,)), upload(gpu, Vbuf, (Vbuf.size,)), upload(gpu, Obuf, (Obuf.size,))
    gemv = w.Gemv.from_device(gpu.device())
    var = variant if variant is not None else (w.GemvVariant.GemvTr if tr else w.GemvVariant.Gemv)
    run_pass(gpu, lambda p: gemv.dispatch_generic(gpu.device(), shapes, p, w.GpuTensorView(vo, to.buffer(), "f32", 3),
                                                  w.GpuTensorView(vm, tm.buffer(), "f32", 3),
                                                  w.GpuTensorView(vv, tv.buffer(), "f32", 3), var))
    got = to.read()
    mask = np.ones(Obuf.size, bool)
    worst = 0.0
    for t in range(T):
        m = np.stack([Mbuf[off[0] + t * sm + j * ldm: off[0] + t * sm + j * ldm + R] for j in range(Cc)], axis=1).astype(np.float64)
        for c in range(ncol):
            v = Vbuf[off[1] + t * sv + c * ldv: off[1] + t * sv + c * ldv + klen].astype(np.float64)
            ref = (m.T if tr else m) @ v
            lo = off[2] + t * so_ + c * ldo
            worst = max(worst, rel_err(got[lo:lo + olen], ref))
            mask[lo:lo + olen] = False
    assert worst < F32_TOL, f"rel err {worst:.3e}"
    assert np.all(got[mask] == sentinel), "wrote outside the output view"


@pytest.mark.parametrize("tr", [False, True])
@pytest.mark.parametrize("R,Cc", [(4, 4), (1, 1), (128, 128), (1024, 256), (130, 67), (3, 1000), (5000, 12), (64, 4096), (8192, 96)])
def test_gemv_shapes(gpu, shapes, R, Cc, tr):
    gemv_case(gpu, shapes, R, Cc, tr=tr)


@pytest.mark.parametrize("tr", [False, True])
def test_gemv_multi_column_batched_views(gpu, shapes, tr):
    gemv_case(gpu, shapes, 256, 128, ncol=3, T=2, tr=tr, pad=4, off=(8, 4, 12))      # aligned sub-views
    gemv_case(gpu, shapes, 250, 131, ncol=5, T=2, tr=tr, pad=3, off=(1, 2, 3))       # unaligned: scalar path
    gemv_case(gpu, shapes, 512, 512, ncol=2, tr=tr)


def test_gemv_fast_variant_rules(gpu, shapes):
    gemv_case(gpu, shapes, 256, 256, variant=w.GemvVariant.GemvFast)
    gemv_case(gpu, shapes, 256, 256, tr=True, variant=w.GemvVariant.GemvTrFast)
    gemv_case(gpu, shapes, 100, 256, tr=True, variant=w.GemvVariant.GemvTrFast)      # gemv.rs:99-104 fallback: still valid
    with pytest.raises(w.DimensionMismatch):                                         # gemv.rs:122
        gemv_case(gpu, shapes, 6, 128, variant=w.GemvVariant.GemvFast)
    gemv = w.Gemv.from_device(gpu.device())
    z = lambda *s: upload(gpu, np.zeros(int(np.prod(s)), np.float32), s)
    with pytest.raises(w.DimensionMismatch, match="Gemv: dimension mismatch"):       # gemv.rs:89
        run_pass(gpu, lambda p: gemv.dispatch(gpu.device(), shapes, p, z(8), z(8, 4), z(5)))


# ------------------------------------------------------------------ op_assign / reduce / dot / column reduce
@pytest.mark.parametrize("n", [1, 3, 4, 5, 63, 1757, (1 << 20) + 3])
@pytest.mark.parametrize("offs", [(0, 0), (1, 0), (0, 1), (3, 2), (4, 8)])
def test_op_assign_lengths_and_misalignment(gpu, shapes, n, offs):
    total = n + 16
    a0 = O.uniform(SEED_A, total) + np.float32(0.5)
    b0 = O.uniform(SEED_B, total) + np.float32(0.5)
    for op in w.OpAssignVariant:
        ta, tb = upload(gpu, a0, (total,)), upload(gpu, b0, (total,))
        va = w.GpuTensorView(w.ViewShape((n, 1, 1), total, total, offs[0]), ta.buffer(), "f32", 1)
        vb = w.GpuTensorView(w.ViewShape((n, 1, 1), total, total, offs[1]), tb.buffer(), "f32", 1)
        opk = w.OpAssign.new(gpu.device(), op)
        run_pass(gpu, lambda p: opk.dispatch(gpu.device(), shapes, p, va, vb))
        ref = a0.copy()
        assert O.op_assign(int(op), ref, O.Shape(n, 1, 1, total, total, offs[0]), b0, O.Shape(n, 1, 1, total, total, offs[1])) == O.ORC_OK
        np.testing.assert_array_equal(ta.read(), ref)


def test_op_assign_dimension_mismatch(gpu, shapes):
    a, b = upload(gpu, np.zeros(8, np.float32), (8,)), upload(gpu, np.zeros(9, np.float32), (9,))
    with pytest.raises(w.DimensionMismatch, match="Op-assign: dimension mismatch"):
        run_pass(gpu, lambda p: w.OpAssign.new(gpu.device(), w.OpAssignVariant.Add).dispatch(gpu.device(), shapes, p, a, b))


@pytest.mark.parametrize("n,off", [(0, 0), (1, 0), (127, 1), (128, 2), (4099, 3), ((1 << 22) + 5, 1)])
def test_reduce_all_ops(gpu, shapes, n, off):
    total = n + 8
    x = O.uniform(SEED_V, total)
    tx = upload(gpu, x, (total,))
    view = w.GpuTensorView(w.ViewShape((n, 1, 1), total, total, off), tx.buffer(), "f32", 1)
    res = w.TensorBuilder.scalar(STORAGE).build(gpu.device())
    xs = x[off:off + n].astype(np.float64)
    for op in w.ReduceOp:
        if op == w.ReduceOp.Prod and n > 4099:
            continue  # underflows to 0 on both sides; nothing to compare
        red = w.Reduce.new(gpu.device(), op)
        run_pass(gpu, lambda p: red.dispatch(gpu.device(), shapes, p, view, res))
        got = float(res.read()[0])
        ref = O.reduce(int(op), x, O.Shape(n, 1, 1, total, total, off))
        exact = {w.ReduceOp.Min: min(xs.min(initial=3.4e38), float(np.float32(3.4e38))), w.ReduceOp.Max: max(xs.max(initial=-3.4e38), float(np.float32(-3.4e38))),
                 w.ReduceOp.Sum: xs.sum(), w.ReduceOp.Prod: xs.prod(), w.ReduceOp.SqNorm: (xs * xs).sum()}[op]
        assert abs(got - ref) <= F32_TOL * abs(ref) + 1e-37, (op, got, ref)
        assert abs(got - exact) <= F32_TOL * abs(exact) + 1e-37, (op, got, exact)


def test_reduce_is_deterministic(gpu, shapes):
    n = (1 << 21) + 17
    tx = upload(gpu, O.uniform(SEED_V, n), (n,))
    res = w.TensorBuilder.scalar(STORAGE).build(gpu.device())
    red = w.Reduce.new(gpu.device(), w.ReduceOp.Sum)
    vals = set()
    for _ in range(5):
        run_pass(gpu, lambda p: red.dispatch(gpu.device(), shapes, p, tx, res))
        vals.add(res.read()[0].tobytes())
    assert len(vals) == 1


@pytest.mark.parametrize("n,offs", [(1, (0, 0)), (1000, (1, 2)), ((1 << 20) + 1, (0, 3)), (4096, (4, 4))])
def test_dot(gpu, shapes, n, offs):
    total = n + 8
    a, b = O.uniform(SEED_A, total), O.uniform(SEED_B, total)
    ta, tb = upload(gpu, a, (total,)), upload(gpu, b, (total,))
    res = w.TensorBuilder.scalar(STORAGE).build(gpu.device())
    va = w.GpuTensorView(w.ViewShape((n, 1, 1), total, total, offs[0]), ta.buffer(), "f32", 1)
    vb = w.GpuTensorView(w.ViewShape((n, 1, 1), total, total, offs[1]), tb.buffer(), "f32", 1)
    run_pass(gpu, lambda p: w.Dot.new(gpu.device()).dispatch(gpu.device(), shapes, p, va, vb, res))
    ref = float(a[offs[0]:offs[0] + n].astype(np.float64) @ b[offs[1]:offs[1] + n].astype(np.float64))
    assert abs(float(res.read()[0]) - ref) <= F32_TOL * abs(ref)


@pytest.mark.parametrize("R,Cc,T,pad,off", [(345, 7, 1, 0, 0), (4096, 33, 2, 4, 8), (100, 1000, 1, 3, 1), (2, 5, 3, 0, 0)])
def test_reduce_columns_equals_one_reduce_per_column(gpu, shapes, R, Cc, T, pad, off):
    ld, sm = R + pad, (R + pad) * Cc + 5
    buf = O.uniform(SEED_A, off + sm * T) + np.float32(0.25)
    tm = upload(gpu, buf, (buf.size,))
    view = w.GpuTensorView(w.ViewShape((R, Cc, T), ld, sm, off), tm.buffer(), "f32", 3)
    out = upload(gpu, np.zeros(Cc * T, np.float32), (Cc * T,))
    for op in w.ReduceOp:
        if op == w.ReduceOp.Prod and R > 400:
            continue
        red = w.Reduce.new(gpu.device(), op)
        run_pass(gpu, lambda p: red.dispatch_columns(gpu.device(), shapes, p, view, out))
        got = out.read()
        for t in range(T):
            for j in range(Cc):
                # the reference's way: Reduce over GpuMatrix::column(j)  (tensor.rs:574-585)
                ref = O.reduce(int(op), buf, O.Shape(R, 1, 1, 1, 1, off + t * sm + j * ld))
                assert abs(got[t * Cc + j] - ref) <= F32_TOL * abs(ref) + 1e-37


# ------------------------------------------------------------------ BASELINE sizes through properties
def test_cfg4_gemv_full_size_properties(gpu, shapes):
    """f32 GEMV 65536 x 4096 (BASELINE configs[3]): too big for the scalar oracle in seconds, so check
    (1) m e_k = column k exactly, (2) linearity m(x+y) = mx + my, (3) sampled rows against float64."""
    M, K = 65536, 4096
    dev = gpu.device()
    m = w.TensorBuilder.matrix(M, K, STORAGE).build(dev)
    run_pass(gpu, lambda p: w.fill_uniform(dev, p, m, SEED_A))
    gemv = w.Gemv.from_device(dev)
    x, y = O.uniform(SEED_V, K), O.uniform(SEED_V + 100, K)
    ek = np.zeros(K, np.float32); ek[1234] = 1.0
    vx, vy, vxy, vek = (upload(gpu, a, (K,)) for a in (x, y, x + y, ek))
    ox, oy, oxy, oek = (w.TensorBuilder.vector(M, STORAGE).build(dev) for _ in range(4))
    run_pass(gpu, lambda p: [gemv.dispatch(dev, shapes, p, o, m, v) for o, v in ((ox, vx), (oy, vy), (oxy, vxy), (oek, vek))])
    np.testing.assert_array_equal(oek.read(), O.uniform(SEED_A, M, 1, col0=1234))
    rx, ry, rxy = ox.read(), oy.read(), oxy.read()
    assert rel_err(rxy, rx.astype(np.float64) + ry.astype(np.float64)) < 5e-6
    rows = np.array([0, 1, 77, 4095, 4096, 32768, 65535])
    mr = np.stack([O.uniform(SEED_A, 1, K, row0=int(r)) for r in rows]).astype(np.float64)
    assert rel_err(rx[rows], mr @ x.astype(np.float64)) < F32_TOL
    # gemv_tr on the same matrix: tr(m) u, u = ones  ==> column sums == Reduce(Sum) over each column
    ones = upload(gpu, np.ones(M, np.float32), (M,))
    ot = w.TensorBuilder.vector(K, STORAGE).build(dev)
    oc = w.TensorBuilder.vector(K, STORAGE).build(dev)
    red = w.Reduce.new(dev, w.ReduceOp.Sum)
    run_pass(gpu, lambda p: (gemv.dispatch_tr(dev, shapes, p, ot, m, ones), red.dispatch_columns(dev, shapes, p, m, oc)))
    assert rel_err(ot.read(), oc.read()) < 5e-6
    col = O.uniform(SEED_A, M, 1, col0=4095).astype(np.float64)
    assert abs(float(ot.read()[4095]) - col.sum()) <= F32_TOL * col.sum()


def test_cfg3_bf16_gemm_4096_properties(gpu, shapes):
    """bf16 GEMM 4096^3 (BASELINE configs[2]): (1) A I = A bit-exactly, (2) sampled rows vs float64 on the same
    bf16-rounded inputs within 1e-2, (3) tr(A) B agrees with the same samples."""
    n = 4096
    dev = gpu.device()
    a = w.TensorBuilder.matrix(n, n, STORAGE).build(dev, "bf16")
    b = w.TensorBuilder.matrix(n, n, STORAGE).build(dev, "bf16")
    run_pass(gpu, lambda p: (w.fill_uniform(dev, p, a, SEED_A), w.fill_uniform(dev, p, b, SEED_B)))
    eye = np.zeros((n, n), np.float32); np.fill_diagonal(eye, 1.0)
    ident = upload(gpu, O.bf16_bits(eye.reshape(-1)), (n, n), "bf16")
    out = w.TensorBuilder.matrix(n, n, STORAGE).build(dev, "bf16")
    gemm = w.Gemm.from_device(dev)
    run_pass(gpu, lambda p: gemm.dispatch(dev, shapes, p, out, a, ident))
    np.testing.assert_array_equal(out.read(), a.read())
    outf = w.TensorBuilder.matrix(n, n, STORAGE).build(dev, "f32")
    path = []
    run_pass(gpu, lambda p: (gemm.dispatch(dev, shapes, p, outf, a, b), path.append(p.last_gemm_path())))
    assert path == [2]
    got = cm(outf.read(), n, n)
    rows = np.array([0, 5, 127, 128, 2049, 4095])
    A = np.stack([O.to_bf16_rne(O.uniform(SEED_A, 1, n, row0=int(r))) for r in rows]).astype(np.float64)
    B = cm(O.to_bf16_rne(O.uniform(SEED_B, n, n)), n, n).astype(np.float64)
    assert rel_err(got[rows, :], A @ B) < 1e-4           # f32 output of bf16 inputs: only accumulation error
    # the bench configuration itself: bf16 OUT, random operands, whatever BLOCK_N / CTA group the heuristic picks (256 / 2)
    cfg = []
    run_pass(gpu, lambda p: (gemm.dispatch(dev, shapes, p, out, a, b), cfg.append(p.last_gemm_config())))
    assert (cfg[0]["bn"], cfg[0]["cg"], cfg[0]["out_dtype"]) == (256, 2, 1)
    got16 = cm(O.bf16_from_bits(out.read()), n, n)
    assert rel_err(got16[rows, :], A @ B) < BF16_TOL
    run_pass(gpu, lambda p: gemm.dispatch_tr(dev, shapes, p, outf, a, b))
    gt = cm(outf.read(), n, n)
    cols = np.stack([O.to_bf16_rne(O.uniform(SEED_A, n, 1, col0=int(r))) for r in rows]).astype(np.float64)   # columns of A = rows of tr(A)
    assert rel_err(gt[rows, :], cols @ B) < 1e-4


def test_cfg2_f32_gemm_large_sampled(gpu, shapes):
    """f32 GEMM N = 2048 (inside BASELINE configs[1]'s sweep): sampled rows vs float64, 1e-5."""
    n = 2048
    dev = gpu.device()
    a, b, c = (w.TensorBuilder.matrix(n, n, STORAGE).build(dev) for _ in range(3))
    run_pass(gpu, lambda p: (w.fill_uniform(dev, p, a, SEED_A), w.fill_uniform(dev, p, b, SEED_B)))
    gemm = w.Gemm.from_device(dev)
    run_pass(gpu, lambda p: gemm.dispatch(dev, shapes, p, c, a, b))
    got = cm(c.read(), n, n)
    rows = np.array([0, 3, 129, 1024, 2047])
    A = np.stack([O.uniform(SEED_A, 1, n, row0=int(r)) for r in rows]).astype(np.float64)
    B = cm(O.uniform(SEED_B, n, n), n, n).astype(np.float64)
    assert rel_err(got[rows, :], A @ B) < F32_TOL


@pytest.mark.parametrize("tr", [False, True])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_gemm_host_buffers_pipelined(gpu, shapes, tr, dtype):
    """wgb_gemm_host == build_init + dispatch + read (the reference tests' sequence, gemm.rs:156-193)."""
    M, N, K = 384, 1100, 520          # N spans several column panels, the last one ragged
    ar, ac = (K, M) if tr else (M, K)
    A, B = O.uniform(SEED_A, ar, ac), O.uniform(SEED_B, K, N)
    if dtype == "bf16":
        A, B = O.to_bf16_rne(A), O.to_bf16_rne(B)
        ha, hb = O.bf16_bits(A), O.bf16_bits(B)
    else:
        ha, hb = A, B
    out = np.full(M * N, -1.0, np.float32)
    gemm = w.Gemm.from_device(gpu.device())
    gemm.dispatch_host(gpu.device(), M, N, K, out, ha, hb, w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm,
                       in_dtype=dtype, out_dtype="f32", n_panels=3)
    a = cm(A, ar, ac).astype(np.float64)
    ref = (a.T if tr else a) @ cm(B, K, N).astype(np.float64)
    assert rel_err(cm(out, M, N), ref) < (1e-4 if dtype == "bf16" else F32_TOL)


def test_gemm_host_enqueue_keeps_products_apart(gpu, shapes):
    """wgb_gemm_host_enqueue: five products in flight over two device slots, each with its own inputs and output; every
    output must be its own product (no slot is reused before its download has left)."""
    M, N, K = 256, 520, 264
    dev = gpu.device()
    gemm = w.Gemm.from_device(dev)
    ins, outs = [], []
    for i in range(5):
        A, B = O.uniform(SEED_A + 16 * i, M, K), O.uniform(SEED_B + 16 * i, K, N)
        out = np.full(M * N, -1.0, np.float32)
        ins.append((A, B))
        outs.append(out)
        gemm.enqueue_host(dev, M, N, K, out, A, B, n_panels=2 + i % 2)
    dev.poll_wait()
    for (A, B), out in zip(ins, outs):
        ref = cm(A, M, K).astype(np.float64) @ cm(B, K, N).astype(np.float64)
        assert rel_err(cm(out, M, N), ref) < F32_TOL
    # and the blocking form still works after enqueued ones
    out = np.full(M * N, -1.0, np.float32)
    gemm.dispatch_host(dev, M, N, K, out, ins[0][0], ins[0][1])
    assert rel_err(cm(out, M, N), cm(ins[0][0], M, K).astype(np.float64) @ cm(ins[0][1], K, N).astype(np.float64)) < F32_TOL


def test_gemm_host_enqueue_sizes_change_while_products_are_in_flight(gpu, shapes):
    """Large products followed at once by small ones (and back): the two device operand slots sit at fixed offsets, so a small
    product's slot can never land inside a large product that is still multiplying or downloading."""
    dev = gpu.device()
    gemm = w.Gemm.from_device(dev)
    jobs = []
    big, small = (1536, 2048, 1024), (128, 264, 136)
    for i, (M, N, K) in enumerate([big, small, small, big, small, big, big, small]):
        A, B = O.uniform(SEED_A + 32 * i, M, K), O.uniform(SEED_B + 32 * i, K, N)
        out = np.full(M * N, -1.0, np.float32)
        gemm.enqueue_host(dev, M, N, K, out, A, B)
        jobs.append((M, N, K, A, B, out))
    dev.poll_wait()
    for M, N, K, A, B, out in jobs:
        ref = cm(A, M, K).astype(np.float64) @ cm(B, K, N).astype(np.float64)
        assert rel_err(cm(out, M, N), ref) < F32_TOL, (M, N, K)


def test_graph_refuses_to_replay_after_a_workspace_reallocation(gpu, shapes, monkeypatch):
    """A recorded 3xTF32 GEMM in its split-kernel form carries pointers into the context's operand-split workspace.  A larger
    eager product reallocates it; replaying the old graph would read freed memory, so wgb_graph_launch must fail instead (and a
    fresh recording works)."""
    monkeypatch.setenv("WGB_TF32_FUSED_SPLIT", "0")      # (the in-kernel split needs no workspace)
    dev = w.GpuInstance.new().device()          # own context: its workspaces start empty
    sh = w.ViewShapeBuffers.new()
    gemm = w.Gemm.from_device(dev)

    def mats(n):
        ta = w.TensorBuilder.matrix(n, n, STORAGE).build_init(dev, O.uniform(SEED_A, n, n))
        tb = w.TensorBuilder.matrix(n, n, STORAGE).build_init(dev, O.uniform(SEED_B, n, n))
        return ta, tb, w.TensorBuilder.matrix(n, n, STORAGE).build(dev)

    def run(fn):
        enc = dev.create_command_encoder()
        with enc.compute_pass("t", None) as p:
            fn(p)
    a, b, c = mats(256)
    run(lambda p: gemm.dispatch_generic(dev, sh, p, c, a, b, w.GemmVariant.Gemm, f32_mode=w.F32Mode.X3Tf32))   # warm
    with dev.capture() as cap:
        run(lambda p: gemm.dispatch_generic(dev, sh, p, c, a, b, w.GemmVariant.Gemm, f32_mode=w.F32Mode.X3Tf32))
    cap.graph.launch()
    want = c.read()
    a2, b2, c2 = mats(1024)
    run(lambda p: gemm.dispatch_generic(dev, sh, p, c2, a2, b2, w.GemmVariant.Gemm, f32_mode=w.F32Mode.X3Tf32))   # grows the workspace
    dev.poll_wait()
    with pytest.raises(w.WgbError, match="reallocated"):
        cap.graph.launch()
    with dev.capture() as cap2:
        run(lambda p: gemm.dispatch_generic(dev, sh, p, c, a, b, w.GemmVariant.Gemm, f32_mode=w.F32Mode.X3Tf32))
    cap2.graph.launch()
    np.testing.assert_array_equal(c.read(), want)


@pytest.mark.parametrize("op", [w.OpAssignVariant.Add, w.OpAssignVariant.Sub, w.OpAssignVariant.Mul, w.OpAssignVariant.Div])
@pytest.mark.parametrize("M,N,K,mode", [(256, 384, 192, None), (100, 60, 52, w.F32Mode.Simt), (512, 512, 1024, w.F32Mode.Tf32)])
def test_gemm_fused_op_assign_equals_the_two_dispatch_chain(gpu, shapes, op, M, N, K, mode):
    """wgb_gemm_op == Gemm::dispatch then OpAssign::dispatch(out, operand) (the reference's separate dispatches)."""
    dev = gpu.device()
    A, B = O.uniform(SEED_A, M, K), O.uniform(SEED_B, K, N)
    E = O.uniform(SEED_V, M * N) + np.float32(0.5)
    ta, tb, te = upload(gpu, A, (M, K)), upload(gpu, B, (K, N)), upload(gpu, E, (M, N))
    fused = upload(gpu, np.zeros(M * N, np.float32), (M, N))
    chain = upload(gpu, np.zeros(M * N, np.float32), (M, N))
    gemm = w.Gemm.from_device(dev)
    opk = w.OpAssign.new(dev, op)

    def go(p):
        gemm.dispatch_op(dev, shapes, p, fused, ta, tb, op, te, f32_mode=mode)
        gemm.dispatch_generic(dev, shapes, p, chain, ta, tb, w.GemmVariant.Gemm, f32_mode=mode)
        opk.dispatch(dev, shapes, p, chain.reshape((M * N,)), te.reshape((M * N,)))
    run_pass(gpu, go)
    np.testing.assert_array_equal(fused.read(), chain.read())          # same kernel arithmetic, then the same IEEE op
    if mode != w.F32Mode.Tf32:
        ref = np.zeros(M * N, np.float32)
        if M % 4 == 0 and N % 4 == 0 and K % 4 == 0:
            assert O.gemm(O.GEMM, ref, O.shape(M, N), A, O.shape(M, K), B, O.shape(K, N)) == O.ORC_OK
        else:
            ref = (cm(A, M, K).astype(np.float64) @ cm(B, K, N).astype(np.float64)).T.reshape(-1).astype(np.float32)
        assert O.op_assign(int(op), ref, O.shape(M * N), E, O.shape(M * N)) == O.ORC_OK
        assert rel_err(fused.read(), ref) < F32_TOL


def test_gemm_fused_accumulate_in_place(gpu, shapes):
    """operand == out: out += m1 * m2 (bf16 operands, f32 accumulate / output), twice."""
    dev = gpu.device()
    n = 256
    A, B = O.to_bf16_rne(O.uniform(SEED_A, n, n)), O.to_bf16_rne(O.uniform(SEED_B, n, n))
    ta, tb = upload(gpu, O.bf16_bits(A), (n, n), "bf16"), upload(gpu, O.bf16_bits(B), (n, n), "bf16")
    out = upload(gpu, np.ones(n * n, np.float32), (n, n))
    gemm = w.Gemm.from_device(dev)
    run_pass(gpu, lambda p: [gemm.dispatch_op(dev, shapes, p, out, ta, tb, w.OpAssignVariant.Add, out) for _ in range(2)])
    ref = 1.0 + 2.0 * (cm(A, n, n).astype(np.float64) @ cm(B, n, n).astype(np.float64))
    assert rel_err(cm(out.read(), n, n), ref) < 1e-4
    with pytest.raises(w.WgbError):
        run_pass(gpu, lambda p: gemm.dispatch_op(dev, shapes, p, out, ta, tb, w.OpAssignVariant.Copy, out))


@pytest.mark.parametrize("op", [w.OpAssignVariant.Add, w.OpAssignVariant.Sub, w.OpAssignVariant.Mul, w.OpAssignVariant.Div])
@pytest.mark.parametrize("variant,M,K,C", [(w.GemvVariant.Gemv, 1024, 1024, 1), (w.GemvVariant.GemvTr, 1024, 1024, 1),
                                           (w.GemvVariant.Gemv, 130, 4100, 3), (w.GemvVariant.GemvTr, 77, 9000, 2),
                                           (w.GemvVariant.Gemv, 40000, 36, 1)])
def test_gemv_fused_op_assign_equals_the_two_dispatch_chain(gpu, shapes, op, variant, M, K, C):
    """wgb_gemv_op == Gemv::dispatch then OpAssign::dispatch(out, operand): bit-identical to the unfused GPU chain (every
    reduction path: single CTA per tile, split reduction with last-CTA fold, ragged scalar tiles, multi-column), and within the
    f32 tolerance of the oracle's gemv + op_assign."""
    dev = gpu.device()
    tr = variant == w.GemvVariant.GemvTr
    mr, mc = (K, M) if tr else (M, K)                       # m is [M x K], or [K x M] for the transposed product
    Mm, V = O.uniform(SEED_A, mr, mc), O.uniform(SEED_V, K, C)
    E = O.uniform(SEED_OUT, M, C) + np.float32(0.5)
    tm, tv, te = upload(gpu, Mm, (mr, mc)), upload(gpu, V, (K, C)), upload(gpu, E, (M, C))
    fused = upload(gpu, np.full(M * C, -3.0, np.float32), (M, C))
    chain = upload(gpu, np.full(M * C, -3.0, np.float32), (M, C))
    gemv, opk = w.Gemv.from_device(dev), w.OpAssign.new(dev, op)

    def go(p):
        gemv.dispatch_op(dev, shapes, p, fused, tm, tv, op, te, variant)
        gemv.dispatch_generic(dev, shapes, p, chain, tm, tv, variant)
        opk.dispatch(dev, shapes, p, chain.reshape((M * C,)), te.reshape((M * C,)))
    run_pass(gpu, go)
    np.testing.assert_array_equal(fused.read(), chain.read())
    a64 = cm(Mm, mr, mc).astype(np.float64)
    prod = ((a64.T if tr else a64) @ cm(V, K, C).astype(np.float64)).T.reshape(-1).astype(np.float32)
    ref = prod.copy()
    assert O.op_assign(int(op), ref, O.shape(M * C), E, O.shape(M * C)) == O.ORC_OK
    # (m*v) - e cancels for some elements: compare against the magnitude of the operands, like the reference's abs-eps tests
    scale = np.maximum(np.abs(ref.astype(np.float64)), np.abs(prod.astype(np.float64)))
    assert float(np.max(np.abs(fused.read().astype(np.float64) - ref) / scale)) < F32_TOL


def test_gemv_fused_residual_update_and_errors(gpu, shapes):
    """operand == out: out = m * v + out (the residual update), applied twice; Copy, row mismatch and partial overlap are rejected."""
    dev = gpu.device()
    M, K = 2048, 512
    Mm, V = O.uniform(SEED_A, M, K), O.uniform(SEED_V, K)
    tm, tv = upload(gpu, Mm, (M, K)), upload(gpu, V, (K,))
    out = upload(gpu, np.ones(M, np.float32), (M,))
    gemv = w.Gemv.from_device(dev)
    run_pass(gpu, lambda p: [gemv.dispatch_op(dev, shapes, p, out, tm, tv, w.OpAssignVariant.Add, out) for _ in range(2)])
    ref = 1.0 + 2.0 * (cm(Mm, M, K).astype(np.float64) @ V.astype(np.float64))
    assert rel_err(out.read(), ref) < F32_TOL
    with pytest.raises(w.WgbError):
        run_pass(gpu, lambda p: gemv.dispatch_op(dev, shapes, p, out, tm, tv, w.OpAssignVariant.Copy, out))
    big = upload(gpu, np.zeros(M + 64, np.float32), (M + 64,))
    with pytest.raises(w.WgbError):                                     # overlapping but not the same view
        run_pass(gpu, lambda p: gemv.dispatch_op(dev, shapes, p, big.rows(0, M), tm, tv, w.OpAssignVariant.Add, big.rows(64, M)))
    with pytest.raises(w.DimensionMismatch):                            # op_assign.rs:82-86
        run_pass(gpu, lambda p: gemv.dispatch_op(dev, shapes, p, out, tm, tv, w.OpAssignVariant.Add, big))
    # K == 0: the product is the zero vector (gemv.wgsl:74,88), so out = 0 + operand
    e = upload(gpu, O.uniform(SEED_OUT, M), (M,))
    empty_m, empty_v = upload(gpu, np.zeros(4, np.float32), (4,)).reshape((M, 0)), upload(gpu, np.zeros(4, np.float32), (4,)).reshape((0,))
    run_pass(gpu, lambda p: gemv.dispatch_op(dev, shapes, p, out, empty_m, empty_v, w.OpAssignVariant.Add, e))
    np.testing.assert_array_equal(out.read(), e.read())


@pytest.mark.parametrize("rop", list(w.ReduceOp))
@pytest.mark.parametrize("variant,R,Cc", [(w.GemvVariant.Gemv, 1024, 1024), (w.GemvVariant.GemvTr, 1024, 1024), (w.GemvVariant.Gemv, 200, 3000),
                                          (w.GemvVariant.GemvTr, 5000, 24), (w.GemvVariant.Gemv, 70001, 36), (w.GemvVariant.GemvTr, 96, 70001),
                                          (w.GemvVariant.Gemv, 5, 7), (w.GemvVariant.Gemv, 300001, 8)])
def test_gemv_fused_reduce_equals_the_two_dispatch_chain(gpu, shapes, rop, variant, R, Cc):
    """wgb_gemv_reduce == Gemv::dispatch into `out`, then Reduce::dispatch(out, result) (gemv.rs:64-137, reduce.rs:100-113), bit for
    bit — every reduce op, both variants, shapes with and without the split of the reduction axis, an output vector large enough
    for a multi-CTA reduce tree, and one beyond what the fused tail emulates (falls back to the second launch)."""
    dev = gpu.device()
    tr = variant == w.GemvVariant.GemvTr
    m = upload(gpu, O.uniform(SEED_A, R, Cc) + np.float32(0.5), (R, Cc))
    nv, nout = (R, Cc) if tr else (Cc, R)
    v = upload(gpu, O.uniform(SEED_V, nv) * np.float32(2.0 / nv if rop == w.ReduceOp.Prod else 1.0), (nv,))   # Prod: factors near 1
    out = w.TensorBuilder.vector(nout, STORAGE).build(dev)
    r_chain, r_fused = w.TensorBuilder.scalar(STORAGE).build(dev), w.TensorBuilder.scalar(STORAGE).build(dev)
    gemv, red = w.Gemv.from_device(dev), w.Reduce.new(dev, rop)
    n0 = dev.launch_count()
    run_pass(gpu, lambda p: (gemv.dispatch_generic(dev, shapes, p, out, m, v, variant), red.dispatch(dev, shapes, p, out, r_chain)))
    n1 = dev.launch_count()
    run_pass(gpu, lambda p: gemv.dispatch_reduce(dev, shapes, p, r_fused, m, v, rop, variant))
    n2 = dev.launch_count()
    a, b = r_chain.read(), r_fused.read()
    assert a.tobytes() == b.tobytes(), (a, b)
    assert rop == w.ReduceOp.Prod or np.isfinite(a[0])
    if nout <= 200000:
        assert n2 - n1 == 1 and n1 - n0 == 2      # one launch instead of two


@pytest.mark.parametrize("rop", list(w.ReduceOp))
@pytest.mark.parametrize("axis", [1, 2])
@pytest.mark.parametrize("M,N,K,tr,dtype,mode", [(256, 384, 192, False, "f32", None), (520, 200, 136, True, "f32", None),
                                                 (1024, 1160, 264, False, "bf16", None), (100, 60, 52, False, "f32", w.F32Mode.Simt),
                                                 (264, 136, 72, True, "f32", w.F32Mode.Tf32)])
def test_gemm_fused_reduce_matches_the_two_dispatch_chain(gpu, shapes, rop, axis, M, N, K, tr, dtype, mode):
    """wgb_gemm_reduce vs the chain Gemm::dispatch, then Reduce per column of the product (wgb_reduce_columns; for axis 2, per row):
    the product is never stored, results agree to f32 rounding (1e-5 relative; the fused form folds 32-row partials in index order,
    the chain reduces each whole column with its own tree) and a second run gives the same bits (deterministic)."""
    dev = gpu.device()
    ar, ac = (K, M) if tr else (M, K)
    scale = np.float32(2.0 / K) if rop == w.ReduceOp.Prod else np.float32(1.0)     # Prod: factors near 1
    A, B = O.uniform(SEED_A, ar, ac) * scale, O.uniform(SEED_B, K, N) + (np.float32(0.5) if rop == w.ReduceOp.Prod else np.float32(0.0))
    if dtype == "bf16":
        A, B = O.to_bf16_rne(A), O.to_bf16_rne(B)
        ta, tb = upload(gpu, O.bf16_bits(A), (ar, ac), "bf16"), upload(gpu, O.bf16_bits(B), (K, N), "bf16")
    else:
        ta, tb = upload(gpu, A, (ar, ac)), upload(gpu, B, (K, N))
    var = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
    gemm = w.Gemm.from_device(dev)
    n_out = N if axis == 1 else M
    r1, r2 = w.TensorBuilder.vector(n_out, STORAGE).build(dev), w.TensorBuilder.vector(n_out, STORAGE).build(dev)
    path = []
    run_pass(gpu, lambda p: (gemm.dispatch_reduce(dev, shapes, p, r1, ta, tb, rop, axis, var, f32_mode=mode), path.append(p.last_gemm_path())))
    run_pass(gpu, lambda p: gemm.dispatch_reduce(dev, shapes, p, r2, ta, tb, rop, axis, var, f32_mode=mode))
    got = r1.read()
    assert got.tobytes() == r2.read().tobytes()
    assert path[0] == (1 if mode == w.F32Mode.Simt else 2 if dtype == "bf16" else 3 if mode == w.F32Mode.Tf32 else 4)
    # the chain: the product stored (f32), then reduced along the axis in float64 for the expectation
    c = w.TensorBuilder.matrix(M, N, STORAGE).build(dev)
    run_pass(gpu, lambda p: gemm.dispatch_generic(dev, shapes, p, c, ta, tb, var, f32_mode=mode))
    C = cm(c.read(), M, N).astype(np.float64)
    red = {w.ReduceOp.Min: np.min, w.ReduceOp.Max: np.max, w.ReduceOp.Sum: np.sum, w.ReduceOp.Prod: np.prod,
           w.ReduceOp.SqNorm: lambda x, axis: np.sum(x * x, axis=axis)}[rop]
    want = red(C, axis=0 if axis == 1 else 1)
    if rop == w.ReduceOp.Prod:      # a product of ~1000 factors may leave the f32 range: compare where it stays inside
        ok = np.isfinite(want) & (np.abs(want) > 1e-30) & (np.abs(want) < 1e30)
        assert ok.sum() > 0 and rel_err(got[ok], want[ok]) < 1e-3
    else:
        assert rel_err(got, want) < 1e-5
    if axis == 1 and rop != w.ReduceOp.Prod:
        # and against the library's own single-launch column reduce of the stored product
        rc = w.TensorBuilder.vector(N, STORAGE).build(dev)
        run_pass(gpu, lambda p: w.Reduce.new(dev, rop).dispatch_columns(dev, shapes, p, c, rc))
        assert rel_err(got, rc.read().astype(np.float64)) < 1e-5


def test_graph_capture_replays_a_dispatch_chain(gpu, shapes):
    """wgb_graph_*: record gemm -> op_assign -> reduce once, replay it, same result as the eager sequence."""
    dev = gpu.device()
    n = 256
    A, B, bias = O.uniform(SEED_A, n, n), O.uniform(SEED_B, n, n), O.uniform(SEED_V, n * n)
    ta, tb, tbias = upload(gpu, A, (n, n)), upload(gpu, B, (n, n)), upload(gpu, bias, (n * n,))
    tc = w.TensorBuilder.matrix(n, n, STORAGE).build(dev)
    res = w.TensorBuilder.scalar(STORAGE).build(dev)
    gemm, add, rsum = w.Gemm.from_device(dev), w.OpAssign.new(dev, w.OpAssignVariant.Add), w.Reduce.new(dev, w.ReduceOp.Sum)

    def chain(p):
        gemm.dispatch(dev, shapes, p, tc, ta, tb)
        add.dispatch(dev, shapes, p, tc.reshape((n * n,)), tbias)
        rsum.dispatch(dev, shapes, p, tc.reshape((n * n,)), res)
    run_pass(gpu, chain)                       # eager (also warms the workspaces)
    eager_c, eager_r = tc.read(), res.read()[0]
    with dev.capture() as cap:
        run_pass(gpu, chain)
    run_pass(gpu, lambda p: w.OpAssign.new(dev, w.OpAssignVariant.Copy).dispatch(dev, shapes, p, tc.reshape((n * n,)), tbias))  # clobber
    n0 = dev.launch_count()
    cap.graph.launch()
    np.testing.assert_array_equal(tc.read(), eager_c)
    assert res.read()[0] == eager_r
    assert dev.launch_count() - n0 >= 3
    ref = (cm(A, n, n).astype(np.float64) @ cm(B, n, n).astype(np.float64)).T.reshape(-1) + bias
    assert rel_err(eager_c, ref) < F32_TOL


def test_timestamps_and_launch_counter(gpu, shapes):
    dev = gpu.device()
    ts = w.GpuTimestamps.new(dev, 8)
    n = 1 << 22
    a, b = (w.TensorBuilder.vector(n, STORAGE).build(dev) for _ in range(2))
    opk = w.OpAssign.new(dev, w.OpAssignVariant.Copy)
    enc = dev.create_command_encoder()
    n0 = dev.launch_count()
    p = enc.compute_pass("timed", ts)
    for _ in range(4):
        opk.dispatch(dev, shapes, p, a, b)
    p.end()
    ts.resolve(enc)
    gpu.queue().submit(enc.finish())
    ms = ts.wait_for_results_ms(dev, gpu.queue())
    assert len(ms) == 2 and ms[1] > 0.0
    assert dev.launch_count() - n0 == 4


def test_context_outlives_its_handle_while_children_exist():
    """wgpu handles are reference counted (a Buffer keeps its Device alive); so are these: destroying the context first leaves
    its buffers / events usable and destroyable, and host bindings whose finalisers run in arbitrary order cannot crash."""
    import subprocess
    import sys
    L = w.lib()
    ctx, buf, ev = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    assert L.wgb_ctx_create(0, ctypes.byref(ctx)) == 0
    data = np.arange(1024, dtype=np.float32)
    assert L.wgb_buffer_create_init(ctx, data.ctypes.data_as(ctypes.c_void_p), data.nbytes, STORAGE, ctypes.byref(buf)) == 0
    assert L.wgb_event_create(ctx, ctypes.byref(ev)) == 0
    assert L.wgb_ctx_destroy(ctx) == 0                       # the caller's reference only
    back = np.zeros_like(data)
    assert L.wgb_buffer_read(ctx, buf, 0, back.ctypes.data_as(ctypes.c_void_p), back.nbytes) == 0
    np.testing.assert_array_equal(back, data)
    assert L.wgb_event_destroy(ev) == 0
    assert L.wgb_buffer_destroy(buf) == 0                    # last child: the context goes with it
    # interpreter exit with module-level tensors caught in a reference cycle (finalisers in arbitrary order)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "exit_probe.py"), "all"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "done all" in r.stdout, (r.returncode, r.stdout, r.stderr)
