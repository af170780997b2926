"""GPU parity tests (-m gpu) for RowMajor / ColumnMajor views (tensor.rs:17-39, shape.wgsl:49-57): every combination of
orderings of out / m1 / m2, with and without the transposed variant, on each GEMM kernel family and on GEMV, against the
oracle's scalar restatement of the row-major addressing (oracle/wgsl_oracle.c orc_gemm_ord / orc_gemv_ord).

Tolerances as in test_gpu_parity.py: 1e-5 relative for f32 (3xTF32 / FFMA), 1e-2 for bf16 output."""
import itertools

import numpy as np
import pytest

import wgmath_b200 as w
from oracle import oracle as O
from tests.helpers import SEED_A, SEED_B, SEED_OUT, SEED_V, oshape, rel_err, run_pass, upload

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5
BF16_TOL = 1e-2
ORD = {0: w.ColumnMajor, 1: w.RowMajor}


def view_of(rows, cols, T, rm, pad, off, extra=8):
    """A [rows x cols x T] view with ordering `rm` inside a padded parent: (ViewShape, parent length)."""
    inner, outer = (cols, rows) if rm else (rows, cols)
    ld = inner + pad
    smat = ld * outer + extra
    return w.ViewShape((rows, cols, T), ld, smat, off), off + smat * T


def addressed(vs, rm):
    """Flat indices of every element of the view, as [T, rows, cols]."""
    r, c, t = vs.size
    i, j, k = np.meshgrid(np.arange(r), np.arange(c), np.arange(t), indexing="ij")
    idx = vs.offset + k * vs.stride_mat + (i * vs.stride + j if rm else i + j * vs.stride)
    return np.transpose(idx, (2, 0, 1))


def gemm_ord_case(gpu, shapes, M, N, K, tr, ro, r1, r2, T=1, dtype="f32", out_dtype="f32", mode=None, pad=(0, 0, 0), off=(0, 0, 0),
                  tol=F32_TOL, op=None, cfg_out=None, out_extra=12):
    ar, ac = (K, M) if tr else (M, K)
    s1, n1 = view_of(ar, ac, T, r1, pad[0], off[0])
    s2, n2 = view_of(K, N, T, r2, pad[1], off[1])
    so, no = view_of(M, N, T, ro, pad[2], off[2], extra=out_extra)
    A, B = O.uniform(SEED_A, n1), O.uniform(SEED_B, n2)
    if dtype == "bf16":
        A, B = O.to_bf16_rne(A), O.to_bf16_rne(B)
        ta, tb = upload(gpu, O.bf16_bits(A), (n1,), "bf16"), upload(gpu, O.bf16_bits(B), (n2,), "bf16")
    else:
        ta, tb = upload(gpu, A, (n1,)), upload(gpu, B, (n2,))
    sentinel = np.float32(-3.0)
    C0 = np.full(no, sentinel, np.float32)
    tc = upload(gpu, O.bf16_bits(C0) if out_dtype == "bf16" else C0, (no,), out_dtype)
    va = w.GpuTensorView(s1, ta.buffer(), dtype, 3, ORD[r1])
    vb = w.GpuTensorView(s2, tb.buffer(), dtype, 3, ORD[r2])
    vc = w.GpuTensorView(so, tc.buffer(), out_dtype, 3, ORD[ro])
    E = None
    if op is not None:   # fused element-wise operand, ordered like out, in its own padded parent
        se, ne = view_of(M, N, T, ro, pad[2] + 4, 4)
        E = O.uniform(SEED_OUT, ne) + np.float32(0.5)
        if out_dtype == "bf16":
            E = O.to_bf16_rne(E)
        te = upload(gpu, O.bf16_bits(E) if out_dtype == "bf16" else E, (ne,), out_dtype)
        ve = w.GpuTensorView(se, te.buffer(), out_dtype, 3, ORD[ro])
    gemm = w.Gemm.from_device(gpu.device())
    variant = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
    path = []

    def go(p):
        if op is None:
            gemm.dispatch_generic(gpu.device(), shapes, p, vc, va, vb, variant, f32_mode=mode)
        else:
            gemm.dispatch_op(gpu.device(), shapes, p, vc, va, vb, op, ve, variant=variant, f32_mode=mode)
        path.append(p.last_gemm_path())
        if cfg_out is not None:
            cfg_out.append(p.last_gemm_config())
    run_pass(gpu, go)
    got = tc.read()
    if out_dtype == "bf16":
        got = O.bf16_from_bits(got)
    ref = np.full(no, sentinel, np.float32)
    assert O.gemm_ord(int(variant), ref, oshape(so), ro, A, oshape(s1), r1, B, oshape(s2), r2) == O.ORC_OK
    idx = addressed(so, ro).reshape(-1)
    want = ref[idx].astype(np.float64)
    if op is not None:
        e = E[addressed(se, ro).reshape(-1)].astype(np.float64)
        want = {w.OpAssignVariant.Add: want + e, w.OpAssignVariant.Sub: want - e, w.OpAssignVariant.Mul: want * e,
                w.OpAssignVariant.Div: want / e}[op]
    err = rel_err(got[idx], want)
    assert err < tol, f"rel err {err:.3e} (path {path}, tr={tr} out/m1/m2 row-major = {ro}{r1}{r2})"
    mask = np.ones(no, bool)
    mask[idx] = False
    assert np.all(got[mask] == sentinel), "wrote outside the output view"
    return path[0]


COMBOS = list(itertools.product([False, True], [0, 1], [0, 1], [0, 1]))   # tr, out, m1, m2


@pytest.mark.parametrize("tr,ro,r1,r2", COMBOS)
def test_gemm_orderings_ffma(gpu, shapes, tr, ro, r1, r2):
    # ragged sizes, odd strides and offsets, batched: the FFMA kernel (path 1)
    assert gemm_ord_case(gpu, shapes, 131, 70, 45, tr, ro, r1, r2, T=2, pad=(3, 5, 1), off=(7, 9, 3)) == 1
    assert gemm_ord_case(gpu, shapes, 5, 3, 7, tr, ro, r1, r2, mode=w.F32Mode.Simt) == 1


@pytest.mark.parametrize("tr,ro,r1,r2", COMBOS)
def test_gemm_orderings_3xtf32(gpu, shapes, tr, ro, r1, r2):
    assert gemm_ord_case(gpu, shapes, 264, 200, 328, tr, ro, r1, r2, T=2, mode=w.F32Mode.X3Tf32, pad=(4, 8, 4), off=(4, 8, 12)) == 4


@pytest.mark.parametrize("tr,ro,r1,r2", COMBOS)
def test_gemm_orderings_single_tf32(gpu, shapes, tr, ro, r1, r2):
    assert gemm_ord_case(gpu, shapes, 256, 384, 512, tr, ro, r1, r2, mode=w.F32Mode.Tf32, tol=2e-3) == 3


@pytest.mark.parametrize("tr,ro,r1,r2", COMBOS)
def test_gemm_orderings_bf16_tcgen05(gpu, shapes, tr, ro, r1, r2):
    # in-place MN-major B operand (row-major m2 / swapped operands of a row-major out); f32 output shows the exact product
    assert gemm_ord_case(gpu, shapes, 384, 520, 192, tr, ro, r1, r2, dtype="bf16", tol=1e-4, pad=(8, 16, 4), off=(8, 16, 4)) == 2
    assert gemm_ord_case(gpu, shapes, 200, 136, 72, tr, ro, r1, r2, T=3, dtype="bf16", out_dtype="bf16", tol=BF16_TOL, pad=(8, 8, 8)) == 2


@pytest.mark.parametrize("ro,r1,r2", [(0, 0, 1), (1, 0, 0), (1, 1, 1), (0, 1, 1)])
def test_gemm_orderings_bf16_tail_strips_and_narrow_n(gpu, shapes, ro, r1, r2):
    # 80 tiles of 256 x 256 on 74 CTA pairs: the 6 tail tiles are cut into column strips (whole 128-byte atoms of an MN-major B)
    assert gemm_ord_case(gpu, shapes, 2048, 2560, 128, False, ro, r1, r2, dtype="bf16", tol=1e-4) == 2
    assert gemm_ord_case(gpu, shapes, 2560, 2048, 128, True, ro, r1, r2, dtype="bf16", tol=1e-4) == 2
    # N <= 128: BLOCK_N = 128, one atom per CTA
    assert gemm_ord_case(gpu, shapes, 512, 96, 256, False, ro, r1, r2, dtype="bf16", tol=1e-4) == 2


@pytest.mark.parametrize("op", [w.OpAssignVariant.Add, w.OpAssignVariant.Div])
@pytest.mark.parametrize("ro,r1,r2", [(1, 0, 0), (1, 1, 0), (0, 0, 1)])
def test_gemm_orderings_fused_op(gpu, shapes, op, ro, r1, r2):
    gemm_ord_case(gpu, shapes, 256, 384, 192, False, ro, r1, r2, op=op)
    gemm_ord_case(gpu, shapes, 100, 60, 52, True, ro, r1, r2, op=op, mode=w.F32Mode.Simt)
    gemm_ord_case(gpu, shapes, 256, 384, 192, False, ro, r1, r2, op=op, dtype="bf16", out_dtype="bf16", tol=BF16_TOL)


@pytest.mark.parametrize("tr", [False, True])
@pytest.mark.parametrize("R,Cc,ncol,T,pad,off", [(1024, 256, 1, 1, 0, 0), (130, 67, 3, 2, 3, 5), (64, 4096, 1, 1, 4, 8), (4100, 36, 2, 1, 0, 4)])
def test_gemv_row_major_matrix(gpu, shapes, tr, R, Cc, ncol, T, pad, off):
    sm, nm = view_of(R, Cc, T, 1, pad, off)
    klen, olen = (R, Cc) if tr else (Cc, R)
    sv, nv = view_of(klen, ncol, T, 0, pad, off)
    so, no = view_of(olen, ncol, T, 0, pad, off)
    Mb, Vb = O.uniform(SEED_A, nm), O.uniform(SEED_V, nv)
    sentinel = np.float32(-9.0)
    Ob = np.full(no, sentinel, np.float32)
    tm, tv, to = upload(gpu, Mb, (nm,)), upload(gpu, Vb, (nv,)), upload(gpu, Ob, (no,))
    gemv = w.Gemv.from_device(gpu.device())
    var = w.GemvVariant.GemvTr if tr else w.GemvVariant.Gemv
    run_pass(gpu, lambda p: gemv.dispatch_generic(gpu.device(), shapes, p, w.GpuTensorView(so, to.buffer(), "f32", 3),
                                                  w.GpuTensorView(sm, tm.buffer(), "f32", 3, w.RowMajor),
                                                  w.GpuTensorView(sv, tv.buffer(), "f32", 3), var))
    got = to.read()
    ref = Ob.copy()
    assert O.gemv_ord(int(var), ref, oshape(so), Mb, oshape(sm), 1, Vb, oshape(sv)) == O.ORC_OK
    idx = addressed(so, 0).reshape(-1)
    assert rel_err(got[idx], ref[idx]) < F32_TOL
    mask = np.ones(no, bool)
    mask[idx] = False
    assert np.all(got[mask] == sentinel)


def test_row_major_views_of_a_tensor(gpu, shapes):
    """as_view(RowMajor) / reshape default strides (tensor.rs:282-297, :525-529) and row-major sub-views."""
    r, c = 96, 160
    a = np.arange(r * c, dtype=np.float32).reshape(r, c) / 1000.0          # numpy C order == row-major storage
    b = O.uniform(SEED_B, c * 128)                                          # column-major [c x 128]
    ta = upload(gpu, a.reshape(-1), (r, c))
    tb = upload(gpu, b, (c, 128))
    out = upload(gpu, np.zeros(r * 128, np.float32), (r, 128))
    va = ta.as_view(w.RowMajor)
    assert va.shape().stride == c and va.ordering is w.RowMajor
    gemm = w.Gemm.from_device(gpu.device())
    run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, out, va, tb))
    ref = a.astype(np.float64) @ b.reshape(128, c).T.astype(np.float64)
    assert rel_err(out.read().reshape(128, r).T, ref) < F32_TOL
    # rows 32..64 x columns 16..80 of the row-major matrix
    sub = va.rows(32, 32).columns(16, 64)
    assert sub.shape().offset == 32 * c + 16 and sub.ordering is w.RowMajor
    out2 = upload(gpu, np.zeros(32 * 128, np.float32), (32, 128))
    run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, out2, sub, tb.rows(16, 64)))
    ref2 = a[32:64, 16:80].astype(np.float64) @ b.reshape(128, c).T[16:80].astype(np.float64)
    assert rel_err(out2.read().reshape(128, 32).T, ref2) < F32_TOL


def test_ordering_errors(gpu, shapes):
    gemm = w.Gemm.from_device(gpu.device())
    z = lambda r, c: upload(gpu, np.zeros(r * c, np.float32), (r, c))
    with pytest.raises(w.DimensionMismatch, match="Gemm: dimension mismatch"):
        run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, z(8, 8), z(8, 4).as_view(w.RowMajor), z(8, 8)))
    with pytest.raises(w.WgbError, match="reaches"):   # bounds are checked with the row-major extent
        bad = w.GpuTensorView(w.ViewShape((4, 16, 1), 16, 64, 1), z(8, 8).buffer(), "f32", 3, w.RowMajor)
        run_pass(gpu, lambda p: gemm.dispatch(gpu.device(), shapes, p, bad, z(4, 8), z(8, 16)))
