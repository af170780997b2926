"""GPU parity tests (-m gpu) for the batched small-matrix factorizations (first half): wgb_geometry_batch
through the reference-shaped host mirror (WgCholesky2.., WgLU2.., WgQR2.., WgSymmetricEigen2.., WgSvd2/3, WgInv) against
oracle/geometry_oracle.c.

Both sides evaluate the WGSL's operation sequence without FMA contraction (nvcc -fmad=false / gcc -ffp-contract=off), with
IEEE division and square root, so the comparison is BIT-EXACT on every output word — except svd2, which goes through
sin / cos / atan (different libm implementations): 4e-6 absolute there, stated in the test.  The reference's own pass criteria
(reconstruction at 1e-4 etc., tests/test_geometry_oracle.py) are re-checked on the GPU output as well."""
import numpy as np
import pytest

import wgmath_b200 as w
from oracle import oracle as O
from tests.helpers import STORAGE, run_pass
from tests.test_geometry_oracle import LEN, lu_fields, lu_reconstruct_error, random_mats, relative_eq, sdp, split
from wgmath_b200 import geometry as G

pytestmark = pytest.mark.gpu

CASES = [("cholesky", d) for d in (2, 3, 4)] + [("lu", d) for d in (2, 3, 4)] + [("qr", d) for d in (2, 3, 4)] + \
        [("eig", d) for d in (2, 3, 4)] + [("svd", d) for d in (2, 3)] + [("inv", d) for d in (2, 3, 4)]
SHADER = {("cholesky", 2): w.WgCholesky2, ("cholesky", 3): w.WgCholesky3, ("cholesky", 4): w.WgCholesky4,
          ("lu", 2): w.WgLU2, ("lu", 3): w.WgLU3, ("lu", 4): w.WgLU4, ("qr", 2): w.WgQR2, ("qr", 3): w.WgQR3, ("qr", 4): w.WgQR4,
          ("eig", 2): w.WgSymmetricEigen2, ("eig", 3): w.WgSymmetricEigen3, ("eig", 4): w.WgSymmetricEigen4,
          ("svd", 2): w.WgSvd2, ("svd", 3): w.WgSvd3}
ORC_OP = {"cholesky": O.GEOM_CHOLESKY, "lu": O.GEOM_LU, "qr": O.GEOM_QR, "eig": O.GEOM_EIG, "svd": O.GEOM_SVD, "inv": O.GEOM_INV}
OUT_NAME = {"cholesky": "mat", "inv": "mat", "lu": "lu", "qr": "qr", "eig": "eig", "svd": "svd"}


def inputs_for(op, dim, seed, n=LEN):
    a = random_mats(dim, seed, n)
    if op in ("cholesky", "lu", "eig"):
        return sdp(a)                                   # what the reference's tests feed these (cholesky.rs:98-102 ...)
    if op == "inv":
        return a + 2.0 * np.eye(dim, dtype=np.float32)
    return a


def gpu_batch(gpu, op, dim, packed_words, n=None, in_first=0, out_first=0, out_len=None, prefill=None):
    """Upload [m, in_words] float32, dispatch, read back [out_len, out_words] float32."""
    dev = gpu.device()
    m = packed_words.shape[0]
    ow = O.geom_out_words(ORC_OP[op], dim)
    out_len = m if out_len is None else out_len
    tin = w.TensorBuilder.vector(m, STORAGE).build_init(dev, packed_words.view(G.Matrix[dim]).reshape(-1), f"mat{dim}")
    out_dt = f"{OUT_NAME[op]}{dim}"
    init = np.zeros((out_len, ow), np.float32) if prefill is None else prefill
    tout = w.TensorBuilder.vector(out_len, STORAGE).build_init(dev, init.view(w.tensor._DT[out_dt][0]).reshape(-1), out_dt)
    n = m - in_first if n is None else n
    src = tin if in_first == 0 and n == m else tin.rows(in_first, n)
    dst = tout if out_first == 0 else tout.rows(out_first, out_len - out_first)
    if op == "inv":
        sh = w.WgInv.from_device(dev)
        run_pass(gpu, lambda p: sh.dispatch(dev, p, dim, src, dst, n))
    else:
        sh = SHADER[(op, dim)].from_device(dev)
        run_pass(gpu, lambda p: sh.dispatch(dev, p, src, dst, n))
    return tout.read().view(np.float32).reshape(out_len, ow)


def assert_matches_oracle(op, dim, got, ref):
    if op == "svd" and dim == 2:
        # sinf / cosf / atanf: CUDA's and glibc's differ by an ulp or two; everything else in svd2 is exact arithmetic on them
        assert np.abs(got - ref).max() <= 4e-6, np.abs(got - ref).max()
        return
    g, r = got.view(np.uint32), ref.view(np.uint32)
    both_nan = np.isnan(got) & np.isnan(ref)            # NaN payloads are not part of the contract (non-SDP Cholesky inputs)
    if not np.array_equal(np.where(both_nan, 0, g), np.where(both_nan, 0, r)):
        bad = np.argwhere((g != r) & ~both_nan)
        i, k = bad[0]
        raise AssertionError(f"{op}{dim}: {len(bad)} words differ (first: element {i} word {k}: {got[i, k]!r} vs {ref[i, k]!r}); "
                             f"max abs diff {np.nanmax(np.abs(got - ref))}")


@pytest.mark.parametrize("op,dim", CASES)
def test_reference_replay_matches_oracle(gpu, op, dim):
    """The reference's test batch (LEN = 345 random matrices) on the GPU: identical to the oracle, and the reference's own
    pass criterion holds for the GPU result."""
    m = inputs_for(op, dim, 1000 + 10 * dim + len(op))
    packed = O.geom_pack(m)
    got = gpu_batch(gpu, op, dim, packed)
    assert_matches_oracle(op, dim, got, O.geom_batch(ORC_OP[op], dim, packed))
    if op == "qr":
        f = split(got, dim, "qr")
        assert np.abs(f["q"].astype(np.float64) @ f["r"].astype(np.float64) - m).max() < 2e-5
    elif op == "eig":
        f = split(got, dim, "eig")
        v, wv = f["vectors"].astype(np.float64), f["values"].astype(np.float64)
        fails = int((~relative_eq(m, np.einsum("nij,nj,nkj->nik", v, wv, v), 1e-4)).sum())
        assert fails <= (0 if dim == 2 else LEN * 2 // 100)           # eig3.rs:117-123
    elif op == "svd":
        f = split(got, dim, "svd")
        rec = np.einsum("nij,nj,njk->nik", f["u"].astype(np.float64), f["s"].astype(np.float64), f["vt"].astype(np.float64))
        assert relative_eq(m, rec, 1e-4).all()                          # svd2.rs:105, svd3.rs:109
    elif op == "lu":
        lu, ia, ib, ln = lu_fields(got, dim)
        assert lu_reconstruct_error(m, lu, ia, ib, ln) < 1e-5
    elif op == "inv":
        inv = split(got, dim, "inv")["m"].astype(np.float64)
        assert np.abs(inv @ m - np.eye(dim)).max() < 1e-5


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 129, 4097])
@pytest.mark.parametrize("op,dim", [("cholesky", 3), ("lu", 2), ("qr", 4), ("eig", 4), ("svd", 3), ("inv", 4)])
def test_ragged_batch_sizes(gpu, op, dim, n):
    """Partial warp tiles and partial blocks."""
    packed = O.geom_pack(inputs_for(op, dim, 77 + n, n))
    assert_matches_oracle(op, dim, gpu_batch(gpu, op, dim, packed), O.geom_batch(ORC_OP[op], dim, packed))


@pytest.mark.parametrize("op,dim", [("lu", 3), ("qr", 2), ("eig", 3), ("svd", 2)])
def test_sub_range_leaves_the_rest_untouched(gpu, op, dim):
    """GpuVector::rows views on both sides (odd element offsets: 40- and 80-byte structs are only 8-byte aligned)."""
    m, first_in, n, first_out, out_len = 300, 37, 201, 5, 260
    packed = O.geom_pack(inputs_for(op, dim, 4242, m))
    ow = O.geom_out_words(ORC_OP[op], dim)
    sentinel = np.full((out_len, ow), -777.25, np.float32)
    got = gpu_batch(gpu, op, dim, packed, n=n, in_first=first_in, out_first=first_out, out_len=out_len, prefill=sentinel.copy())
    ref = sentinel.copy()
    ref[first_out:first_out + n] = O.geom_batch(ORC_OP[op], dim, packed[first_in:first_in + n])
    assert_matches_oracle(op, dim, got, ref)


def test_special_inputs(gpu):
    """Zero matrices, identity, exactly singular columns (LU `continue`, lu.wgsl:55-58), diagonal inputs (eig2 / eig3 early
    outs), negative entries; NaN-producing inputs must terminate and match the oracle's NaN pattern."""
    for dim in (2, 3, 4):
        eye = np.eye(dim, dtype=np.float32)
        mats = np.stack([np.zeros((dim, dim), np.float32), eye, -eye, np.diag(np.arange(1, dim + 1)).astype(np.float32),
                         np.ones((dim, dim), np.float32), np.triu(np.ones((dim, dim), np.float32)),
                         np.fliplr(eye).copy(), (random_mats(dim, 5, 1)[0] - 0.5)])
        mats[4][:, 0] = 0.0
        packed = O.geom_pack(mats)
        for op in ("cholesky", "lu", "qr", "eig", "svd", "inv"):
            if op == "svd" and dim == 4:
                continue
            got, ref = gpu_batch(gpu, op, dim, packed), O.geom_batch(ORC_OP[op], dim, packed)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (op, dim)
            if op == "svd" and dim == 2:
                assert np.nanmax(np.abs(got - ref)) <= 4e-6
            else:
                fin = ~np.isnan(ref)
                assert np.array_equal(got[fin].view(np.uint32), ref[fin].view(np.uint32)), (op, dim)


def test_in_place_cholesky_and_inverse(gpu):
    dev = gpu.device()
    for op, cls in (("cholesky", w.WgCholesky4), ("inv", None)):
        packed = O.geom_pack(inputs_for(op, 4, 99, 1000))
        t = w.TensorBuilder.vector(1000, STORAGE).build_init(dev, packed.view(G.Matrix[4]).reshape(-1), "mat4")
        if cls is None:
            sh = w.WgInv.from_device(dev)
            run_pass(gpu, lambda p: sh.dispatch(dev, p, 4, t, t))
        else:
            sh = cls.from_device(dev)
            run_pass(gpu, lambda p: sh.dispatch(dev, p, t, t))
        assert_matches_oracle(op, 4, t.read().view(np.float32).reshape(1000, 16), O.geom_batch(ORC_OP[op], 4, packed))


def test_error_behaviour(gpu):
    dev = gpu.device()
    t2 = w.TensorBuilder.vector(64, STORAGE).build(dev, "mat2")
    t3 = w.TensorBuilder.vector(64, STORAGE).build(dev, "mat3")
    q2 = w.TensorBuilder.vector(32, STORAGE).build(dev, "qr2")
    with pytest.raises(TypeError):                                       # element type mismatch (a compile error in Rust)
        run_pass(gpu, lambda p: w.WgQR2.from_device(dev).dispatch(dev, p, t3, q2))
    with pytest.raises(w.WgbError) as e:                                 # more inputs than the output can hold
        run_pass(gpu, lambda p: w.WgQR2.from_device(dev).dispatch(dev, p, t2, q2))
    assert "exceed the buffer" in str(e.value)
    with pytest.raises(w.WgbError):                                      # shifted in-place ranges race between warps
        run_pass(gpu, lambda p: w.WgCholesky2.from_device(dev).dispatch(dev, p, t2.rows(0, 32), t2.rows(16, 32)))
    enc = dev.create_command_encoder()
    p = enc.compute_pass("t", None)
    rc = w.lib().wgb_geometry_batch(p._h, G.GEOM_SVD, 4, t2.buffer()._h, 0, t2.buffer()._h, 0, 1)   # no 4x4 SVD in the reference
    p.end()
    assert rc != 0
    # empty dispatch: nothing queued, no error (kernel.rs:144)
    before = dev.launch_count()
    run_pass(gpu, lambda p: w.WgQR2.from_device(dev).dispatch(dev, p, t2, q2, 0))
    assert dev.launch_count() == before


@pytest.mark.parametrize("op,dim", [("cholesky", 4), ("lu", 4), ("qr", 3), ("eig", 4), ("eig", 3), ("svd", 3), ("inv", 3)])
def test_large_batch_matches_oracle(gpu, op, dim):
    """2^20 matrices: many resident waves of the grid-stride loop; full comparison with the (OpenMP) oracle."""
    n = 1 << 20
    packed = O.geom_pack(inputs_for(op, dim, 31337 + dim, n))
    got = gpu_batch(gpu, op, dim, packed)
    assert_matches_oracle(op, dim, got, O.geom_batch(ORC_OP[op], dim, packed))
