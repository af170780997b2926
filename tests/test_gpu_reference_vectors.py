"""GPU parity tests (-m gpu) against golden vectors computed from the reference's own shader source.

tests/golden/ref_wgsl_*.npz: the unmodified WGSL of the reference, executed by tests/golden/wgsl_interp.py with the reference's
dispatch grids (tests/golden/make_reference_vectors.py; tests/test_reference_vectors.py pins the oracle to the same files bit for
bit).  Here the CUDA path, called through the reference-shaped host API over the C ABI, runs the same views over the same
buffers:
  * gemm / gemv: every element of the view within 1e-5 relative (BASELINE.json north_star; the device sums in a different
    order), every element outside the view untouched (the reference's shaders write whole 4-row blocks, i.e. also into the
    padding rows of a ragged view — that spill-over is not reproduced, nothing else may be touched);
  * op_assign: bit-exact, whole buffer;  reduce: 1e-5 relative (min / max bit-exact);
  * factorizations: bit-exact words (svd2 through sin / cos / atan: 4e-6)."""
import os
import sys

import numpy as np
import pytest

import wgmath_b200 as w
from tests.helpers import STORAGE, run_pass
from tests.test_gpu_geometry import assert_matches_oracle, gpu_batch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
import reference_cases as C  # noqa: E402

F32_TOL = 1e-5


@pytest.fixture(scope="module")
def linalg_vectors():
    return np.load(os.path.join(GOLD, "ref_wgsl_linalg.npz"))


@pytest.fixture(scope="module")
def geometry_vectors():
    return np.load(os.path.join(GOLD, "ref_wgsl_geometry.npz"))


def view(shape, tensor):
    r, c, t, stride, stride_mat, off = shape
    return w.GpuTensorView(w.ViewShape((r, c, t), stride, stride_mat, off), tensor.buffer(), "f32", 3)


def view_mask(shape, length):
    """boolean mask of the buffer elements a (column-major) view addresses"""
    r, c, t, stride, stride_mat, off = shape
    m = np.zeros(length, bool)
    for k in range(t):
        for j in range(c):
            base = off + k * stride_mat + j * stride
            m[base:base + r] = True
    return m


def buf(gpu, a):
    return w.TensorBuilder.vector(len(a), STORAGE).build_init(gpu.device(), a)


def check_view(got, want, initial, shape):
    inside = view_mask(shape, len(got))
    rel = np.abs(got[inside].astype(np.float64) - want[inside]) / np.abs(want[inside].astype(np.float64))
    assert rel.max() < F32_TOL, rel.max()
    np.testing.assert_array_equal(got[~inside].view(np.uint32), initial[~inside].view(np.uint32))


@pytest.mark.parametrize("case", C.gemm_cases(), ids=lambda c: c["name"])
def test_gemm_matches_the_reference_shaders(gpu, shapes, linalg_vectors, case):
    b = C.inputs(case)
    out, m1, m2 = buf(gpu, b["out"]), buf(gpu, b["m1"]), buf(gpu, b["m2"])
    gemm = w.Gemm.from_device(gpu.device())
    variant = w.GemmVariant(C.GEMM_VARIANTS[case["variant"]])
    run_pass(gpu, lambda p: gemm.dispatch_generic(gpu.device(), shapes, p, view(case["so"], out), view(case["s1"], m1), view(case["s2"], m2), variant))
    check_view(out.read(), linalg_vectors[case["name"]], b["out"], case["so"])
    np.testing.assert_array_equal(m1.read(), b["m1"])
    np.testing.assert_array_equal(m2.read(), b["m2"])


@pytest.mark.parametrize("case", C.gemv_cases(), ids=lambda c: c["name"])
def test_gemv_matches_the_reference_shaders(gpu, shapes, linalg_vectors, case):
    b = C.inputs(case)
    out, m, v = buf(gpu, b["out"]), buf(gpu, b["m"]), buf(gpu, b["v"])
    gemv = w.Gemv.from_device(gpu.device())
    variant = w.GemvVariant(C.GEMV_VARIANTS[case["variant"]])
    run_pass(gpu, lambda p: gemv.dispatch_generic(gpu.device(), shapes, p, view(case["so"], out), view(case["sm"], m), view(case["sv"], v), variant))
    check_view(out.read(), linalg_vectors[case["name"]], b["out"], case["so"])


@pytest.mark.parametrize("case", C.op_assign_cases(), ids=lambda c: c["name"])
def test_op_assign_matches_the_reference_shader_bit_for_bit(gpu, shapes, linalg_vectors, case):
    b = C.inputs(case)
    a, bb = buf(gpu, b["a"]), buf(gpu, b["b"])
    op = w.OpAssign.new(gpu.device(), w.OpAssignVariant(C.OP_ASSIGN[case["op"]]))
    va = w.GpuTensorView(w.ViewShape((case["sa"][0], 1, 1), case["sa"][3], case["sa"][4], case["sa"][5]), a.buffer(), "f32", 1)
    vb = w.GpuTensorView(w.ViewShape((case["sb"][0], 1, 1), case["sb"][3], case["sb"][4], case["sb"][5]), bb.buffer(), "f32", 1)
    run_pass(gpu, lambda p: op.dispatch(gpu.device(), shapes, p, va, vb))
    np.testing.assert_array_equal(a.read().view(np.uint32), linalg_vectors[case["name"]].view(np.uint32))


@pytest.mark.parametrize("case", C.reduce_cases(), ids=lambda c: c["name"])
def test_reduce_matches_the_reference_shader(gpu, shapes, linalg_vectors, case):
    b = C.inputs(case)
    x = buf(gpu, b["x"])
    res = w.TensorBuilder.scalar(STORAGE).build_init(gpu.device(), b["out"])
    red = w.Reduce.new(gpu.device(), w.ReduceOp(C.REDUCE[case["op"]]))
    n, off = case["s"][0], case["s"][5]
    vx = w.GpuTensorView(w.ViewShape((n, 1, 1), case["s"][3], case["s"][4], off), x.buffer(), "f32", 1)
    run_pass(gpu, lambda p: red.dispatch(gpu.device(), shapes, p, vx, res))
    got, want = float(res.read()[0]), float(linalg_vectors[case["name"]][0])
    if case["op"] in ("min", "max") or n <= 1:
        assert got == want                                        # order-independent: exact, the +-3.4e38 initial values included
    else:
        assert abs(got - want) <= F32_TOL * abs(want)


@pytest.mark.parametrize("op,dim", C.geometry_cases(), ids=lambda x: str(x))
def test_factorizations_match_the_reference_shaders(gpu, geometry_vectors, op, dim):
    from oracle import oracle as O
    packed = O.geom_pack(C.geometry_inputs(op, dim))
    got = gpu_batch(gpu, op, dim, packed)
    assert_matches_oracle(op, dim, got, geometry_vectors[f"{op}{dim}"])


@pytest.fixture(scope="module")
def scan_sort_vectors():
    return np.load(os.path.join(GOLD, "ref_wgsl_scan_sort.npz"))


@pytest.mark.parametrize("case", C.scan_cases(), ids=lambda c: c["name"])
def test_prefix_sum_matches_the_reference_shaders(gpu, scan_sort_vectors, case):
    v = C.scan_input(case)
    t = w.TensorBuilder.vector(v.size, STORAGE).build_init(gpu.device(), v, "u32")
    ps = w.WgPrefixSum.from_device(gpu.device())
    ws = w.PrefixSumWorkspace.with_capacity(gpu.device(), v.size)
    run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, ws, t))
    np.testing.assert_array_equal(t.read(), scan_sort_vectors["scan/" + case["name"]])


@pytest.mark.parametrize("case", C.sort_cases(), ids=lambda c: c["name"])
def test_radix_sort_matches_the_reference_shaders(gpu, scan_sort_vectors, case):
    from tests.test_gpu_scan_sort import sort_on_gpu
    keys, values = C.sort_input(case)
    gk, gv = sort_on_gpu(gpu, keys, values, case["n_sort"], case["bits"])
    n = case["n_sort"]
    np.testing.assert_array_equal(gk[:n], scan_sort_vectors[f"sort/{case['name']}/keys"][:n])
    np.testing.assert_array_equal(gv[:n], scan_sort_vectors[f"sort/{case['name']}/values"][:n])
