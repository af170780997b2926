"""Host-side mirror of the factorization libraries of wgebra::geometry (first half):

  WgCholesky2 / 3 / 4          wgebra/src/geometry/cholesky.rs:21-38  + cholesky.wgsl
  WgLU2 / 3 / 4, GpuLU*        wgebra/src/geometry/lu.rs:25-79        + lu.wgsl
  WgQR2 / 3 / 4, GpuQR*        wgebra/src/geometry/qr2.rs:9-27 (qr3.rs, qr4.rs)       + qr2.wgsl ...
  WgSymmetricEigen2 / 3 / 4    wgebra/src/geometry/eig2.rs:10-26 (eig3.rs, eig4.rs)   + eig2.wgsl ...
  WgSvd2 / WgSvd3, GpuSvd*     wgebra/src/geometry/svd2.rs:9-23, svd3.rs:10-28        + svd2.wgsl, svd3.wgsl
  WgInv                        wgebra/src/geometry/inv.rs:3-8         + inv.wgsl

In the reference each of these is a `Shader` that contributes WGSL functions to other kernels; the only kernels it builds from
them are the per-module test kernels `out[i] = f(in[i])` (e.g. cholesky.rs:53-63).  Here the function library is
csrc/geometry.cuh and `dispatch` is that batched kernel, through the C ABI (wgb_geometry_batch) — the call the reference's
tests make as `KernelDispatch::new(device, &mut pass, &pipeline).bind0([inputs.buffer(), result.buffer()]).dispatch(len)`.

Element types are numpy structured dtypes with WGSL's storage layout (what `GpuVector::init` / `encase` upload in the
reference): a matrix field is indexed `m[column][row]`; 3x3 matrices have 4-float columns (Matrix4x3 in the reference's tests,
cholesky.rs:151-153).  There is no CPU fallback: without the CUDA library these classes raise."""
from __future__ import annotations

import numpy as np

from ._lib import check, lib
from .tensor import GpuTensor, GpuTensorView, register_dtype

GEOM_CHOLESKY, GEOM_LU, GEOM_QR, GEOM_SYMMETRIC_EIGEN, GEOM_SVD, GEOM_INV = range(6)

_f4, _u4 = np.dtype("<f4"), np.dtype("<u4")


def _cs(dim: int) -> int:
    return 2 if dim == 2 else 4


def _mat(dim: int):
    return (_f4, (dim, _cs(dim)))


# Matrix2<f32> / Matrix4x3<f32> / Matrix4<f32> as uploaded by the reference's tests
Matrix = {d: np.dtype([("m", *_mat(d))]) for d in (2, 3, 4)}
# lu.rs:25-56 gpu_output_types!: {lu, p: {ia, ib, len}} — flattened, with WGSL's vec3<u32> padding
GpuLU = {
    2: np.dtype([("lu", *_mat(2)), ("ia", _u4, (2,)), ("ib", _u4, (2,)), ("len", _u4), ("_pad", _u4)]),
    3: np.dtype([("lu", *_mat(3)), ("ia", _u4, (4,)), ("ib", _u4, (3,)), ("len", _u4)]),
    4: np.dtype([("lu", *_mat(4)), ("ia", _u4, (4,)), ("ib", _u4, (4,)), ("len", _u4), ("_pad", _u4, (3,))]),
}
GpuQR = {d: np.dtype([("q", *_mat(d)), ("r", *_mat(d))]) for d in (2, 3, 4)}                       # qr3.rs:15-20
GpuSymmetricEigen = {d: np.dtype([("eigenvectors", *_mat(d)), ("eigenvalues", _f4, (_cs(d),))])     # eig3.rs:16-21
                     for d in (2, 3, 4)}
GpuSvd = {d: np.dtype([("u", *_mat(d)), ("s", _f4, (_cs(d),)), ("vt", *_mat(d))]) for d in (2, 3)}  # svd2.rs:12-19, svd3.rs:15-22

for _d in (2, 3, 4):
    register_dtype(f"mat{_d}", Matrix[_d])
    register_dtype(f"lu{_d}", GpuLU[_d])
    register_dtype(f"qr{_d}", GpuQR[_d])
    register_dtype(f"eig{_d}", GpuSymmetricEigen[_d])
    if _d < 4:
        register_dtype(f"svd{_d}", GpuSvd[_d])


def pack(mats: np.ndarray) -> np.ndarray:
    """[n, dim, dim] matrices indexed [i][row][col] -> structured array of Matrix[dim] (column-major, padded)."""
    mats = np.asarray(mats, np.float32)
    n, dim, _ = mats.shape
    out = np.zeros(n, Matrix[dim])
    out["m"][:, :, :dim] = np.transpose(mats, (0, 2, 1))
    return out


def unpack(field: np.ndarray) -> np.ndarray:
    """A matrix field [n, dim, col_stride] (indexed [i][col][row]) -> [n, dim, dim] indexed [i][row][col]."""
    dim = field.shape[1]
    return np.transpose(field[:, :, :dim], (0, 2, 1))


class _GeometryShader:
    """Common part of the Wg* structs below."""
    OP = DIM = None
    IN = OUT = None                # registered dtype names
    FILE_PATH = SRC = None         # the reference's `Shader` consts name the WGSL source; here the CUDA source
    OUT_TYPE = None                # numpy dtype of one output element

    def __init__(self, device):
        self._device = device

    @classmethod
    def from_device(cls, device):
        lib()
        assert lib().wgb_geometry_in_bytes(cls.DIM) == Matrix[cls.DIM].itemsize
        assert lib().wgb_geometry_out_bytes(cls.OP, cls.DIM) == cls.OUT_TYPE.itemsize
        return cls(device)

    def dispatch(self, device, pass_, inputs, outputs, n: int | None = None) -> None:
        """outputs[i] = f(inputs[i]) for the first n elements (default: all of `inputs`).  `inputs` is a
        GpuVector<Matrix>, `outputs` a GpuVector of the result struct; vector views (GpuVector::rows) select a sub-range."""
        ib, i0, ilen, idt = _range(inputs)
        ob, o0, olen, odt = _range(outputs)
        if idt != self.IN or odt != self.OUT:
            raise TypeError(f"{type(self).__name__}.dispatch: expected GpuVector<{self.IN}> -> GpuVector<{self.OUT}>, "
                            f"got {idt} -> {odt}")
        n = ilen if n is None else int(n)
        check(lib().wgb_geometry_batch(pass_._h, self.OP, self.DIM, ib._h, i0, ob._h, o0, n))


def _range(x):
    if isinstance(x, GpuTensorView):
        return x.buffer(), x.view_shape.offset, x.view_shape.size[0], x.dtype
    assert isinstance(x, GpuTensor)
    return x.buffer(), 0, x.len(), x.dtype


def _make(name: str, op: int, dim: int, out_name: str, out_type: np.dtype, doc: str):
    cls = type(name, (_GeometryShader,), {"OP": op, "DIM": dim, "IN": f"mat{dim}", "OUT": out_name, "OUT_TYPE": out_type,
                                          "FILE_PATH": "wgmath_b200/csrc/geometry.cuh", "__doc__": doc})
    return cls


WgCholesky2, WgCholesky3, WgCholesky4 = (
    _make(f"WgCholesky{d}", GEOM_CHOLESKY, d, f"mat{d}", Matrix[d],
          f"cholesky.rs:21-38: Cholesky factor of a symmetric-definite-positive {d}x{d} matrix (lower triangle of the output).")
    for d in (2, 3, 4))
WgLU2, WgLU3, WgLU4 = (
    _make(f"WgLU{d}", GEOM_LU, d, f"lu{d}", GpuLU[d], f"lu.rs:65-79: LU with partial pivoting of a {d}x{d} matrix -> GpuLU{d}.")
    for d in (2, 3, 4))
WgQR2, WgQR3, WgQR4 = (
    _make(f"WgQR{d}", GEOM_QR, d, f"qr{d}", GpuQR[d], f"qr{d}.rs:22-27: Householder QR of a {d}x{d} matrix -> GpuQR{d}.")
    for d in (2, 3, 4))
WgSymmetricEigen2, WgSymmetricEigen3, WgSymmetricEigen4 = (
    _make(f"WgSymmetricEigen{d}", GEOM_SYMMETRIC_EIGEN, d, f"eig{d}", GpuSymmetricEigen[d],
          f"eig{d}.rs:23-27: eigendecomposition of a symmetric {d}x{d} matrix -> GpuSymmetricEigen{d}.")
    for d in (2, 3, 4))
WgSvd2, WgSvd3 = (
    _make(f"WgSvd{d}", GEOM_SVD, d, f"svd{d}", GpuSvd[d], f"svd{d}.rs: SVD of a {d}x{d} matrix -> GpuSvd{d}.")
    for d in (2, 3))


class WgInv:
    """inv.rs:3-8 (one shader with inv2 / inv3 / inv4, inv.wgsl:8-88): closed-form inverse; `dim` picks the function."""

    def __init__(self, device):
        self._device = device

    @staticmethod
    def from_device(device) -> "WgInv":
        lib()
        return WgInv(device)

    def dispatch(self, device, pass_, dim: int, inputs, outputs, n: int | None = None) -> None:
        ib, i0, ilen, idt = _range(inputs)
        ob, o0, olen, odt = _range(outputs)
        if idt != f"mat{dim}" or odt != f"mat{dim}":
            raise TypeError(f"WgInv.dispatch: expected GpuVector<mat{dim}> -> GpuVector<mat{dim}>, got {idt} -> {odt}")
        check(lib().wgb_geometry_batch(pass_._h, GEOM_INV, dim, ib._h, i0, ob._h, o0, ilen if n is None else int(n)))
