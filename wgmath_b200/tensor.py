"""GpuTensor / views / TensorBuilder — the host-side mirror of
/root/reference/crates/wgcore/src/tensor.rs (line numbers below refer to that file).

Element types: the reference is generic over `T: Pod`; here a tensor carries a dtype tag,
"f32" (numpy float32), "bf16" (numpy uint16 bit patterns), "u32" (numpy uint32: the GpuVector<u32> / GpuScalar<u32>
operands of the scan and sort primitives) or a registered struct type (register_dtype: the small-matrix element types of
wgebra::geometry)."""
from __future__ import annotations

import ctypes
from math import prod
from typing import Optional, Sequence

import numpy as np

from ._lib import check, lib
from .shapes import ViewShape


class BufferUsages:
    """wgpu::BufferUsages bit values (passed through to the C ABI unchanged)."""
    MAP_READ = 1 << 0
    MAP_WRITE = 1 << 1
    COPY_SRC = 1 << 2
    COPY_DST = 1 << 3
    UNIFORM = 1 << 6
    STORAGE = 1 << 7


_DT = {"f32": (np.float32, 4), "bf16": (np.uint16, 2), "u32": (np.uint32, 4)}


def register_dtype(name: str, dtype: np.dtype) -> str:
    """A `T: Pod` / `T: ShaderType` element type beyond the scalar ones: a numpy structured dtype laid out like the WGSL
    storage struct (geometry.py registers Matrix2 / Matrix4x3 / Matrix4 and the GpuLU* / GpuQR* / ... structs)."""
    dtype = np.dtype(dtype)
    _DT[name] = (dtype, dtype.itemsize)
    return name


class Buffer:
    """wgpu::Buffer."""

    def __init__(self, device, handle, nbytes: int):
        self.device, self._h, self.nbytes = device, handle, nbytes

    def size(self) -> int:
        return self.nbytes

    def device_ptr(self) -> int:
        p = ctypes.c_void_p()
        check(lib().wgb_buffer_device_ptr(self._h, ctypes.byref(p)))
        return p.value or 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().wgb_buffer_destroy(h)
            except Exception:
                pass


class TensorBuilder:
    """:65-187."""

    def __init__(self, shape: Sequence[int], usage: int):
        self.shape, self.usage, self._label = tuple(int(s) for s in shape), usage, None

    @staticmethod
    def scalar(usage: int) -> "TensorBuilder":                      # :74-76
        return TensorBuilder((), usage)

    @staticmethod
    def vector(dim: int, usage: int) -> "TensorBuilder":            # :81-83
        return TensorBuilder((dim,), usage)

    @staticmethod
    def matrix(nrows: int, ncols: int, usage: int) -> "TensorBuilder":  # :88-90
        return TensorBuilder((nrows, ncols), usage)

    @staticmethod
    def tensor(shape: Sequence[int], usage: int) -> "TensorBuilder":    # :95-101
        return TensorBuilder(shape, usage)

    def len(self) -> int:                                           # :104-106
        return prod(self.shape)

    def label(self, label: str) -> "TensorBuilder":
        self._label = label
        return self

    def build(self, device, dtype: str = "f32") -> "GpuTensor":     # :112-129
        nbytes = _DT[dtype][1] * self.len()
        h = ctypes.c_void_p()
        check(lib().wgb_buffer_create(device._h, nbytes, self.usage, ctypes.byref(h)))
        return GpuTensor(self.shape, Buffer(device, h, nbytes), dtype)

    def build_bytes(self, device, data: bytes, dtype: str = "f32") -> "GpuTensor":  # :149-161
        h = ctypes.c_void_p()
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data) if len(data) else None
        check(lib().wgb_buffer_create_init(device._h, buf, len(data), self.usage, ctypes.byref(h)))
        return GpuTensor(self.shape, Buffer(device, h, len(data)), dtype)

    def build_init(self, device, data, dtype: Optional[str] = None) -> "GpuTensor":  # :175-186
        arr = np.asarray(data)
        if dtype is None:
            dtype = "bf16" if arr.dtype == np.uint16 else "u32" if arr.dtype == np.uint32 else "f32"
        arr = np.ascontiguousarray(arr.reshape(-1), dtype=_DT[dtype][0])
        n = self.len()
        assert arr.size >= n, (f"Incorrect number of elements provided for initializing Tensor."
                               f"Expected at least {n}, found {arr.size}")   # :176-182
        arr = arr[:n]
        h = ctypes.c_void_p()
        check(lib().wgb_buffer_create_init(device._h, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes, self.usage,
                                           ctypes.byref(h)))
        return GpuTensor(self.shape, Buffer(device, h, arr.nbytes), dtype)


class ColumnMajor:
    """tensor.rs:17-33 — `MatrixOrdering` marker: element (i, j) at offset + i + j * stride (shape.wgsl:59-62)."""

    @staticmethod
    def is_row_major() -> bool:
        return False

    @staticmethod
    def is_column_major() -> bool:
        return True


class RowMajor:
    """tensor.rs:19-39 — element (i, j) at offset + i * stride + j (shape.wgsl:49-53)."""

    @staticmethod
    def is_row_major() -> bool:
        return True

    @staticmethod
    def is_column_major() -> bool:
        return False


class GpuTensorView:
    """:416-420 — a borrowed (buffer, ViewShape) pair; `dim` is the view's tensor order, `ordering` the reference's
    `Ordering` type parameter (ColumnMajor unless stated)."""

    def __init__(self, view_shape: ViewShape, buffer: Buffer, dtype: str, dim: int, ordering=ColumnMajor):
        self.view_shape, self._buffer, self.dtype, self.dim, self.ordering = view_shape, buffer, dtype, dim, ordering

    def shape(self) -> ViewShape:                                   # :424-426
        return self.view_shape

    def buffer(self) -> Buffer:                                     # :429-431
        return self._buffer

    # GpuVectorView (:434-463)
    def is_empty(self) -> bool:
        return self.len() == 0

    def len(self) -> int:
        return self.view_shape.size[0]

    def _with(self, size, stride, stride_mat, offset, dim) -> "GpuTensorView":
        return GpuTensorView(ViewShape(tuple(size), stride, stride_mat, offset), self._buffer, self.dtype, dim, self.ordering)

    # The reference's sub-view constructors are generic over `Ordering` but use the column-major offset formulas for both
    # (:484-510); here a row-major view steps by its own addressing (rows are `stride` apart, columns 1 apart).
    def rows(self, first_row: int, nrows: int) -> "GpuTensorView":
        s = self.view_shape
        if self.dim == 1:                                           # :445-462
            assert first_row + nrows <= self.len(), f"Rows slice range out of bounds: {first_row}..{first_row + nrows}"
            return self._with((nrows, 1, 1), s.stride, s.stride_mat, s.offset + first_row, 1)
        step = s.stride if self.ordering.is_row_major() else 1
        return self._with((nrows, s.size[1], 1), s.stride, s.stride_mat, s.offset + step * first_row, 2)   # :498-510

    def columns(self, first_col: int, ncols: int) -> "GpuTensorView":   # :484-496
        s = self.view_shape
        step = 1 if self.ordering.is_row_major() else s.stride
        return self._with((s.size[0], ncols, 1), s.stride, s.stride_mat, s.offset + step * first_col, 2)

    def matrix(self, matrix_id: int) -> "GpuTensorView":            # :466-481 (GpuCubeView::matrix)
        s = self.view_shape
        assert matrix_id < s.size[2]
        return self._with((s.size[0], s.size[1], 1), s.stride, 1, s.offset + s.stride_mat * matrix_id, 2)


class GpuTensor:
    """:192-399 (GpuScalar / GpuVector / GpuMatrix / GpuCube are this class with 0..3 dims)."""

    def __init__(self, shape: Sequence[int], buffer: Buffer, dtype: str = "f32"):
        self._shape, self._buffer, self.dtype = tuple(shape), buffer, dtype

    # -- :198-219
    def is_empty(self) -> bool:
        return self.len() == 0

    def len(self) -> int:
        return prod(self._shape)

    def bytes_len(self) -> int:
        return _DT[self.dtype][1] * self.len()

    def shape(self):
        return self._shape

    def buffer(self) -> Buffer:
        return self._buffer

    def into_inner(self) -> Buffer:
        return self._buffer

    # -- copies :227-265
    def copy_from(self, encoder, source: "GpuTensor") -> None:
        assert self.len() == source.len()
        check(lib().wgb_buffer_copy(self._buffer.device._h, None, self._buffer._h, 0, source._buffer._h, 0, self.bytes_len()))

    def copy_from_view(self, encoder, source) -> None:
        source = as_view(source, max(len(self._shape), 1))
        assert source.view_shape.size[0] == (1 if len(self._shape) == 0 else self._shape[0])
        es = _DT[self.dtype][1]
        check(lib().wgb_buffer_copy(self._buffer.device._h, None, self._buffer._h, 0, source.buffer()._h,
                                    source.view_shape.offset * es, self.bytes_len()))

    # -- views :282-297, :514-541
    def as_view(self, ordering=ColumnMajor) -> GpuTensorView:
        """`as_view::<Ordering>()`: the same buffer read with the given ordering (a RowMajor view of an r x c tensor
        expects the r rows stored one after the other)."""
        return self.as_embedded_view(len(self._shape), ordering)

    def as_embedded_view(self, dim2: int = 3, ordering=ColumnMajor) -> GpuTensorView:
        assert dim2 >= len(self._shape), "Can only embed into a higher-order tensor view."
        embedded = [1] * dim2
        embedded[:len(self._shape)] = self._shape
        return self.reshape(embedded, ordering=ordering)

    def reshape(self, shape: Sequence[int], stride: Optional[int] = None, stride_mat: Optional[int] = None,
                ordering=ColumnMajor) -> GpuTensorView:
        shape = [int(s) for s in shape]
        assert prod(shape) <= self.len()                            # :520
        size = [1, 1, 1]
        size[:len(shape)] = shape
        s0 = shape[0] if len(shape) > 0 else 1
        s1 = shape[1] if len(shape) > 1 else 1
        default_stride = s0 if ordering.is_column_major() else s1   # :525-529
        return GpuTensorView(ViewShape(tuple(size), default_stride if stride is None else stride,
                                       s0 * s1 if stride_mat is None else stride_mat, 0),
                             self._buffer, self.dtype, len(shape), ordering)

    # -- GpuMatrix :561-626
    def column(self, i: int) -> GpuTensorView:                      # :574-585
        r = self._shape[0]
        return GpuTensorView(ViewShape((r, 1, 1), 1, 1, r * i), self._buffer, self.dtype, 1)

    def slice(self, ij, dims) -> GpuTensorView:                     # :587-598
        (i, j), (nrows, ncols) = ij, dims
        r, c = self._shape[0], self._shape[1]
        # NOTE the reference computes `offset = i + j * nrows` with the *slice's* nrows (:594); that addresses
        # the wrong column for j > 0 unless nrows == parent rows.  Deliberately fixed here: parent rows.
        return GpuTensorView(ViewShape((nrows, ncols, 1), r, r * c, i + j * r), self._buffer, self.dtype, 2)

    def columns(self, first_col: int, ncols: int) -> GpuTensorView:  # :600-612
        r, c = self._shape[0], self._shape[1]
        return GpuTensorView(ViewShape((r, ncols, 1), r, r * c, first_col * r), self._buffer, self.dtype, 2)

    def rows(self, first_row: int, nrows: int) -> GpuTensorView:
        if len(self._shape) == 1:                                   # GpuVector::rows :669-681
            n = self._shape[0]
            return GpuTensorView(ViewShape((nrows, 1, 1), n, n, first_row), self._buffer, self.dtype, 1)
        r, c = self._shape[0], self._shape[1]                       # GpuMatrix::rows :614-626
        return GpuTensorView(ViewShape((nrows, c, 1), r, r * c, first_row), self._buffer, self.dtype, 2)

    # -- read-back :300-384
    def read(self, device=None) -> np.ndarray:
        out = np.empty(self.len(), dtype=_DT[self.dtype][0])
        self.read_to(device, out)
        return out

    def read_to(self, device, out: np.ndarray) -> None:
        assert out.flags["C_CONTIGUOUS"] and out.nbytes == self.bytes_len()
        check(lib().wgb_buffer_read(self._buffer.device._h, self._buffer._h, 0, out.ctypes.data_as(ctypes.c_void_p), out.nbytes))

    def slow_read(self, gpu=None) -> np.ndarray:                    # :340-355 (no staging needed with CUDA)
        return self.read()

    # -- constructors :544-705
    @staticmethod
    def init(device, data, usage: int, dtype: Optional[str] = None) -> "GpuTensor":
        """GpuMatrix::init (:561-571, takes a numpy matrix and stores it column-major) /
        GpuVector::init (:659-666) / GpuScalar::init (:699-704)."""
        arr = np.asarray(data)
        if arr.ndim == 2:
            return TensorBuilder.matrix(arr.shape[0], arr.shape[1], usage).build_init(device, np.asfortranarray(arr).reshape(-1, order="F"), dtype)
        if arr.ndim == 1:
            return TensorBuilder.vector(arr.shape[0], usage).build_init(device, arr, dtype)
        return TensorBuilder.scalar(usage).build_init(device, arr.reshape(1), dtype)

    @staticmethod
    def uninit(device, shape: Sequence[int], usage: int, dtype: str = "f32") -> "GpuTensor":
        return TensorBuilder.tensor(shape, usage).build(device, dtype)


GpuScalar = GpuVector = GpuMatrix = GpuCube = GpuTensor
GpuScalarView = GpuVectorView = GpuMatrixView = GpuCubeView = GpuTensorView


def as_view(x, dim: int) -> GpuTensorView:
    """`impl Into<GpuTensorView<..., DIM>>` (:403-409): tensors embed into a higher-order view."""
    if isinstance(x, GpuTensorView):
        return x
    if isinstance(x, GpuTensor):
        return x.as_embedded_view(max(dim, len(x.shape())))
    raise TypeError(f"expected a GpuTensor or GpuTensorView, got {type(x).__name__}")
