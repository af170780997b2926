"""Gemm / Gemv / OpAssign / Reduce — the host-side mirror of
/root/reference/crates/wgebra/src/linalg/{gemm,gemv,op_assign,reduce}.rs: same names, argument
order and error behaviour (dimension mismatches raise, empty dispatches are skipped), calling the
sm_100a kernels through the C ABI instead of composing WGSL pipelines."""
from __future__ import annotations

import ctypes
import enum

import numpy as np

from ._lib import BF16, F32, check, lib
from .tensor import GpuTensor, as_view

_DTYPE_CODE = {"f32": F32, "bf16": BF16}


class GemmVariant(enum.IntEnum):       # gemm.rs:25-35
    Gemm = 0
    GemmFast = 1
    GemmTr = 2
    GemmTrFast = 3


class GemvVariant(enum.IntEnum):       # gemv.rs:24-34
    Gemv = 0
    GemvFast = 1
    GemvTr = 2
    GemvTrFast = 3


class OpAssignVariant(enum.IntEnum):   # op_assign.rs:15-26
    Add = 0
    Sub = 1
    Mul = 2
    Div = 3
    Copy = 4


class ReduceOp(enum.IntEnum):          # reduce.rs:16-27
    Min = 0
    Max = 1
    Sum = 2
    Prod = 3
    SqNorm = 4


class F32Mode(enum.IntEnum):           # wgb_f32_mode
    Auto = 0
    X3Tf32 = 1
    Tf32 = 2
    Simt = 3


def _ord(view) -> int:
    return 1 if view.ordering.is_row_major() else 0   # wgb_ordering


class Gemm:
    """gemm.rs:9-127.  The four `ComputePipeline` fields of the reference become kernel-family tags: the CUDA
    kernels are compiled into the library, so `from_device` only checks that the device is usable."""

    def __init__(self, device):
        self.gemm, self.gemm_fast, self.gemm_tr, self.gemm_tr_fast = "gemm", "gemm_fast", "gemm_tr", "gemm_tr_fast"
        self._device = device
        self.f32_mode = F32Mode.Auto

    @staticmethod
    def from_device(device) -> "Gemm":
        return Gemm(device)

    def dispatch(self, device, shapes, pass_, out, m1, m2):                       # gemm.rs:39-49
        self.dispatch_generic(device, shapes, pass_, out, m1, m2, GemmVariant.Gemm)

    def dispatch_tr(self, device, shapes, pass_, out, m1, m2):                    # gemm.rs:52-62
        self.dispatch_generic(device, shapes, pass_, out, m1, m2, GemmVariant.GemmTr)

    def dispatch_generic(self, device, shapes, pass_, out, m1, m2, variant: GemmVariant, f32_mode=None):  # gemm.rs:65-127
        out, m1, m2 = as_view(out, 3), as_view(m1, 3), as_view(m2, 3)
        if m1.dtype != m2.dtype:
            raise TypeError("Gemm: m1 and m2 must have the same element type")
        so, s1, s2 = (shapes.get(device, v.shape()).to_c() for v in (out, m1, m2))   # gemm.rs:98-100
        mode = self.f32_mode if f32_mode is None else f32_mode
        if any(v.ordering.is_row_major() for v in (out, m1, m2)):
            # GpuTensorView<.., RowMajor, ..> operands (tensor.rs:19-39; shape.wgsl:49-57): the ordering travels per operand
            check(lib().wgb_gemm_ord(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), _ord(out), m1.buffer()._h,
                                     ctypes.byref(s1), _ord(m1), m2.buffer()._h, ctypes.byref(s2), _ord(m2),
                                     _DTYPE_CODE[m1.dtype], _DTYPE_CODE[out.dtype], int(mode), -1, None, None))
            return
        check(lib().wgb_gemm_ex(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), m1.buffer()._h, ctypes.byref(s1),
                                m2.buffer()._h, ctypes.byref(s2), _DTYPE_CODE[m1.dtype], _DTYPE_CODE[out.dtype], int(mode)))


    def dispatch_reduce(self, device, shapes, pass_, result, m1, m2, reduce_op: "ReduceOp", axis: int = 1, variant=GemmVariant.Gemm,
                        f32_mode=None):
        """result[j] = reduce_op over the column j of m1 * m2 (axis=1), or result[i] over the row i (axis=2), in one pass over the
        operands (wgb_gemm_reduce): the reference's Gemm::dispatch followed by one Reduce::dispatch per GpuMatrix::column(j), with
        the product never stored.  `result` is a vector (or vector view) of N (axis 1) or M (axis 2) f32."""
        r, m1, m2 = as_view(result, 3), as_view(m1, 3), as_view(m2, 3)
        rs, s1, s2 = (shapes.get(device, v.shape()).to_c() for v in (r, m1, m2))
        mode = self.f32_mode if f32_mode is None else f32_mode
        check(lib().wgb_gemm_reduce(pass_._h, int(variant), axis, int(reduce_op), r.buffer()._h, ctypes.byref(rs), m1.buffer()._h,
                                    ctypes.byref(s1), m2.buffer()._h, ctypes.byref(s2), _DTYPE_CODE[m1.dtype], int(mode)))

    def dispatch_op(self, device, shapes, pass_, out, m1, m2, op: "OpAssignVariant", operand, variant=GemmVariant.Gemm, f32_mode=None):
        """out = (m1 * m2) (op) operand in one launch (wgb_gemm_op): Gemm::dispatch + OpAssign::dispatch(out, operand) fused
        into the GEMM epilogue.  `operand` may be `out` itself (accumulate into out)."""
        out, m1, m2, e = as_view(out, 3), as_view(m1, 3), as_view(m2, 3), as_view(operand, 3)
        so, s1, s2, se = (shapes.get(device, v.shape()).to_c() for v in (out, m1, m2, e))
        mode = self.f32_mode if f32_mode is None else f32_mode
        if any(v.ordering.is_row_major() for v in (out, m1, m2, e)):
            if e.ordering is not out.ordering:
                raise TypeError("Gemm.dispatch_op: operand and out must share one ordering")
            check(lib().wgb_gemm_ord(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), _ord(out), m1.buffer()._h,
                                     ctypes.byref(s1), _ord(m1), m2.buffer()._h, ctypes.byref(s2), _ord(m2),
                                     _DTYPE_CODE[m1.dtype], _DTYPE_CODE[out.dtype], int(mode), int(op), e.buffer()._h,
                                     ctypes.byref(se)))
            return
        check(lib().wgb_gemm_op(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), m1.buffer()._h, ctypes.byref(s1),
                                m2.buffer()._h, ctypes.byref(s2), _DTYPE_CODE[m1.dtype], _DTYPE_CODE[out.dtype], int(mode), int(op),
                                e.buffer()._h, ctypes.byref(se)))

    def dispatch_host(self, device, M: int, N: int, K: int, out_host, m1_host, m2_host, variant=GemmVariant.Gemm,
                      in_dtype: str = "f32", out_dtype: str = "f32", f32_mode=None, n_panels: int = 0) -> None:
        """Host-buffer GEMM (wgb_gemm_host): upload, multiply and download pipelined by column panel.  `*_host` are
        numpy arrays or raw host pointers (ctypes.c_void_p) of dense column-major matrices.  Blocking."""
        def ptr(x):
            return x if isinstance(x, ctypes.c_void_p) else x.ctypes.data_as(ctypes.c_void_p)
        mode = self.f32_mode if f32_mode is None else f32_mode
        check(lib().wgb_gemm_host(device._h, int(variant), M, N, K, ptr(out_host), ptr(m1_host), ptr(m2_host),
                                  _DTYPE_CODE[in_dtype], _DTYPE_CODE[out_dtype], int(mode), n_panels))

    def enqueue_host(self, device, M: int, N: int, K: int, out_host, m1_host, m2_host, variant=GemmVariant.Gemm,
                     in_dtype: str = "f32", out_dtype: str = "f32", f32_mode=None, n_panels: int = 0) -> None:
        """wgb_gemm_host_enqueue: the same product queued without waiting (wgpu's submit-now / map-later model,
        tensor.rs:300-384).  `out_host` is complete after `device.poll_wait()`; consecutive products overlap on the host
        link (download of product i under the upload of product i+1)."""
        def ptr(x):
            return x if isinstance(x, ctypes.c_void_p) else x.ctypes.data_as(ctypes.c_void_p)
        mode = self.f32_mode if f32_mode is None else f32_mode
        check(lib().wgb_gemm_host_enqueue(device._h, int(variant), M, N, K, ptr(out_host), ptr(m1_host), ptr(m2_host),
                                          _DTYPE_CODE[in_dtype], _DTYPE_CODE[out_dtype], int(mode), n_panels))

    @staticmethod
    def flush_host(device) -> None:
        """wgb_gemm_host_flush: later work on the queue (e.g. a timestamp) waits for every enqueued product's download."""
        check(lib().wgb_gemm_host_flush(device._h))


class Gemv:
    """gemv.rs:9-137."""

    def __init__(self, device):
        self.gemv, self.gemv_fast, self.gemv_tr, self.gemv_tr_fast = "gemv", "gemv_fast", "gemv_tr", "gemv_tr_fast"
        self._device = device

    @staticmethod
    def from_device(device) -> "Gemv":
        return Gemv(device)

    def dispatch(self, device, shapes, pass_, out, m, v):                         # gemv.rs:38-48
        self.dispatch_generic(device, shapes, pass_, out, m, v, GemvVariant.Gemv)

    def dispatch_tr(self, device, shapes, pass_, out, m, v):                      # gemv.rs:51-61
        self.dispatch_generic(device, shapes, pass_, out, m, v, GemvVariant.GemvTr)

    def dispatch_generic(self, device, shapes, pass_, out, m, v, variant: GemvVariant):   # gemv.rs:64-137
        out, m, v = as_view(out, 3), as_view(m, 3), as_view(v, 3)
        so, sm, sv = (shapes.get(device, x.shape()).to_c() for x in (out, m, v))
        if m.ordering.is_row_major():
            check(lib().wgb_gemv_ord(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), m.buffer()._h, ctypes.byref(sm),
                                     1, v.buffer()._h, ctypes.byref(sv)))
            return
        check(lib().wgb_gemv(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), m.buffer()._h, ctypes.byref(sm),
                             v.buffer()._h, ctypes.byref(sv)))

    def dispatch_op(self, device, shapes, pass_, out, m, v, op: "OpAssignVariant", operand, variant=GemvVariant.Gemv):
        """out = (m * v) (op) operand in one launch (wgb_gemv_op): Gemv::dispatch + OpAssign::dispatch(out, operand) fused into
        the GEMV's store.  `operand` may be `out` itself (the residual update out = m * v + out)."""
        out, m, v, e = as_view(out, 3), as_view(m, 3), as_view(v, 3), as_view(operand, 3)
        so, sm, sv, se = (shapes.get(device, x.shape()).to_c() for x in (out, m, v, e))
        check(lib().wgb_gemv_op(pass_._h, int(variant), out.buffer()._h, ctypes.byref(so), m.buffer()._h, ctypes.byref(sm),
                                1 if m.ordering.is_row_major() else 0, v.buffer()._h, ctypes.byref(sv), int(op), e.buffer()._h,
                                ctypes.byref(se)))

    def dispatch_reduce(self, device, shapes, pass_, result, m, v, reduce_op: "ReduceOp", variant=GemvVariant.Gemv):
        """result = reduce_op(m * v) in one launch (wgb_gemv_reduce): Gemv::dispatch + Reduce::dispatch(out, result) with the product
        vector never leaving the library; bit-identical to the two-dispatch chain through a 16-byte aligned `out`."""
        m, v = as_view(m, 3), as_view(v, 3)
        sm, sv = shapes.get(device, m.shape()).to_c(), shapes.get(device, v.shape()).to_c()
        check(lib().wgb_gemv_reduce(pass_._h, int(variant), int(reduce_op), result.buffer()._h, m.buffer()._h, ctypes.byref(sm),
                                    1 if m.ordering.is_row_major() else 0, v.buffer()._h, ctypes.byref(sv)))


class OpAssign:
    """op_assign.rs:43-94: `OpAssign(pipeline, variant)` -> fields `.0` / `.1` are `pipeline` / `variant`."""
    SRC = "op_assign.wgsl (replaced by wgmath_b200/csrc/level1.cu)"
    FILE_PATH = "wgebra/src/op_assign.wgsl"

    def __init__(self, device, op: OpAssignVariant):
        self.pipeline, self.variant = "op_assign", OpAssignVariant(op)

    @staticmethod
    def new(device, op: OpAssignVariant) -> "OpAssign":                           # op_assign.rs:52-67
        return OpAssign(device, op)

    def dispatch(self, device, shapes, pass_, in_out_a, in_b):                    # op_assign.rs:71-94
        a, b = as_view(in_out_a, 1), as_view(in_b, 1)
        sa, sb = shapes.get(device, a.shape()).to_c(), shapes.get(device, b.shape()).to_c()
        check(lib().wgb_op_assign(pass_._h, int(self.variant), a.buffer()._h, ctypes.byref(sa), b.buffer()._h, ctypes.byref(sb)))


class Reduce:
    """reduce.rs:62-124."""
    SRC = "reduce.wgsl (replaced by wgmath_b200/csrc/level1.cu)"
    FILE_PATH = "wgebra/src/reduce.wgsl"

    def __init__(self, device, op: ReduceOp):
        self.pipeline, self.op = "reduce", ReduceOp(op)

    @staticmethod
    def new(device, op: ReduceOp) -> "Reduce":                                    # reduce.rs:71-96
        return Reduce(device, op)

    def dispatch(self, device, shapes, pass_, value, result: GpuTensor):          # reduce.rs:100-113
        v = as_view(value, 1)
        sv = shapes.get(device, v.shape()).to_c()
        check(lib().wgb_reduce(pass_._h, int(self.op), v.buffer()._h, ctypes.byref(sv), result.buffer()._h))

    def dispatch_columns(self, device, shapes, pass_, matrix, out):
        """Extension (wgb_reduce_columns): Reduce over every column of a matrix view in one launch."""
        m, o = as_view(matrix, 3), as_view(out, 1)
        sm, so_ = shapes.get(device, m.shape()).to_c(), shapes.get(device, o.shape()).to_c()
        check(lib().wgb_reduce_columns(pass_._h, int(self.op), m.buffer()._h, ctypes.byref(sm), o.buffer()._h, ctypes.byref(so_)))

    def eval_cpu(self, val: np.ndarray) -> float:                                 # reduce.rs:116-124
        val = np.asarray(val, dtype=np.float32)
        return float({ReduceOp.Min: val.min, ReduceOp.Max: val.max, ReduceOp.Prod: val.prod, ReduceOp.Sum: val.sum,
                      ReduceOp.SqNorm: lambda: (val * val).sum()}[self.op]())


class Dot:
    """Extension (wgb_dot): result = sum_i a[i] * b[i].  The reference only has Reduce(SqNorm)."""

    @staticmethod
    def new(device) -> "Dot":
        return Dot()

    def dispatch(self, device, shapes, pass_, a, b, result: GpuTensor):
        a, b = as_view(a, 1), as_view(b, 1)
        sa, sb = shapes.get(device, a.shape()).to_c(), shapes.get(device, b.shape()).to_c()
        check(lib().wgb_dot(pass_._h, a.buffer()._h, ctypes.byref(sa), b.buffer()._h, ctypes.byref(sb), result.buffer()._h))


def fill_uniform(device, pass_, target, seed: int, row0: int = 0, col0: int = 0) -> None:
    """Seeded U[0,1) fill in HBM (wgb_fill_uniform); bit-identical to oracle.uniform()."""
    v = as_view(target, 3)
    s = v.shape().to_c()
    check(lib().wgb_fill_uniform(pass_._h, v.buffer()._h, ctypes.byref(s), _DTYPE_CODE[v.dtype], seed, row0, col0))
