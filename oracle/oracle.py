"""ctypes front-end for the CPU oracle (oracle/wgsl_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of wgsl_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never by the
product package wgmath_b200/.

Also holds the seeded synthetic-input generator: a counter-based
splitmix64 stream, value = f(seed, i, j) independent of sharding, U[0,1) with 24 random
bits (the distribution of nalgebra's `new_random`, reference gemm.rs:152-153).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwgsl_oracle.so")

GEMM, GEMM_FAST, GEMM_TR, GEMM_TR_FAST = 0, 1, 2, 3
GEMV, GEMV_FAST, GEMV_TR, GEMV_TR_FAST = 0, 1, 2, 3
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_COPY = 0, 1, 2, 3, 4
RED_MIN, RED_MAX, RED_SUM, RED_PROD, RED_SQNORM = 0, 1, 2, 3, 4
ORC_OK, ORC_DIM_MISMATCH = 0, 2

SEED_BASE = 0x5EED0000  # + operand index: A=1, B=2, v=3, out=4


class Shape(ctypes.Structure):
    """Byte-identical to wgcore::shapes::ViewShape (crates/wgcore/src/shapes.rs:9-21)."""

    _fields_ = [("nrows", ctypes.c_uint32), ("ncols", ctypes.c_uint32), ("nmats", ctypes.c_uint32),
                ("stride", ctypes.c_uint32), ("stride_mat", ctypes.c_uint32), ("offset", ctypes.c_uint32)]


def shape(nrows, ncols=1, nmats=1, stride=None, stride_mat=None, offset=0) -> Shape:
    stride = nrows if stride is None else stride
    stride_mat = nrows * ncols if stride_mat is None else stride_mat
    return Shape(nrows, ncols, nmats, stride, stride_mat, offset)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("wgsl_oracle.c", "scan_sort_oracle.c", "geometry_oracle.c")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH) for src in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        sp = ctypes.POINTER(Shape)
        L.orc_gemm.argtypes = [ctypes.c_int, fp, sp, fp, sp, fp, sp]
        L.orc_gemv.argtypes = [ctypes.c_int, fp, sp, fp, sp, fp, sp, ctypes.POINTER(ctypes.c_int)]
        L.orc_op_assign.argtypes = [ctypes.c_int, fp, sp, fp, sp]
        L.orc_reduce.argtypes = [ctypes.c_int, fp, sp, fp]
        L.orc_gemm_f64.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_uint32,
                                   ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, fp, sp, fp, sp]
        L.orc_gemm_f64.restype = None
        L.orc_gemm_ord.argtypes = [ctypes.c_int, fp, sp, ctypes.c_int, fp, sp, ctypes.c_int, fp, sp, ctypes.c_int]
        L.orc_gemv_ord.argtypes = [ctypes.c_int, fp, sp, fp, sp, ctypes.c_int, fp, sp]
        up = ctypes.POINTER(ctypes.c_uint32)
        L.orc_prefix_sum.argtypes = [up, ctypes.c_uint32]
        L.orc_radix_sort.argtypes = [up, up, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, up, up]
        L.orc_geom_batch.argtypes = [ctypes.c_int, ctypes.c_int, fp, ctypes.c_void_p, ctypes.c_uint64]
        L.orc_geom_out_words.argtypes = [ctypes.c_int, ctypes.c_int]
        L.orc_geom_out_words.restype = ctypes.c_uint32
        L.orc_geom_in_words.argtypes = [ctypes.c_int]
        L.orc_geom_in_words.restype = ctypes.c_uint32
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_shape_index.restype = ctypes.c_uint32
        L.orc_shape_index.argtypes = [ctypes.POINTER(Shape), ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.POINTER(ctypes.c_uint32)]
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def gemm(variant, out, so, m1, s1, m2, s2) -> int:
    """out/m1/m2: flat float32 buffers (the wgpu storage buffers); s*: Shape views into them."""
    return lib().orc_gemm(variant, _fp(out), ctypes.byref(so), _fp(m1), ctypes.byref(s1), _fp(m2), ctypes.byref(s2))


def gemv(variant, out, so, m, sm, v, sv):
    ran = ctypes.c_int(-1)
    rc = lib().orc_gemv(variant, _fp(out), ctypes.byref(so), _fp(m), ctypes.byref(sm), _fp(v), ctypes.byref(sv),
                        ctypes.byref(ran))
    return rc, ran.value


def gemm_ord(variant, out, so, out_rm, m1, s1, m1_rm, m2, s2, m2_rm) -> int:
    """gemm on views that are each column-major (0) or row-major (1): shape.wgsl:49-57 addressing."""
    return lib().orc_gemm_ord(variant, _fp(out), ctypes.byref(so), int(out_rm), _fp(m1), ctypes.byref(s1), int(m1_rm),
                              _fp(m2), ctypes.byref(s2), int(m2_rm))


def gemv_ord(variant, out, so, m, sm, m_rm, v, sv) -> int:
    return lib().orc_gemv_ord(variant, _fp(out), ctypes.byref(so), _fp(m), ctypes.byref(sm), int(m_rm), _fp(v),
                              ctypes.byref(sv))


def shape_index(s: Shape, row_major: bool, fn: int, i: int, j: int = 0, t: int = 0):
    """shape.wgsl's iv (fn 0) / im (1) / it (2) and with_vec4_elts for a column- or row-major view: (index, 6 u32 of the vec4 shape)"""
    v4 = (ctypes.c_uint32 * 6)()
    idx = lib().orc_shape_index(ctypes.byref(s), int(row_major), fn, i, j, t, v4)
    return idx, tuple(v4)


def op_assign(op, a, sa, b, sb) -> int:
    return lib().orc_op_assign(op, _fp(a), ctypes.byref(sa), _fp(b), ctypes.byref(sb))


def reduce(op, x, s) -> float:
    r = ctypes.c_float(0.0)
    lib().orc_reduce(op, _fp(x), ctypes.byref(s), ctypes.byref(r))
    return r.value


def gemm_f64(tr, M, N, K, nmats, m1, s1, m2, s2) -> np.ndarray:
    out = np.empty(nmats * M * N, dtype=np.float64)
    lib().orc_gemm_f64(int(tr), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), M, N, K, nmats,
                       _fp(m1), ctypes.byref(s1), _fp(m2), ctypes.byref(s2))
    return out.reshape(nmats, N, M)  # [t][col][row]: column-major matrices


def num_threads() -> int:
    return lib().orc_num_threads()


def use_all_cores() -> int:
    """Overrides an inherited OMP_NUM_THREADS (torchrun sets it to 1): the CPU baseline uses every core the process may
    run on.  Returns the thread count now in effect."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(n)
    return num_threads()


# ---- seeded inputs -------------------------------------------------------------------

def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, nrows: int, ncols: int = 1, row0: int = 0, col0: int = 0) -> np.ndarray:
    """Column-major [ncols][nrows] float32 array of U[0,1): element (i,j) depends only on
    (seed, row0+i, col0+j).  Matches wgb_fill_uniform in the CUDA library bit for bit."""
    i = (np.arange(nrows, dtype=np.uint64) + np.uint64(row0))[None, :]
    j = (np.arange(ncols, dtype=np.uint64) + np.uint64(col0))[:, None]
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0xD1342543DE82EF95)) ^ ((j << np.uint64(32)) | i)
    u = _splitmix64(key)
    return ((u >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(ncols * nrows)


def to_bf16_rne(x: np.ndarray) -> np.ndarray:
    """Round float32 -> bfloat16 (round-to-nearest-even), returned as float32 values."""
    b = x.astype(np.float32).view(np.uint32)
    lsb = (b >> np.uint32(16)) & np.uint32(1)
    r = (b + np.uint32(0x7FFF) + lsb) & np.uint32(0xFFFF0000)
    return r.view(np.float32)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 (already bf16-representable or not) -> uint16 bf16 bit patterns (RNE)."""
    return (to_bf16_rne(x).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def bf16_from_bits(h: np.ndarray) -> np.ndarray:
    return (h.astype(np.uint32) << np.uint32(16)).view(np.float32)


def _up(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


def prefix_sum(data: np.ndarray) -> int:
    """WgPrefixSum::dispatch (wgrapier/src/dynamics/prefix_sum.rs:49-99): in-place exclusive u32 prefix sum."""
    return lib().orc_prefix_sum(_up(data), data.size)


def radix_sort(keys: np.ndarray, values: np.ndarray, n_sort: int, sorting_bits: int, out_keys: np.ndarray,
               out_values: np.ndarray) -> int:
    """RadixSort::dispatch (wgparry/src/utils/radix_sort/mod.rs:111-223)."""
    assert keys.size == values.size and out_keys.size >= keys.size and out_values.size >= keys.size
    return lib().orc_radix_sort(_up(keys), _up(values), keys.size, n_sort, sorting_bits, _up(out_keys), _up(out_values))


# wgebra::geometry batched factorizations (geometry_oracle.c)
GEOM_CHOLESKY, GEOM_LU, GEOM_QR, GEOM_EIG, GEOM_SVD, GEOM_INV = range(6)


def geom_in_words(dim: int) -> int:
    return lib().orc_geom_in_words(dim)


def geom_out_words(op: int, dim: int) -> int:
    return lib().orc_geom_out_words(op, dim)


def geom_batch(op: int, dim: int, mats: np.ndarray) -> np.ndarray:
    """out[i] = f(in[i]) in WGSL storage layout (the test kernels of wgebra/src/geometry/*.rs).  `mats` is a float32 array
    [n, geom_in_words(dim)]; returns float32 [n, geom_out_words(op, dim)] (view as uint32 for the LU permutation words)."""
    iw, ow = geom_in_words(dim), geom_out_words(op, dim)
    assert ow, f"unsupported geometry op {op} for dim {dim}"
    mats = np.ascontiguousarray(mats, np.float32).reshape(-1, iw)
    out = np.zeros((mats.shape[0], ow), np.float32)
    rc = lib().orc_geom_batch(op, dim, _fp(mats), out.ctypes.data_as(ctypes.c_void_p), mats.shape[0])
    assert rc == 0
    return out


def geom_pack(mats: np.ndarray) -> np.ndarray:
    """[n, dim, dim] matrices (mats[i][r][c]) -> WGSL storage layout [n, dim * col_stride] (column-major, vec3 columns padded)."""
    n, dim, _ = mats.shape
    cs = 2 if dim == 2 else 4
    out = np.zeros((n, dim, cs), np.float32)
    out[:, :, :dim] = np.transpose(mats, (0, 2, 1))
    return out.reshape(n, dim * cs)


def geom_unpack(words: np.ndarray, dim: int) -> np.ndarray:
    """Inverse of geom_pack for one matrix field: [n, dim * col_stride] -> [n, dim, dim] (row, column)."""
    cs = 2 if dim == 2 else 4
    return np.transpose(words.reshape(-1, dim, cs)[:, :, :dim], (0, 2, 1))
