"""ctypes binding of libwgebra_b200.so (include/wgb200.h).  There is no fallback: if the
library is missing or no sm_100 device is usable, importing callers get a loud error."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwgebra_b200.so")

OK, ERR_INVALID, ERR_DIM_MISMATCH, ERR_CUDA, ERR_UNSUPPORTED, ERR_OOM, ERR_NCCL, ERR_OOB, ERR_NO_DEVICE = range(9)
F32, BF16 = 0, 1
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


class WgbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[wgb status {status}] {msg}")
        self.status = status


class DimensionMismatch(AssertionError):
    """The reference panics (assert_eq!) on mismatched operand dimensions
    (gemm.rs:91-95, gemv.rs:89-90,122, op_assign.rs:82-86); the host shim raises this."""


class ViewShapeC(ctypes.Structure):
    _fields_ = [("size", ctypes.c_uint32 * 3), ("stride", ctypes.c_uint32), ("stride_mat", ctypes.c_uint32),
                ("offset", ctypes.c_uint32)]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -m wgmath_b200.build` "
                          "(nvcc, sm_100a). wgmath_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, sz, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
    pvp = ctypes.POINTER(vp)
    sp = ctypes.POINTER(ViewShapeC)
    sigs = {
        "wgb_abi_version": ([], ci),
        "wgb_last_error_string": ([], ctypes.c_char_p),
        "wgb_ctx_create": ([ci, pvp], ci),
        "wgb_ctx_destroy": ([vp], ci),
        "wgb_ctx_sync": ([vp], ci),
        "wgb_ctx_device_info": ([vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(sz),
                                 ctypes.c_char_p, sz], ci),
        "wgb_ctx_launch_count": ([vp, ctypes.POINTER(u64)], ci),
        "wgb_ctx_stream": ([vp, pvp], ci),
        "wgb_pass_begin": ([vp, ctypes.c_char_p, vp, vp, pvp], ci),
        "wgb_pass_end": ([vp], ci),
        "wgb_submit": ([vp], ci),
        "wgb_pass_last_gemm_path": ([vp, ctypes.POINTER(ci)], ci),
        "wgb_pass_last_gemm_config": ([vp, ctypes.POINTER(ci)], ci),
        "wgb_graph_capture_begin": ([vp], ci),
        "wgb_graph_capture_end": ([vp, pvp], ci),
        "wgb_graph_launch": ([vp], ci),
        "wgb_graph_destroy": ([vp], ci),
        "wgb_buffer_create": ([vp, sz, u32, pvp], ci),
        "wgb_buffer_create_init": ([vp, vp, sz, u32, pvp], ci),
        "wgb_buffer_wrap": ([vp, vp, sz, pvp], ci),
        "wgb_buffer_destroy": ([vp], ci),
        "wgb_buffer_size": ([vp, ctypes.POINTER(sz)], ci),
        "wgb_buffer_device_ptr": ([vp, pvp], ci),
        "wgb_buffer_write": ([vp, vp, sz, vp, sz], ci),
        "wgb_buffer_copy": ([vp, vp, vp, sz, vp, sz, sz], ci),
        "wgb_buffer_read": ([vp, vp, sz, vp, sz], ci),
        "wgb_host_alloc": ([sz, pvp], ci),
        "wgb_host_free": ([vp], ci),
        "wgb_gemm": ([vp, ci, vp, sp, vp, sp, vp, sp], ci),
        "wgb_gemm_ex": ([vp, ci, vp, sp, vp, sp, vp, sp, ci, ci, ci], ci),
        "wgb_gemm_op": ([vp, ci, vp, sp, vp, sp, vp, sp, ci, ci, ci, ci, vp, sp], ci),
        "wgb_gemm_ord": ([vp, ci, vp, sp, ci, vp, sp, ci, vp, sp, ci, ci, ci, ci, ci, vp, sp], ci),
        "wgb_gemv_ord": ([vp, ci, vp, sp, vp, sp, ci, vp, sp], ci),
        "wgb_gemv_op": ([vp, ci, vp, sp, vp, sp, ci, vp, sp, ci, vp, sp], ci),
        "wgb_gemv_reduce": ([vp, ci, ci, vp, vp, sp, ci, vp, sp], ci),
        "wgb_gemm_reduce": ([vp, ci, ci, ci, vp, sp, vp, sp, vp, sp, ci, ci], ci),
        "wgb_debug_tc_trace": ([vp, ci, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)], ci),
        "wgb_gemm_host": ([vp, ci, u32, u32, u32, vp, vp, vp, ci, ci, ci, ci], ci),
        "wgb_gemm_host_enqueue": ([vp, ci, u32, u32, u32, vp, vp, vp, ci, ci, ci, ci], ci),
        "wgb_gemm_host_flush": ([vp], ci),
        "wgb_gemv": ([vp, ci, vp, sp, vp, sp, vp, sp], ci),
        "wgb_op_assign": ([vp, ci, vp, sp, vp, sp], ci),
        "wgb_reduce": ([vp, ci, vp, sp, vp], ci),
        "wgb_prefix_sum": ([vp, vp, sp], ci),
        "wgb_radix_sort": ([vp, vp, sp, vp, sp, vp, u32, vp, sp, vp, sp], ci),
        "wgb_geometry_in_bytes": ([ci], u32),
        "wgb_geometry_out_bytes": ([ci, ci], u32),
        "wgb_geometry_batch": ([vp, ci, ci, vp, u64, vp, u64, u64], ci),
        "wgb_dot": ([vp, vp, sp, vp, sp, vp], ci),
        "wgb_reduce_columns": ([vp, ci, vp, sp, vp, sp], ci),
        "wgb_fill_uniform": ([vp, vp, sp, ci, u64, u32, u32], ci),
        "wgb_event_create": ([vp, pvp], ci),
        "wgb_event_destroy": ([vp], ci),
        "wgb_event_record": ([vp, vp], ci),
        "wgb_event_elapsed_ms": ([vp, vp, ctypes.POINTER(ctypes.c_float)], ci),
        "wgb_comm_get_unique_id": ([vp], ci),
        "wgb_comm_init_rank": ([vp, ci, ci, vp], ci),
        "wgb_comm_destroy": ([vp], ci),
        "wgb_gemm_row_sharded": ([vp, ci, vp, vp, sp, vp, sp, ci, ci, ci, ci], ci),
        "wgb_peer_gather_create": ([vp, ci, ci, sz, pvp], ci),
        "wgb_peer_gather_create_ex": ([vp, ci, ci, sz, ci, pvp], ci),
        "wgb_peer_gather_connect_local": ([vp, pvp], ci),
        "wgb_peer_gather_buffer_at": ([vp, ci, pvp], ci),
        "wgb_peer_gather_wait": ([vp, vp, ci], ci),
        "wgb_gemm_row_sharded_fused_ex": ([vp, ci, vp, vp, sp, vp, sp, ci, ci, ci, u32], ci),
        "wgb_peer_gather_export": ([vp, vp], ci),
        "wgb_peer_gather_connect": ([vp, vp], ci),
        "wgb_peer_gather_buffer": ([vp, pvp], ci),
        "wgb_peer_gather_destroy": ([vp], ci),
        "wgb_peer_gather_disconnect": ([vp], ci),
        "wgb_peer_gather_region_bytes": ([sz, ci], sz),
        "wgb_peer_gather_create_external": ([vp, ci, ci, sz, ci, pvp, vp, pvp], ci),
        "wgb_peer_gather_debug_flags": ([vp, ctypes.POINTER(u32)], ci),
        "wgb_debug_link_stream": ([vp, ci, pvp, ci, vp, sz, ci, ci, ctypes.POINTER(ctypes.c_float)], ci),
        "wgb_gemm_row_sharded_fused": ([vp, ci, vp, vp, sp, vp, sp, ci, ci, ci], ci),
        "wgb_gemm_row_sharded_fused_host_enqueue": ([vp, ci, vp, u32, u32, u32, vp, vp, vp, ci, ci, ci, ci], ci),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(L, name)  # AttributeError here == the .so does not export what wgb200.h declares
        fn.argtypes = argtypes
        fn.restype = restype
    if L.wgb_abi_version() != 1:
        raise ImportError("libwgebra_b200.so ABI version mismatch")
    _lib = L
    return L


EXPORTED = ["wgb_abi_version", "wgb_last_error_string", "wgb_ctx_create", "wgb_ctx_destroy", "wgb_ctx_sync",
            "wgb_ctx_device_info", "wgb_ctx_launch_count", "wgb_ctx_stream", "wgb_pass_begin", "wgb_pass_end",
            "wgb_submit", "wgb_pass_last_gemm_path", "wgb_pass_last_gemm_config", "wgb_graph_capture_begin", "wgb_graph_capture_end", "wgb_graph_launch",
            "wgb_graph_destroy", "wgb_buffer_create", "wgb_buffer_create_init", "wgb_buffer_wrap",
            "wgb_buffer_destroy", "wgb_buffer_size", "wgb_buffer_device_ptr", "wgb_buffer_write", "wgb_buffer_copy",
            "wgb_buffer_read", "wgb_host_alloc", "wgb_host_free", "wgb_gemm", "wgb_gemm_ex", "wgb_gemm_op", "wgb_gemm_ord", "wgb_gemm_host", "wgb_gemm_host_enqueue", "wgb_gemm_host_flush", "wgb_gemv",
            "wgb_gemv_ord", "wgb_gemv_op", "wgb_gemv_reduce", "wgb_gemm_reduce", "wgb_debug_tc_trace",
            "wgb_op_assign", "wgb_reduce", "wgb_prefix_sum", "wgb_radix_sort", "wgb_geometry_in_bytes", "wgb_geometry_out_bytes", "wgb_geometry_batch", "wgb_dot", "wgb_reduce_columns", "wgb_fill_uniform", "wgb_event_create",
            "wgb_event_destroy", "wgb_event_record", "wgb_event_elapsed_ms", "wgb_comm_get_unique_id",
            "wgb_comm_init_rank", "wgb_comm_destroy", "wgb_gemm_row_sharded", "wgb_peer_gather_create",
            "wgb_peer_gather_export", "wgb_peer_gather_connect", "wgb_peer_gather_buffer", "wgb_peer_gather_destroy",
            "wgb_gemm_row_sharded_fused", "wgb_gemm_row_sharded_fused_host_enqueue", "wgb_peer_gather_create_ex", "wgb_peer_gather_disconnect", "wgb_peer_gather_debug_flags", "wgb_debug_link_stream", "wgb_peer_gather_region_bytes", "wgb_peer_gather_create_external",
            "wgb_peer_gather_connect_local", "wgb_peer_gather_buffer_at", "wgb_peer_gather_wait", "wgb_gemm_row_sharded_fused_ex"]


def check(status: int) -> None:
    if status == OK:
        return
    msg = lib().wgb_last_error_string().decode("utf-8", "replace")
    if status == ERR_DIM_MISMATCH:
        raise DimensionMismatch(msg)
    raise WgbError(status, msg)
