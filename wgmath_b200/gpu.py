"""GpuInstance / Device / Queue / CommandEncoder / ComputePass — the host-side mirror of
/root/reference/crates/wgcore/src/gpu.rs:7-79 and kernel.rs:7-27 over the C ABI.

wgpu's record-then-submit model maps onto one in-order CUDA stream per device: a dispatch is
enqueued when it is recorded, `queue.submit` only flushes, and `read` / `poll(wait)` synchronise."""
from __future__ import annotations

import ctypes
from typing import Optional

from . import _lib
from ._lib import check, lib


class Device:
    """Stands in for wgpu::Device on this path: owns the wgb_ctx (device ordinal + queue stream)."""

    def __init__(self, ordinal: int = 0):
        h = ctypes.c_void_p()
        check(lib().wgb_ctx_create(ordinal, ctypes.byref(h)))
        self._h = h
        self.ordinal = ordinal

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().wgb_ctx_destroy(h)
            except Exception:
                pass

    def create_command_encoder(self, desc=None) -> "CommandEncoder":
        return CommandEncoder(self)

    def capture(self) -> "_Capture":
        """`with device.capture() as cap: <dispatches>` records them; `cap.graph.launch()` replays (wgb_graph_*)."""
        return _Capture(self)

    def poll_wait(self) -> None:
        """device.poll(PollType::wait()) (tensor.rs:304-312)."""
        check(lib().wgb_ctx_sync(self._h))

    def info(self) -> dict:
        sm, maj, mn = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        mem = ctypes.c_size_t()
        name = ctypes.create_string_buffer(256)
        check(lib().wgb_ctx_device_info(self._h, ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mn), ctypes.byref(mem), name, 256))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mn.value), "total_mem": mem.value}

    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        check(lib().wgb_ctx_launch_count(self._h, ctypes.byref(n)))
        return n.value


class Graph:
    """A recorded dispatch sequence (CUDA graph): `Device.capture()` ... `with` body ... then `.launch()` replays it."""

    def __init__(self, device: "Device", handle):
        self._device, self._h = device, handle

    def launch(self) -> None:
        check(lib().wgb_graph_launch(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().wgb_graph_destroy(h)
            except Exception:
                pass


class _Capture:
    def __init__(self, device: "Device"):
        self._device, self.graph = device, None

    def __enter__(self):
        check(lib().wgb_graph_capture_begin(self._device._h))
        return self

    def __exit__(self, et, ev, tb):
        h = ctypes.c_void_p()
        st = lib().wgb_graph_capture_end(self._device._h, ctypes.byref(h))
        if et is None:
            check(st)
            self.graph = Graph(self._device, h)
        return False


class Queue:
    def __init__(self, device: Device):
        self._device = device

    def submit(self, command_buffers=None) -> None:
        """queue.submit(Some(encoder.finish())) (gemm.rs:192)."""
        check(lib().wgb_submit(self._device._h))


class ComputePass:
    """kernel.rs:15-26.  `end()` is Rust's `drop(pass)`; also usable as a context manager."""

    def __init__(self, encoder: "CommandEncoder", label: str, timestamps=None):
        begin = end = None
        if timestamps is not None:
            begin, end = timestamps.next_compute_pass_timestamp_writes()
        h = ctypes.c_void_p()
        check(lib().wgb_pass_begin(encoder.device._h, label.encode(), begin, end, ctypes.byref(h)))
        self._h = h
        self.device = encoder.device

    def end(self) -> None:
        h, self._h = self._h, None
        if h:
            check(lib().wgb_pass_end(h))

    def last_gemm_path(self) -> int:
        p = ctypes.c_int()
        check(lib().wgb_pass_last_gemm_path(self._h, ctypes.byref(p)))
        return p.value

    def last_gemm_config(self) -> dict:
        """The tcgen05 kernel instantiation / plan of the last GEMM on this pass (wgb_pass_last_gemm_config)."""
        c = (ctypes.c_int * 13)()
        check(lib().wgb_pass_last_gemm_config(self._h, c))
        keys = ("kind", "a_mn", "b_mn", "bn", "passes", "out_dtype", "cg", "epi_tma", "nsplit", "splitk", "dests", "units", "fused_split")
        return dict(zip(keys, list(c)))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.end()

    def __del__(self):
        try:
            self.end()
        except Exception:
            pass


class CommandEncoder:
    def __init__(self, device: Device):
        self.device = device

    def compute_pass(self, label: str, timestamps=None) -> ComputePass:
        """CommandEncoderExt::compute_pass (kernel.rs:7-27)."""
        return ComputePass(self, label, timestamps)

    def finish(self):
        return self


class GpuInstance:
    """gpu.rs:7-79."""

    def __init__(self, ordinal: int = 0):
        self._device = Device(ordinal)
        self._queue = Queue(self._device)

    @staticmethod
    def new(ordinal: int = 0) -> "GpuInstance":
        return GpuInstance(ordinal)

    without_gl = new
    with_backends = new

    def device(self) -> Device:
        return self._device

    device_arc = device

    def queue(self) -> Queue:
        return self._queue

    def adapter(self) -> dict:
        return self._device.info()
