"""Host-side mirror of the two integer primitives next to the linalg path:

  WgPrefixSum / PrefixSumWorkspace   wgrapier/src/dynamics/prefix_sum.rs:22-224
  RadixSort / RadixSortWorkspace     wgparry/src/utils/radix_sort/mod.rs:67-223

Same names, argument order and error behaviour as the reference; the work happens in csrc/scan_sort.cu through the C ABI
(wgb_prefix_sum / wgb_radix_sort).  The workspace types of the reference hold auxiliary GPU buffers; here those live in the
context, so the classes only keep the reference's interface (and its capacity bookkeeping, which callers may inspect)."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib
from .tensor import as_view


class PrefixSumWorkspace:
    """prefix_sum.rs:119-224.  `stages` mirrors the level lengths the reference would allocate (ceil(n / 256) ... 1)."""

    THREADS = 256

    def __init__(self):
        self.stages: list[int] = []
        self.num_stages = 0

    @staticmethod
    def new() -> "PrefixSumWorkspace":
        return PrefixSumWorkspace()

    @staticmethod
    def with_capacity(device, buffer_len: int) -> "PrefixSumWorkspace":
        ws = PrefixSumWorkspace()
        ws.reserve(device, buffer_len)
        return ws

    def reserve(self, device, buffer_len: int) -> None:            # :185-224
        stages = []
        stage_len = -(-buffer_len // self.THREADS)
        while stage_len > 1:                                       # (the reference loops forever for buffer_len == 0)
            stages.append(stage_len)
            stage_len = -(-stage_len // self.THREADS)
        stages.append(1)
        self.stages, self.num_stages = stages, len(stages)


class WgPrefixSum:
    """prefix_sum.rs:22-117: in-place exclusive prefix sum of a GpuVector<u32> (wrapping)."""

    THREADS = 256

    def __init__(self, device):
        self._device = device
        self.prefix_sum, self.add_data_grp = "prefix_sum", "add_data_grp"   # the reference's two pipelines, by name

    @staticmethod
    def from_device(device) -> "WgPrefixSum":
        lib()
        return WgPrefixSum(device)

    def dispatch(self, device, pass_, workspace: PrefixSumWorkspace, data) -> None:   # :49-99
        v = as_view(data, 1)
        if v.dtype != "u32":
            raise TypeError("WgPrefixSum.dispatch: data must be a GpuVector<u32>")
        workspace.reserve(device, v.view_shape.size[0])
        s = v.view_shape.to_c()
        check(lib().wgb_prefix_sum(pass_._h, v.buffer()._h, ctypes.byref(s)))

    @staticmethod
    def eval_cpu(v: np.ndarray) -> None:                           # :101-117
        """The reference's sequential CPU version, in place on a uint32 array."""
        if v.size == 0:
            return
        c = np.cumsum(v, dtype=np.uint32)                          # wraps modulo 2^32 like the u32 adds
        v[1:] = c[:-1]
        v[0] = 0


class RadixSortWorkspace:
    """radix_sort/mod.rs:82-109 (pass uniforms, count / reduced buffers, indirect-dispatch sizes, ping-pong outputs)."""

    def __init__(self, device=None):
        self._device = device

    @staticmethod
    def new(device) -> "RadixSortWorkspace":
        return RadixSortWorkspace(device)


class RadixSort:
    """radix_sort/mod.rs:67-223: stable LSD sort of (u32 key, u32 value) pairs."""

    def __init__(self, device):
        self._device = device

    @staticmethod
    def from_device(device) -> "RadixSort":
        lib()
        return RadixSort(device)

    def dispatch(self, device, pass_, workspace: RadixSortWorkspace, input_keys, input_values, n_sort, sorting_bits: int,
                 output_keys, output_values) -> None:              # :111-223
        ik, iv, ok, ov = (as_view(x, 1) for x in (input_keys, input_values, output_keys, output_values))
        for x in (ik, iv, ok, ov):
            if x.dtype != "u32":
                raise TypeError("RadixSort.dispatch: keys and values must be GpuVector<u32>")
        assert ik.view_shape.size[0] == iv.view_shape.size[0], \
            "Input keys and values must have the same number of elements"          # :121-125
        assert sorting_bits <= 32, "Can only sort up to 32 bits"                    # :126
        sk, sv, so, sw = (x.view_shape.to_c() for x in (ik, iv, ok, ov))
        check(lib().wgb_radix_sort(pass_._h, ik.buffer()._h, ctypes.byref(sk), iv.buffer()._h, ctypes.byref(sv),
                                   n_sort.buffer()._h, int(sorting_bits), ok.buffer()._h, ctypes.byref(so), ov.buffer()._h,
                                   ctypes.byref(sw)))
