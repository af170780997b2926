"""GpuTimestamps — mirrors /root/reference/crates/wgcore/src/timestamps.rs:9-248 with CUDA events."""
from __future__ import annotations

import ctypes
from typing import List, Tuple

from ._lib import check, lib


class GpuTimestamps:
    def __init__(self, device, capacity: int = 64):
        self._device = device
        self._events: List[ctypes.c_void_p] = []
        for _ in range(capacity):
            h = ctypes.c_void_p()
            check(lib().wgb_event_create(device._h, ctypes.byref(h)))
            self._events.append(h)
        self._len = 0

    @staticmethod
    def new(device, capacity: int = 64) -> "GpuTimestamps":
        return GpuTimestamps(device, capacity)

    def __del__(self):
        for h in getattr(self, "_events", []):
            try:
                lib().wgb_event_destroy(h)
            except Exception:
                pass

    def clear(self) -> None:
        self._len = 0

    def len(self) -> int:
        return self._len

    def next_compute_pass_timestamp_writes(self) -> Tuple[ctypes.c_void_p, ctypes.c_void_p]:
        """timestamps.rs:63-70: reserve a begin and an end slot for one compute pass."""
        if self._len + 2 > len(self._events):
            raise IndexError("GpuTimestamps capacity exceeded")
        b, e = self._events[self._len], self._events[self._len + 1]
        self._len += 2
        return b, e

    def resolve(self, encoder) -> None:
        """timestamps.rs:119-134: nothing to copy with CUDA events."""

    def wait_for_results_ms(self, device=None, queue=None) -> List[float]:
        """timestamps.rs:226-230: per-slot times in ms relative to the first slot."""
        out = [0.0] * self._len
        ms = ctypes.c_float()
        for i in range(1, self._len):
            check(lib().wgb_event_elapsed_ms(self._events[0], self._events[i], ctypes.byref(ms)))
            out[i] = ms.value
        return out
