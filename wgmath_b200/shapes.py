"""ViewShape / ViewShapeBuffers — mirrors /root/reference/crates/wgcore/src/shapes.rs:9-116."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Tuple

from ._lib import ViewShapeC


@dataclass(frozen=True)
class ViewShape:
    """shapes.rs:9-21: `size` = [rows, cols, mats]; strides and offset in elements (u32)."""
    size: Tuple[int, int, int]
    stride: int
    stride_mat: int
    offset: int

    def to_c(self) -> ViewShapeC:
        for v in (*self.size, self.stride, self.stride_mat, self.offset):
            if not 0 <= v <= 0xFFFFFFFF:
                raise OverflowError("ViewShape fields are u32 (shapes.rs:12-21)")
        return ViewShapeC((ctypes.c_uint32 * 3)(*self.size), self.stride, self.stride_mat, self.offset)

    def f32_to_vec4(self, column_major: bool = True) -> "ViewShape":
        """shapes.rs:25-39 (kept for source compatibility; the CUDA kernels do not need it)."""
        size = (self.size[0] // 4, self.size[1], self.size[2]) if column_major else (self.size[0], self.size[1] // 4, self.size[2])
        return ViewShape(size, self.stride // 4, self.stride_mat // 4, self.offset // 4)



class ViewShapeBuffers:
    """shapes.rs:46-116.  In the reference this caches one 24-byte uniform buffer per distinct shape
    ("emulated push constants").  CUDA passes the shape by value as a kernel parameter, so the cache is a
    shim kept for source compatibility: `get` returns the shape itself."""

    def __init__(self):
        self._seen = set()

    @staticmethod
    def new() -> "ViewShapeBuffers":
        return ViewShapeBuffers()

    def clear_tmp(self) -> None:
        pass

    def put_tmp(self, device, queue, shape: ViewShape) -> None:
        self._seen.add(shape)

    def contains(self, shape: ViewShape) -> bool:
        return shape in self._seen

    def get(self, device, shape: ViewShape) -> ViewShape:
        self._seen.add(shape)
        return shape
