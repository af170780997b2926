"""wgmath_b200 — B200-native (sm_100a CUDA) backend for the wgebra dense-linear-algebra surface.

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of
include/wgb200.h, built into libwgebra_b200.so) and this host-side mirror of the reference's
Rust interface (wgcore::{gpu,tensor,shapes,kernel,timestamps} + wgebra::linalg)."""
from ._lib import DimensionMismatch, WgbError, lib  # noqa: F401
from .gpu import CommandEncoder, ComputePass, Device, GpuInstance, Graph, Queue  # noqa: F401
from .linalg import (Dot, F32Mode, Gemm, GemmVariant, Gemv, GemvVariant, OpAssign, OpAssignVariant, Reduce,  # noqa: F401
                     ReduceOp, fill_uniform)
from .geometry import (GpuLU, GpuQR, GpuSvd, GpuSymmetricEigen, Matrix, WgCholesky2, WgCholesky3, WgCholesky4, WgInv, WgLU2,  # noqa: F401
                       WgLU3, WgLU4, WgQR2, WgQR3, WgQR4, WgSvd2, WgSvd3, WgSymmetricEigen2, WgSymmetricEigen3,
                       WgSymmetricEigen4)
from .primitives import PrefixSumWorkspace, RadixSort, RadixSortWorkspace, WgPrefixSum  # noqa: F401
from .shapes import ViewShape, ViewShapeBuffers  # noqa: F401
from .tensor import (BufferUsages, ColumnMajor, RowMajor, GpuCube, GpuCubeView, GpuMatrix, GpuMatrixView, GpuScalar, GpuTensor,  # noqa: F401
                     GpuTensorView, GpuVector, GpuVectorView, TensorBuilder, as_view)
from .timestamps import GpuTimestamps  # noqa: F401
