"""Shared helpers for the parity tests: seeded inputs, the oracle, and the GPU call sequence
of the reference's own unit tests (gemm.rs:177-192 etc.)."""
import numpy as np

import wgmath_b200 as w
from oracle import oracle as O

U = w.BufferUsages
STORAGE = U.STORAGE | U.COPY_SRC | U.COPY_DST

SEED_A, SEED_B, SEED_V, SEED_OUT = O.SEED_BASE + 1, O.SEED_BASE + 2, O.SEED_BASE + 3, O.SEED_BASE + 4


def rel_err(got, ref):
    """max |got - ref| / |ref| elementwise (inputs are U[0,1), so |ref| is bounded away from 0)."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    denom = np.maximum(np.abs(ref), 1e-30)
    return float(np.max(np.abs(got - ref) / denom)) if ref.size else 0.0


def run_pass(gpu, fn):
    """encoder -> compute_pass -> fn(pass) -> drop(pass) -> submit (gemm.rs:177-192)."""
    enc = gpu.device().create_command_encoder()
    p = enc.compute_pass("test", None)
    try:
        fn(p)
    finally:
        p.end()
    gpu.queue().submit(enc.finish())


def upload(gpu, arr, shape, dtype=None):
    return w.TensorBuilder.tensor(shape, STORAGE).build_init(gpu.device(), arr, dtype)


def oshape(vs: w.ViewShape) -> O.Shape:
    return O.Shape(vs.size[0], vs.size[1], vs.size[2], vs.stride, vs.stride_mat, vs.offset)
