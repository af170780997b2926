"""Generates the committed fixtures in tests/golden/.

The reference cannot be built or run here (Rust + wgpu, no toolchain) and holds no
golden vectors for this path, so the fixtures are: seeded inputs (oracle.uniform, seeds from
BASELINE.md §5) and outputs computed in float64 with numpy — an implementation independent of both
the oracle's C restatement and the CUDA kernels.  Both are checked against these files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402


def cm(flat, r, c):
    return flat.reshape(c, r).T.astype(np.float64)


def main():
    m1, m2 = O.uniform(O.SEED_BASE + 1, 64, 64), O.uniform(O.SEED_BASE + 2, 64, 64)
    a, b = cm(m1, 64, 64), cm(m2, 64, 64)
    np.savez(os.path.join(HERE, "cfg1_gemm64.npz"), m1=m1, m2=m2,
             gemm=(a @ b).T.reshape(-1), gemm_tr=(a.T @ b).T.reshape(-1))
    m, v, x128, x345 = O.uniform(O.SEED_BASE + 1, 128, 64), O.uniform(O.SEED_BASE + 3, 64), O.uniform(O.SEED_BASE + 3, 128), \
        O.uniform(O.SEED_BASE + 3, 345)
    mm, x = cm(m, 128, 64), x345.astype(np.float64)
    np.savez(os.path.join(HERE, "level12.npz"), m=m, v=v, x128=x128, x345=x345,
             gemv=mm @ v.astype(np.float64), gemv_tr=mm.T @ x128.astype(np.float64),
             min=x.min(), max=x.max(), sum=x.sum(), prod=x.prod(), sqnorm=(x * x).sum())
    geometry()


def geometry():
    """wgebra::geometry fixtures: seeded U[0,1) matrices a (general) and s = a^T a + 0.05 I (SDP), with float64 factorizations
    from numpy / scipy (nalgebra's conventions: lower Cholesky factor, QR with a non-negative diagonal of r)."""
    out = {}
    for dim in (2, 3, 4):
        a = O.uniform(O.SEED_BASE + 10 + dim, 64 * dim * dim).reshape(64, dim, dim)
        a64 = a.astype(np.float64)
        s = (np.einsum("nki,nkj->nij", a64, a64) + 0.05 * np.eye(dim)).astype(np.float32)
        s64 = s.astype(np.float64)
        q, r = np.linalg.qr(a64)
        sg = np.sign(np.diagonal(r, axis1=1, axis2=2))
        sg[sg == 0] = 1.0
        out.update({f"a{dim}": a, f"s{dim}": s, f"chol{dim}": np.linalg.cholesky(s64), f"q{dim}": q * sg[:, None, :],
                    f"r{dim}": r * sg[:, :, None], f"eigvals{dim}": np.linalg.eigvalsh(s64), f"inv{dim}": np.linalg.inv(s64),
                    f"det{dim}": np.linalg.det(a64), f"sv{dim}": np.linalg.svd(a64, compute_uv=False)})
    np.savez(os.path.join(HERE, "geometry.npz"), **out)


if __name__ == "__main__":
    main()
