"""The cases behind tests/golden/ref_wgsl_*.npz: which reference shader, which entry point, which views over which buffers.

Shared by the generator (make_reference_vectors.py: runs the reference's WGSL through tests/golden/wgsl_interp.py and stores
the output buffers) and by the tests (which rebuild the same inputs, run the oracle / the CUDA path on them and compare with
the stored buffers).  Pure data: nothing here reads /root/reference.

A view is the six u32 of wgcore::shapes::ViewShape (nrows, ncols, nmats, stride, stride_mat, offset).  Buffers are flat f32
arrays filled from the seeded generator (oracle.uniform) unless stated otherwise; every buffer is long enough for the vec4
accesses of the shaders (they read and write whole 4-row blocks, with_vec4_elts in shape.wgsl).
"""
from __future__ import annotations

import zlib

import numpy as np

GEMM_VARIANTS = {"gemm": 0, "gemm_fast": 1, "gemm_tr": 2, "gemm_tr_fast": 3}      # oracle / wgb200.h numbering
GEMV_VARIANTS = {"gemv": 0, "gemv_fast": 1, "gemv_tr": 2, "gemv_tr_fast": 3}
OP_ASSIGN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "copy": 4}
REDUCE = {"min": 0, "max": 1, "sum": 2, "prod": 3, "sqnorm": 4}


def dense(nrows, ncols=1, nmats=1):
    return (nrows, ncols, nmats, nrows, nrows * ncols, 0)


def _gemm_case(name, variant, so, s1, s2, len_out, len_1, len_2, zero_k_padding=None):
    return dict(kind="gemm", name=f"{name}/{variant}", variant=variant, so=so, s1=s1, s2=s2, len_out=len_out, len_1=len_1,
                len_2=len_2, zero_k_padding=zero_k_padding)


def gemm_cases():
    cases = []
    # BASELINE configs[0]: 64^3 dense
    for v in ("gemm", "gemm_tr"):
        cases.append(_gemm_case("cfg0_64", v, dense(64, 64), dense(64, 64), dense(64, 64), 64 * 64, 64 * 64, 64 * 64))
    # the reference's own test size (gemm.rs:144-202: 256 x 256 operands, all four variants)
    for v in GEMM_VARIANTS:
        cases.append(_gemm_case("ref_test_256", v, dense(256, 256), dense(256, 256), dense(256, 256), 256 * 256, 256 * 256, 256 * 256))
    # sub-views of batched parents: rows 8..40 of 64 x 48 matrices times columns 4..24 of 48 x 40 matrices, padded output
    M, K, N, T = 32, 48, 20, 3
    cases.append(_gemm_case("views", "gemm", (M, N, T, 36, 36 * 24, 4), (M, K, T, 64, 64 * 48, 8), (K, N, T, 48, 48 * 40, 48 * 4),
                            4 + 36 * 24 * T, 64 * 48 * T, 48 * 40 * T))
    # the transposed variant reads a K x M view: rows 8..56 of the 64-row parents, 32 of their 48 columns starting at column 4
    cases.append(_gemm_case("views", "gemm_tr", (M, N, T, 36, 36 * 24, 4), (K, M, T, 64, 64 * 48, 8 + 4 * 64), (K, N, T, 48, 48 * 40, 48 * 4),
                            4 + 36 * 24 * T, 64 * 48 * T, 48 * 40 * T))
    # the workgroup variants: K = 512 is two rounds of 64 invocations x 4 columns (gemm_fast) / eight rounds of 64 rows (gemm_tr_fast)
    M, K, N, T = 8, 512, 8, 2
    cases.append(_gemm_case("fast_views", "gemm_fast", (M, N, T, 12, 12 * 8, 4), (M, K, T, 16, 16 * K, 4), (K, N, T, K, K * N, 8),
                            4 + 12 * 8 * T, 4 + 16 * K * T, 8 + K * N * T))
    cases.append(_gemm_case("fast_views", "gemm_tr_fast", (M, N, T, 12, 12 * 8, 4), (K, M, T, K + 8, (K + 8) * M, 4), (K, N, T, K, K * N, 8),
                            4 + 12 * 8 * T, 4 + (K + 8) * M * T, 8 + K * N * T))
    # dimensions that are not multiples of 4: the shaders work on whole 4 x 4 blocks, so the buffers are padded to 32 x 48,
    # 48 x 20 and 32 x 20 and the K padding (columns 46..47 of m1, rows 46..47 of m2) is zero, as a caller of the reference
    # must arrange for the product to be A * B
    M, K, N = 30, 46, 18
    cases.append(_gemm_case("ragged", "gemm", (M, N, 1, 32, 32 * 20, 0), (M, K, 1, 32, 32 * 48, 0), (K, N, 1, 48, 48 * 20, 0),
                            32 * 20, 32 * 48, 48 * 20, zero_k_padding=("cols", "rows")))
    cases.append(_gemm_case("ragged", "gemm_tr", (M, N, 1, 32, 32 * 20, 0), (K, M, 1, 48, 48 * 32, 0), (K, N, 1, 48, 48 * 20, 0),
                            32 * 20, 48 * 32, 48 * 20, zero_k_padding=("rows", "rows")))
    return cases


def _gemv_case(name, variant, so, sm, sv, len_out, len_m, len_v):
    return dict(kind="gemv", name=f"{name}/{variant}", variant=variant, so=so, sm=sm, sv=sv, len_out=len_out, len_m=len_m, len_v=len_v)


def gemv_cases():
    cases = []
    n = 1024                                                     # gemv.rs:153-197
    for v in GEMV_VARIANTS:
        cases.append(_gemv_case("ref_test_1024", v, dense(n), dense(n, n), dense(n), n, n * n, n))
    # several right-hand sides (out.ncols = 2) and a batch of 2, as sub-views with offsets
    M, K, C, T = 24, 40, 2, 2
    cases.append(_gemv_case("views", "gemv", (M, C, T, 28, 28 * C, 4), (M, K, T, 32, 32 * K, 8), (K, C, T, 44, 44 * C, 4),
                            4 + 28 * C * T, 8 + 32 * K * T, 4 + 44 * C * T))
    cases.append(_gemv_case("views", "gemv_tr", (M, C, T, 28, 28 * C, 4), (K, M, T, 44, 44 * M, 4), (K, C, T, 44, 44 * C, 4),
                            4 + 28 * C * T, 4 + 44 * M * T, 4 + 44 * C * T))
    M, K = 8, 256                                                # two rounds of 32 invocations x 4 columns / eight rounds of 32 rows
    cases.append(_gemv_case("fast_views", "gemv_fast", (M, C, T, 12, 12 * C, 4), (M, K, T, 16, 16 * K, 4), (K, C, T, K + 4, (K + 4) * C, 4),
                            4 + 12 * C * T, 4 + 16 * K * T, 4 + (K + 4) * C * T))
    cases.append(_gemv_case("fast_views", "gemv_tr_fast", (M, C, T, 12, 12 * C, 4), (K, M, T, K + 4, (K + 4) * M, 4), (K, C, T, K + 4, (K + 4) * C, 4),
                            4 + 12 * C * T, 4 + (K + 4) * M * T, 4 + (K + 4) * C * T))
    # dimensions that are not multiples of 4 (buffers padded to whole vec4 blocks, K padding zero, as for the ragged gemm cases)
    M, K = 30, 46
    cases.append(_gemv_case("ragged", "gemv", (M, 1, 1, 32, 32, 0), (M, K, 1, 32, 32 * 48, 0), (K, 1, 1, 48, 48, 0), 32, 32 * 48, 48) | dict(zero_k_padding=("cols", "rows")))
    cases.append(_gemv_case("ragged", "gemv_tr", (M, 1, 1, 32, 32, 0), (K, M, 1, 48, 48 * 32, 0), (K, 1, 1, 48, 48, 0), 32, 48 * 32, 48) | dict(zero_k_padding=("rows", "rows")))
    return cases


def op_assign_cases():
    cases = []
    for op in OP_ASSIGN:
        # op_assign.rs:123-129: LEN = 1757, v0[i] = i + 0.1, v1[i] = 10 i + 0.1 (the one deterministic fixture of the reference)
        cases.append(dict(kind="op_assign", name=f"ref_test_1757/{op}", op=op, sa=dense(1757), sb=dense(1757), len_a=1757, len_b=1757,
                          fixture="op_assign.rs"))
        cases.append(dict(kind="op_assign", name=f"views/{op}", op=op, sa=(100, 1, 1, 100, 100, 3), sb=(100, 1, 1, 100, 100, 5),
                          len_a=110, len_b=120, fixture=None))
    return cases


def reduce_cases():
    cases = []
    for op in REDUCE:
        for n, off in ((345, 0), (0, 0), (1, 0), (127, 0), (128, 0), (129, 0), (1000, 7)):   # 345: reduce.rs:139-179
            cases.append(dict(kind="reduce", name=f"n{n}_off{off}/{op}", op=op, s=(n, 1, 1, n, n, off), len=max(n + off, 4)))
    return cases


def all_linalg_cases():
    return gemm_cases() + gemv_cases() + op_assign_cases() + reduce_cases()


# ------------------------------------------------------------------------------------------------ input buffers
def _uniform(seed, n):
    from oracle import oracle as O
    return O.uniform(0x5EED0000 + seed, n).copy()


def _zero_k(buf, shape, what, K):
    """zero the K padding of a column-major view: columns K..ceil4(K) ('cols') or rows K..ceil4(K) ('rows')"""
    nrows, ncols, nmats, stride, stride_mat, offset = shape
    K4 = (K + 3) // 4 * 4
    for t in range(nmats):
        base = offset + t * stride_mat
        if what == "cols":
            for j in range(K, K4):
                buf[base + j * stride: base + j * stride + (nrows + 3) // 4 * 4] = 0.0
        else:
            for j in range((ncols + 3) // 4 * 4):
                buf[base + j * stride + K: base + j * stride + K4] = 0.0


def inputs(case):
    """{buffer name: float32 array} for a case, output buffers included (pre-filled: the shaders overwrite, never accumulate)"""
    k = case["kind"]
    h = zlib.crc32(case["name"].encode()) & 0xFFFF
    if k == "gemm":
        m1, m2 = _uniform(h + 1, case["len_1"]), _uniform(h + 2, case["len_2"])
        if case["zero_k_padding"]:
            K = case["s2"][0]
            _zero_k(m1, case["s1"], case["zero_k_padding"][0], K)
            _zero_k(m2, case["s2"], case["zero_k_padding"][1], K)
        return dict(out=np.full(case["len_out"], -7.0, np.float32), m1=m1, m2=m2)
    if k == "gemv":
        m, v = _uniform(h + 1, case["len_m"]), _uniform(h + 3, case["len_v"])
        if case.get("zero_k_padding"):
            K = case["sv"][0]
            _zero_k(m, case["sm"], case["zero_k_padding"][0], K)
            _zero_k(v, case["sv"], case["zero_k_padding"][1], K)
        return dict(out=_uniform(h + 4, case["len_out"]), m=m, v=v)
    if k == "op_assign":
        if case["fixture"]:
            i = np.arange(case["len_a"], dtype=np.float32)
            return dict(a=i + np.float32(0.1), b=i * np.float32(10.0) + np.float32(0.1))
        return dict(a=_uniform(h + 1, case["len_a"]) + np.float32(0.5), b=_uniform(h + 2, case["len_b"]) + np.float32(0.5))
    if k == "reduce":
        return dict(x=_uniform(h + 1, case["len"]) + np.float32(0.5), out=np.full(1, -7.0, np.float32))
    raise KeyError(k)


def shape_bytes(shape):
    return np.array(shape, np.uint32).view(np.uint8).copy()


# ------------------------------------------------------------------------------------------------ wgebra::geometry
# out[i] = f(in[i]) over a batch, as the test kernels embedded in the reference's Rust tests do (cholesky.rs:58-70 ...).
# Output struct sizes in 4-byte words, WGSL storage layout (oracle/geometry_oracle.c header).
GEOMETRY_OUT_WORDS = {("cholesky", 2): 4, ("cholesky", 3): 12, ("cholesky", 4): 16, ("lu", 2): 10, ("lu", 3): 20, ("lu", 4): 28,
                      ("qr", 2): 8, ("qr", 3): 24, ("qr", 4): 32, ("eig", 2): 6, ("eig", 3): 16, ("eig", 4): 20,
                      ("svd", 2): 10, ("svd", 3): 28, ("inv", 2): 4, ("inv", 3): 12, ("inv", 4): 16}
GEOMETRY_OPS = {"cholesky": 0, "lu": 1, "qr": 2, "eig": 3, "svd": 4, "inv": 5}      # oracle / wgb200.h numbering
GEOMETRY_BATCH = 48


def geometry_cases():
    return sorted(GEOMETRY_OUT_WORDS)


def geometry_inputs(op, dim, n=GEOMETRY_BATCH):
    """[n, dim, dim] float32: U[0,1) matrices; symmetric positive definite (a^T a + 0.05 I, rounded to f32 and symmetrised) for
    cholesky and eig.  A few special matrices lead the batch: identity, diagonal, a matrix needing a row swap in LU."""
    a = _uniform(100 + dim, n * dim * dim).reshape(n, dim, dim)
    if op in ("cholesky", "eig"):
        a64 = a.astype(np.float64)
        s = np.einsum("nki,nkj->nij", a64, a64) + 0.05 * np.eye(dim)
        a = ((s + np.transpose(s, (0, 2, 1))) * 0.5).astype(np.float32)
    a[0] = np.eye(dim, dtype=np.float32)
    a[1] = np.diag(np.arange(1, dim + 1).astype(np.float32))
    if op not in ("cholesky", "eig"):
        a[2, 0, 0] = 0.0                                          # zero pivot: partial pivoting must swap
    return a


# ------------------------------------------------------------------------------------------------ prefix sum / radix sort (u32)
def _u32(seed, n):
    """32 random bits per element from two 24-bit draws of the seeded generator"""
    a = (_uniform(seed, n).astype(np.float64) * (1 << 24)).astype(np.uint32)
    b = (_uniform(seed + 1, n).astype(np.float64) * (1 << 24)).astype(np.uint32)
    return (a << np.uint32(8)) ^ b


def scan_cases():
    """prefix_sum.rs:243-288: LEN = 15071, inputs all ones / iota / random % 10000; plus a length that needs three levels of
    block totals (66000 -> 258 -> 2 -> 1) and the degenerate lengths"""
    return [dict(name="ones_15071", n=15071, fill="ones"), dict(name="iota_15071", n=15071, fill="iota"),
            dict(name="random_15071", n=15071, fill="random"), dict(name="random_66000", n=66000, fill="random"),
            dict(name="random_256", n=256, fill="random"), dict(name="random_1", n=1, fill="random"), dict(name="wrapping_700", n=700, fill="large")]


def scan_input(case):
    n = case["n"]
    if case["fill"] == "ones":
        return np.ones(n, np.uint32)
    if case["fill"] == "iota":
        return np.arange(n, dtype=np.uint32)
    if case["fill"] == "large":
        return _u32(300 + n, n)                                   # sums wrap modulo 2^32
    return _u32(200 + n, n) % np.uint32(10000)


def sort_cases():
    """radix_sort/mod.rs:238-330 sorts 15 keys (values = 2 key + 5) at full width; here also several workgroups, a pair count
    below the buffer length, odd and even pass counts and keys with bits above the sorted ones"""
    return [dict(name="n15_bits32", len=15, n_sort=15, bits=32), dict(name="n1500_bits32", len=1500, n_sort=1500, bits=32),
            dict(name="n1000_of_1500_bits12", len=1500, n_sort=1000, bits=12), dict(name="n2500_bits16", len=2500, n_sort=2500, bits=16),
            dict(name="n300_bits6_duplicates", len=300, n_sort=300, bits=6)]


def sort_input(case):
    keys = _u32(400 + case["len"] + case["bits"], case["len"])
    if case["name"].endswith("duplicates"):
        keys = keys % np.uint32(7) + (keys & np.uint32(0xFFFF0000))      # few distinct low digits, noise above the sorted bits
    values = keys * np.uint32(2) + np.uint32(5)                           # mod.rs:267
    return keys, values


# ------------------------------------------------------------------------------------------------ shape.wgsl index functions
SHAPE_VIEWS = [(10, 7, 3, 12, 100, 5), (64, 64, 1, 64, 4096, 0), (30, 18, 2, 32, 640, 8), (1, 1, 1, 1, 1, 3), (7, 5, 4, 9, 63, 2),
               (4096, 4096, 1, 4096, 1 << 24, 0)]
SHAPE_QUERIES = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (3, 4, 2), (6, 4, 3), (9, 6, 1), (29, 17, 1), (4095, 4095, 0)]   # (i, j, t)
