"""GPU parity tests (-m gpu) for the integer primitives of SURVEY.md §8(f) 4 — wgb_prefix_sum and wgb_radix_sort through the
reference-shaped host mirror (WgPrefixSum, RadixSort) — against the CPU oracle.  Everything is bit-exact.  Full sizes (2^26) go
through size-independent properties: last element + total, sortedness, stability, permutation checksum."""
import numpy as np
import pytest

import wgmath_b200 as w
from oracle import oracle as O
from tests.helpers import STORAGE, run_pass
from tests.test_scan_sort_oracle import reference_prefix_inputs, reference_sort_keys, seq_exclusive, stable_sorted

pytestmark = pytest.mark.gpu


def vec_u32(gpu, arr):
    return w.TensorBuilder.vector(arr.size, STORAGE).build_init(gpu.device(), arr, "u32")


def test_gpu_prefix_sum_reference_replay(gpu):
    """prefix_sum.rs:243-288: LEN = 15071, three inputs, exact equality with the CPU scan."""
    ps = w.WgPrefixSum.from_device(gpu.device())
    for v in reference_prefix_inputs():
        t = vec_u32(gpu, v)
        ws = w.PrefixSumWorkspace.with_capacity(gpu.device(), v.size)
        run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, ws, t))
        ref = v.copy()
        assert O.prefix_sum(ref) == O.ORC_OK
        np.testing.assert_array_equal(t.read(), ref)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 255, 256, 257, 4095, 4096, 4097, 8191, 65537, 1_000_003])
def test_gpu_prefix_sum_lengths_and_wraparound(gpu, n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    t = vec_u32(gpu, v)
    ps = w.WgPrefixSum.from_device(gpu.device())
    run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, w.PrefixSumWorkspace.new(), t))
    ref = v.copy()
    O.prefix_sum(ref)
    np.testing.assert_array_equal(t.read(), ref)


def test_gpu_prefix_sum_sub_view_leaves_the_rest_untouched(gpu):
    n, first, cnt = 20000, 4099, 9001                       # unaligned start: scalar path of the kernel
    v = np.arange(n, dtype=np.uint32) * np.uint32(2654435761)
    t = vec_u32(gpu, v)
    ps = w.WgPrefixSum.from_device(gpu.device())
    run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, w.PrefixSumWorkspace.new(), t.rows(first, cnt)))
    ref = v.copy()
    ref[first:first + cnt] = seq_exclusive(v[first:first + cnt])
    np.testing.assert_array_equal(t.read(), ref)


def test_gpu_prefix_sum_empty_is_a_noop(gpu):
    t = vec_u32(gpu, np.arange(8, dtype=np.uint32))
    ps = w.WgPrefixSum.from_device(gpu.device())
    run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, w.PrefixSumWorkspace.new(), t.rows(3, 0)))
    np.testing.assert_array_equal(t.read(), np.arange(8, dtype=np.uint32))


def test_gpu_prefix_sum_full_size_properties(gpu):
    """n = 2^26 (the bench size): ones -> iota; random -> spot values and the last element against a CPU cumsum."""
    n = 1 << 26
    ps = w.WgPrefixSum.from_device(gpu.device())
    t = vec_u32(gpu, np.ones(n, np.uint32))
    run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, w.PrefixSumWorkspace.new(), t))
    got = t.read()
    assert got[0] == 0 and got[-1] == n - 1 and (np.diff(got[:: 4093].astype(np.int64)) == 4093).all()
    rng = np.random.default_rng(26)
    v = rng.integers(0, 1000, n, dtype=np.uint32)
    t2 = vec_u32(gpu, v)
    run_pass(gpu, lambda p: ps.dispatch(gpu.device(), p, w.PrefixSumWorkspace.new(), t2))
    np.testing.assert_array_equal(t2.read(), seq_exclusive(v))


def sort_on_gpu(gpu, keys, vals, n_sort, bits, out_fill=(0xDEADBEEF, 0xFEEDFACE), out_len=None):
    dev = gpu.device()
    out_len = keys.size if out_len is None else out_len
    tk, tv = vec_u32(gpu, keys), vec_u32(gpu, vals)
    ok, ov = vec_u32(gpu, np.full(out_len, out_fill[0], np.uint32)), vec_u32(gpu, np.full(out_len, out_fill[1], np.uint32))
    ns = w.TensorBuilder.scalar(STORAGE).build_init(dev, np.array([n_sort], np.uint32), "u32")
    sort = w.RadixSort.from_device(dev)
    ws = w.RadixSortWorkspace.new(dev)
    run_pass(gpu, lambda p: sort.dispatch(dev, p, ws, tk, tv, ns, bits, ok, ov))
    np.testing.assert_array_equal(tk.read(), keys)            # inputs are never modified
    np.testing.assert_array_equal(tv.read(), vals)
    return ok.read(), ov.read()


def test_gpu_radix_sort_reference_replay(gpu):
    """radix_sort/mod.rs:238-330 test_sorting: 15 keys x 128 variations, values = 2 * key + 5, 32 bits."""
    for i in range(0, 128, 3):
        keys = reference_sort_keys(i)
        vals = (keys * 2 + 5).astype(np.uint32)
        gk, gv = sort_on_gpu(gpu, keys, vals, keys.size, 32)
        rk, rv = keys.copy(), vals.copy()
        assert O.radix_sort(keys, vals, keys.size, 32, rk, rv) == O.ORC_OK
        np.testing.assert_array_equal(gk, rk)
        np.testing.assert_array_equal(gv, rv)


@pytest.mark.parametrize("bits", [0, 1, 4, 8, 12, 20, 24, 29, 32])
@pytest.mark.parametrize("n,n_sort", [(15, 15), (4096, 4096), (4097, 4000), (70001, 70001), (5000, 0), (300_000, 299_999)])
def test_gpu_radix_sort_matches_the_oracle(gpu, bits, n, n_sort):
    rng = np.random.default_rng(bits * 1000 + n)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys[::7] = keys[0]                                       # duplicates make stability observable through the values
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = sort_on_gpu(gpu, keys, vals, n_sort, bits)
    rk, rv = np.full(n, 0xDEADBEEF, np.uint32), np.full(n, 0xFEEDFACE, np.uint32)
    assert O.radix_sort(keys, vals, n_sort, bits, rk, rv) == O.ORC_OK
    np.testing.assert_array_equal(gk, rk)                     # includes the untouched tail past n_sort
    np.testing.assert_array_equal(gv, rv)


def test_gpu_radix_sort_few_distinct_keys_and_n_sort_beyond_length(gpu):
    n = 50_000
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 3, n, dtype=np.uint32) * np.uint32(0x01010101)     # every digit sees only 3 populated bins
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = sort_on_gpu(gpu, keys, vals, n + 1000, 32)       # *n_sort > length: clamped to the vector length
    rk, rv = stable_sorted(keys, vals, n, 32)
    np.testing.assert_array_equal(gk, rk)
    np.testing.assert_array_equal(gv, rv)


def test_gpu_radix_sort_dimension_mismatch_and_bits(gpu):
    dev = gpu.device()
    a, b = vec_u32(gpu, np.zeros(8, np.uint32)), vec_u32(gpu, np.zeros(9, np.uint32))
    ns = w.TensorBuilder.scalar(STORAGE).build_init(dev, np.array([8], np.uint32), "u32")
    sort = w.RadixSort.from_device(dev)
    with pytest.raises(AssertionError):                        # the reference's assert_eq! (mod.rs:121-125)
        run_pass(gpu, lambda p: sort.dispatch(dev, p, w.RadixSortWorkspace.new(dev), a, b, ns, 32, a, b))
    with pytest.raises(AssertionError):                        # assert!(sorting_bits <= 32) (mod.rs:126)
        run_pass(gpu, lambda p: sort.dispatch(dev, p, w.RadixSortWorkspace.new(dev), a, a, ns, 33, a, a))
    o = vec_u32(gpu, np.zeros(8, np.uint32))
    with pytest.raises(w.WgbError):                            # outputs overlapping inputs are rejected at the C ABI
        run_pass(gpu, lambda p: sort.dispatch(dev, p, w.RadixSortWorkspace.new(dev), a, o, ns, 32, a, o))


def test_gpu_radix_sort_full_size_properties(gpu):
    """n = 2^26 pairs (the bench size): sorted, stable (values increase inside runs of equal keys), and a permutation
    (values are exactly 0..n-1; keys follow their values)."""
    n = 1 << 26
    rng = np.random.default_rng(2026)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys[: n // 8] &= np.uint32(0xFFFF)                       # many duplicates
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = sort_on_gpu(gpu, keys, vals, n, 32)
    d = np.diff(gk.astype(np.int64))
    assert (d >= 0).all()
    eq = d == 0
    assert (np.diff(gv.astype(np.int64))[eq] > 0).all()
    assert (np.bincount(gv, minlength=n) == 1).all()
    np.testing.assert_array_equal(keys[gv], gk)
