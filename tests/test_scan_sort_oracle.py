"""CPU tests (no GPU) for the scan / sort restatement in oracle/scan_sort_oracle.c: the reference's own unit tests
(wgrapier prefix_sum.rs:243-288, wgparry radix_sort/mod.rs:238-330) replayed against the oracle, plus edge cases.  Integer
results are unique, so every comparison is exact."""
import numpy as np
import pytest

import wgmath_b200 as w
from oracle import oracle as O


def seq_exclusive(v):
    out = np.zeros_like(v)
    if v.size > 1:
        out[1:] = np.cumsum(v[:-1].astype(np.uint64)).astype(np.uint64) & 0xFFFFFFFF
    return out


def reference_prefix_inputs(n=15071):
    """prefix_sum.rs:253-257: all ones, iota, random % 10000 (seeded here)."""
    rng = np.random.default_rng(15071)
    return [np.ones(n, np.uint32), np.arange(n, dtype=np.uint32), (rng.integers(0, 2 ** 32, n, dtype=np.uint64) % 10000).astype(np.uint32)]


def reference_sort_keys(i):
    """radix_sort/mod.rs:246-262."""
    return np.array([5 + i * 4, i, 6, 123, 74657, 123, 999, 2 ** 24 + 123, 6, 7, 8, 0, i * 2, 16 + i, 128 * i], dtype=np.uint32)


def stable_sorted(keys, vals, n_sort, bits):
    nb = 4 * ((bits + 3) // 4)
    mask = np.uint32(0xFFFFFFFF if nb >= 32 else (1 << nb) - 1)
    idx = np.argsort(keys[:n_sort] & mask, kind="stable")
    return keys[:n_sort][idx], vals[:n_sort][idx]


def test_reference_prefix_sum_replay():
    for v in reference_prefix_inputs():
        got = v.copy()
        assert O.prefix_sum(got) == O.ORC_OK
        cpu = v.copy()
        w.WgPrefixSum.eval_cpu(cpu)                      # the reference's eval_cpu (prefix_sum.rs:101-117), mirrored
        np.testing.assert_array_equal(got, cpu)
        np.testing.assert_array_equal(got, seq_exclusive(v))


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 511, 65536, 65537, 256 ** 2 * 3 + 11])
def test_prefix_sum_lengths_and_wraparound(n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)       # sums wrap modulo 2^32
    got = v.copy()
    assert O.prefix_sum(got) == O.ORC_OK
    np.testing.assert_array_equal(got, seq_exclusive(v))


def test_prefix_sum_workspace_levels():
    ws = w.PrefixSumWorkspace.with_capacity(None, 15071)
    assert ws.stages == [59, 1] and ws.num_stages == 2                        # prefix_sum.rs:185-206
    ws.reserve(None, 256 * 256 + 1)
    assert ws.stages == [257, 2, 1]
    ws.reserve(None, 200)
    assert ws.stages == [1]


def test_reference_radix_sort_replay():
    for i in range(128):
        keys = reference_sort_keys(i)
        vals = (keys * 2 + 5).astype(np.uint32)
        ok, ov = keys.copy(), vals.copy()                                     # the test initialises outputs with the inputs
        assert O.radix_sort(keys, vals, keys.size, 32, ok, ov) == O.ORC_OK
        rk, rv = stable_sorted(keys, vals, keys.size, 32)                     # cpu_argsort (mod.rs:232-236) is stable
        np.testing.assert_array_equal(ok, rk)
        np.testing.assert_array_equal(ov, rv)


@pytest.mark.parametrize("bits", [0, 1, 4, 7, 8, 12, 20, 29, 32])
@pytest.mark.parametrize("n,n_sort", [(15, 15), (1024, 1024), (1025, 1000), (70001, 70001), (5000, 0)])
def test_radix_sort_bits_counts_and_untouched_tail(bits, n, n_sort):
    rng = np.random.default_rng(bits * 1000 + n)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys[::7] = keys[0]                                                       # duplicates: stability is observable
    vals = np.arange(n, dtype=np.uint32)
    ok, ov = np.full(n, 0xDEADBEEF, np.uint32), np.full(n, 0xFEEDFACE, np.uint32)
    k0, v0 = keys.copy(), vals.copy()
    assert O.radix_sort(keys, vals, n_sort, bits, ok, ov) == O.ORC_OK
    np.testing.assert_array_equal(keys, k0)
    np.testing.assert_array_equal(vals, v0)
    if bits == 0:
        assert (ok == 0xDEADBEEF).all() and (ov == 0xFEEDFACE).all()         # zero passes: nothing is written (mod.rs:156)
        return
    rk, rv = stable_sorted(keys, vals, n_sort, bits)
    np.testing.assert_array_equal(ok[:n_sort], rk)
    np.testing.assert_array_equal(ov[:n_sort], rv)
    assert (ok[n_sort:] == 0xDEADBEEF).all() and (ov[n_sort:] == 0xFEEDFACE).all()


def test_radix_sort_rejects_more_than_32_bits():
    k = np.zeros(4, np.uint32)
    assert O.radix_sort(k, k.copy(), 4, 33, k.copy(), k.copy()) != O.ORC_OK
