/*
 * scan_sort_oracle.c — CPU restatement of the two integer primitives SURVEY.md §8(f) lists next to the linalg path:
 * the exclusive prefix sum of wgrapier and the key/value radix sort of wgparry.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as wgsl_oracle.c: only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load it).
 *
 * What it restates (paths relative to /root/reference/crates/):
 *   wgrapier/src/dynamics/prefix_sum.wgsl:35-128   prefix_sum: one 256-element Blelloch up-sweep / down-sweep per workgroup,
 *                                                  block total to aux[bid]
 *   wgrapier/src/dynamics/prefix_sum.wgsl:139-147  add_data_grp: data[i] += aux[block of i]
 *   wgrapier/src/dynamics/prefix_sum.rs:49-99      dispatch: scan data, scan the aux levels, add back from the coarsest level
 *   wgrapier/src/dynamics/prefix_sum.rs:185-224    PrefixSumWorkspace::reserve: level lengths ceil(n / 256) ... 1
 *   wgparry/src/utils/radix_sort/mod.rs:111-223    dispatch: ceil(sorting_bits / 4) passes of 4 bits, ping-pong buffers,
 *                                                  the last pass lands in output_keys / output_values
 *   wgparry/src/utils/radix_sort/sort_count.wgsl   per-workgroup (1024 keys) digit histogram, counts[bin * num_wgs + wg]
 *   .../sort_reduce.wgsl, sort_scan.wgsl, sort_scan_add.wgsl   together: exclusive scan of counts[] in (bin, wg) order
 *   .../sort_scatter.wgsl                          stable placement of each 256-key sub-block at its bin offsets
 *
 * All arithmetic is u32 (wrapping).  PARITY PINNING: the reference's own tests compare with a sequential CPU scan
 * (prefix_sum.rs:101-117 eval_cpu; LEN = 15071, inputs all-ones / iota / random % 10000, prefix_sum.rs:243-288) and with a
 * stable CPU argsort (radix_sort/mod.rs:238-330: 15 keys x 128 variations, values = 2 * key + 5).  tests/test_oracle.py replays
 * both against this file; integer results are unique, so parity is bit-exact.  In addition the reference's own shaders
 * (prefix_sum.wgsl; sorting.wgsl + the six sort_*.wgsl, with the host sequences of prefix_sum.rs:47-99 and mod.rs:204-322) are
 * executed by tests/golden/wgsl_interp.py and their outputs (tests/golden/ref_wgsl_scan_sort.npz) must equal this file's
 * (tests/test_reference_vectors.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_OK = 0, ORC_DIM_MISMATCH = 2, ORC_UNSUPPORTED = 4 };

#define PS_WG 256u

static uint32_t next_power_of_two(uint32_t v) { /* prefix_sum.wgsl:152-163 */
    v--;
    v |= v >> 1;
    v |= v >> 2;
    v |= v >> 4;
    v |= v >> 8;
    v |= v >> 16;
    v++;
    return v;
}

/* prefix_sum.wgsl:35-128: one workgroup.  Threads become loops over tid between the barriers. */
static void k_prefix_sum_wg(uint32_t *data, uint32_t data_len, uint32_t *aux, uint32_t bid) {
    uint32_t workspace[PS_WG];
    if (bid * PS_WG >= data_len) return;
    const uint32_t data_block_len = data_len - bid * PS_WG;
    uint32_t shared_len = next_power_of_two(data_block_len);
    if (shared_len < 1u) shared_len = 1u;
    if (shared_len > PS_WG) shared_len = PS_WG;
    for (uint32_t tid = 0; tid < PS_WG; ++tid) {
        const uint32_t elt = tid + bid * PS_WG;
        workspace[tid] = elt < data_len ? data[elt] : 0u;
    }
    for (uint32_t d = shared_len / 2, offset = 1; d > 0; d /= 2, offset *= 2) /* up-sweep :62-82 */
        for (uint32_t tid = 0; tid < d; ++tid) {
            const uint32_t ia = tid * 2u * offset + offset - 1u, ib = (tid * 2u + 1u) * offset + offset - 1u;
            workspace[ib] = workspace[ia] + workspace[ib];
        }
    aux[bid] = workspace[shared_len - 1]; /* :85-89 */
    workspace[shared_len - 1] = 0u;
    for (uint32_t d = 1, offset = shared_len / 2; d < shared_len; d *= 2, offset /= 2) /* down-sweep :93-116 */
        for (uint32_t tid = 0; tid < d; ++tid) {
            const uint32_t ia = tid * 2u * offset + offset - 1u, ib = (tid * 2u + 1u) * offset + offset - 1u;
            const uint32_t a = workspace[ia], b = workspace[ib];
            workspace[ia] = b;
            workspace[ib] = a + b;
        }
    for (uint32_t tid = 0; tid < PS_WG; ++tid) {
        const uint32_t elt = tid + bid * PS_WG;
        if (elt < data_len) data[elt] = workspace[tid];
    }
}

static void k_prefix_sum(uint32_t *data, uint32_t data_len, uint32_t *aux, uint32_t ngroups) {
    for (uint32_t bid = 0; bid < ngroups; ++bid) k_prefix_sum_wg(data, data_len, aux, bid);
}

static void k_add_data_grp(uint32_t *data, uint32_t data_len, const uint32_t *aux, uint32_t ngroups) { /* :139-147 */
    for (uint32_t bid = 0; bid < ngroups; ++bid)
        for (uint32_t t = 0; t < PS_WG; ++t) {
            const uint32_t tid = bid * PS_WG + t;
            if (tid < data_len) data[tid] += aux[bid];
        }
}

/* WgPrefixSum::dispatch (prefix_sum.rs:49-99) on a whole vector of n u32, in place.  n == 0: the reference's
 * PrefixSumWorkspace::reserve never terminates (0.div_ceil(256) == 0 != 1, prefix_sum.rs:188-206); here it is a no-op. */
int orc_prefix_sum(uint32_t *data, uint32_t n) {
    if (n == 0) return ORC_OK;
    uint32_t lens[8];
    uint32_t *bufs[8];
    int num_stages = 0;
    uint32_t stage_len = (n + PS_WG - 1) / PS_WG; /* reserve(): :186-206 */
    while (stage_len != 1) {
        lens[num_stages++] = stage_len;
        stage_len = (stage_len + PS_WG - 1) / PS_WG;
    }
    lens[num_stages++] = 1;
    for (int i = 0; i < num_stages; ++i) bufs[i] = (uint32_t *)calloc(lens[i], sizeof(uint32_t));
    k_prefix_sum(data, n, bufs[0], lens[0]);                                             /* :64-68 */
    for (int i = 0; i < num_stages - 1; ++i) k_prefix_sum(bufs[i], lens[i], bufs[i + 1], lens[i + 1]); /* :70-78 */
    if (num_stages > 2)                                                                  /* :80-90 */
        for (int i = num_stages - 3; i >= 0; --i) k_add_data_grp(bufs[i], lens[i], bufs[i + 1], lens[i + 1]);
    if (num_stages > 1) k_add_data_grp(data, n, bufs[0], lens[0]);                       /* :92-96 */
    for (int i = 0; i < num_stages; ++i) free(bufs[i]);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------------------------- */
#define RS_WG 256u
#define RS_EPT 4u
#define RS_BLOCK (RS_WG * RS_EPT) /* sorting.wgsl: BLOCK_SIZE = 1024 */
#define RS_BINS 16u

/* One 4-bit pass: sort_count -> (sort_reduce, sort_scan, sort_scan_add) -> sort_scatter. */
static void radix_pass(uint32_t shift, uint32_t num_keys, const uint32_t *src, const uint32_t *values, uint32_t *counts,
                       uint32_t *out, uint32_t *out_values) {
    const uint32_t num_wgs = (num_keys + RS_BLOCK - 1) / RS_BLOCK;
    /* sort_count.wgsl: histogram of each workgroup's 1024 keys, stored bin-major */
    memset(counts, 0, (size_t)num_wgs * RS_BINS * sizeof(uint32_t));
    for (uint32_t wg = 0; wg < num_wgs; ++wg)
        for (uint32_t e = 0; e < RS_BLOCK; ++e) {
            const uint32_t idx = wg * RS_BLOCK + e;
            if (idx < num_keys) counts[((src[idx] >> shift) & 0xFu) * num_wgs + wg] += 1u;
        }
    /* sort_reduce.wgsl (sum per bin and 1024-workgroup chunk), sort_scan.wgsl (exclusive scan of those sums) and
     * sort_scan_add.wgsl (exclusive scan inside each chunk + the chunk's base): an exclusive scan of counts[] in storage
     * order, i.e. by (bin, workgroup). */
    uint32_t run = 0;
    for (size_t i = 0; i < (size_t)num_wgs * RS_BINS; ++i) {
        const uint32_t c = counts[i];
        counts[i] = run;
        run += c;
    }
    /* sort_scatter.wgsl: per workgroup, four sub-blocks of 256 keys in order (data_index = start + tid + i * 256); inside a
     * sub-block keys of equal digit keep their thread order (the two 2-bit split steps are stable), and the per-bin cursor
     * advances by the sub-block's histogram. */
    for (uint32_t wg = 0; wg < num_wgs; ++wg) {
        uint32_t cursor[RS_BINS];
        for (uint32_t b = 0; b < RS_BINS; ++b) cursor[b] = counts[b * num_wgs + wg];
        for (uint32_t e = 0; e < RS_BLOCK; ++e) {
            const uint32_t idx = wg * RS_BLOCK + e;
            if (idx >= num_keys) break; /* padding keys (~0) sort last and fall outside [0, num_keys) */
            const uint32_t k = src[idx], pos = cursor[(k >> shift) & 0xFu]++;
            if (pos < num_keys) {
                out[pos] = k;
                out_values[pos] = values[idx];
            }
        }
    }
}

/* RadixSort::dispatch (radix_sort/mod.rs:111-223).  The first n_sort entries of output_* receive the pairs of the first
 * n_sort entries of input_*, stably ordered by the low 4 * ceil(sorting_bits / 4) key bits; entries beyond n_sort and the
 * inputs are untouched; sorting_bits == 0 runs no pass at all (outputs untouched). */
int orc_radix_sort(const uint32_t *input_keys, const uint32_t *input_values, uint32_t len, uint32_t n_sort, uint32_t sorting_bits,
                   uint32_t *output_keys, uint32_t *output_values) {
    if (sorting_bits > 32) return ORC_UNSUPPORTED; /* assert!(sorting_bits <= 32) :126 */
    if (n_sort > len) n_sort = len;
    const uint32_t num_passes = (sorting_bits + 3) / 4;
    if (num_passes == 0 || len == 0) return ORC_OK;
    const uint32_t max_wgs = (len + RS_BLOCK - 1) / RS_BLOCK;
    uint32_t *counts = (uint32_t *)malloc((size_t)max_wgs * RS_BINS * sizeof(uint32_t));
    uint32_t *pong_k = (uint32_t *)malloc((size_t)len * sizeof(uint32_t));
    uint32_t *pong_v = (uint32_t *)malloc((size_t)len * sizeof(uint32_t));
    const uint32_t *cur_k = input_keys, *cur_v = input_values;
    uint32_t *out_k = output_keys, *out_v = output_values, *alt_k = pong_k, *alt_v = pong_v;
    if (num_passes % 2 == 0) { /* :163-166 */
        uint32_t *t = out_k; out_k = alt_k; alt_k = t;
        t = out_v; out_v = alt_v; alt_v = t;
    }
    for (uint32_t p = 0; p < num_passes; ++p) {
        radix_pass(p * 4, n_sort, cur_k, cur_v, counts, out_k, out_v);
        /* :212-221: after pass 0 the current pair is the one just written and the other buffer becomes the target */
        const uint32_t *nk = out_k, *nv = out_v;
        if (p == 0) {
            out_k = alt_k; out_v = alt_v;
        } else {
            out_k = (uint32_t *)cur_k; out_v = (uint32_t *)cur_v;
        }
        cur_k = nk; cur_v = nv;
    }
    free(counts);
    free(pong_k);
    free(pong_v);
    return ORC_OK;
}
