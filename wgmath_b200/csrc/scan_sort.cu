// scan_sort.cu — the two integer primitives SURVEY.md §8(f) lists next to the linalg path, HBM-bound u32 work:
//
//   wgb_prefix_sum   replaces WgPrefixSum::dispatch (/root/reference/crates/wgrapier/src/dynamics/prefix_sum.rs:49-99 +
//                    prefix_sum.wgsl:35-147): in-place exclusive prefix sum of a u32 vector (wrapping adds).  The reference scans
//                    256-element blocks with a Blelloch tree, recurses over the block totals (one dispatch per level) and
//                    adds them back level by level: 3 passes over the data (read+write, read+write of aux, read+write) and
//                    2 * levels dispatches, one 256-thread workgroup per 256 elements with ~18 barriers each.  Here: reduce /
//                    scan-the-partials / apply with one warp per 1024 elements and no block barriers, the apply step walking
//                    backwards so its re-read of the tail comes from L2 (12 B per element, less what L2 still holds).
//   wgb_radix_sort   replaces RadixSort::dispatch (crates/wgparry/src/utils/radix_sort/mod.rs:111-223 + sort_*.wgsl): stable
//                    LSD sort of (u32 key, u32 value) pairs by the low 4 * ceil(sorting_bits / 4) key bits, pair count read
//                    from device memory.  The reference runs ceil(bits / 4) passes of 5 dispatches (count, reduce, scan,
//                    scan_add, scatter), each pass reading the keys twice: 20 B per pair per 4 bits.  Here the same
//                    count / scan / scatter structure with 8-bit digits (half the passes: 20 B per pair per 8 bits), 8192-pair
//                    tiles ranked with warp match instead of 2-bit split scans, tiles sorted in shared memory and written out
//                    in runs, and the scatter walking the tiles backwards so part of its key re-read comes from L2.  (A
//                    single-sweep variant with per-digit decoupled look-back was built first: 23 Gpair/s — the look-back chain
//                    is the critical path at B200 bandwidth, see profiles/README.md.)  A stable sort by the same masked key
//                    has exactly one result, so the output is bit-identical to the reference's.
//
// Data layout: plain u32 arrays (GpuVector<u32>), element offsets from the view shape.  Scan partials, digit counts and the
// ping-pong buffers live in context workspace slots; nothing needs clearing between calls.
#include <cstdlib>

#include "common.cuh"

namespace wgb {

namespace {

// ---------------------------------------------------------------------------------------------
// exclusive prefix sum: reduce, scan the partials, scan with carry-in
// ---------------------------------------------------------------------------------------------
// Unit of work = one warp and 1024 consecutive elements ("warp tile": 8 rounds of one 128-bit access per lane, round r /
// lane l covers elements [128 r + 4 l, +4)); warps never synchronise with each other, so every warp streams like a copy
// kernel.  Three steps, all in-order launches (no spinning, safe under graph capture and profiler replay):
//   1. scan_reduce_kernel   partial[w] = sum of warp tile w                       (4 B / element read)
//   2. the partials are scanned by the same routine, recursively (ceil(n / 1024) values; one CTA when <= 4096)
//   3. scan_apply_kernel    exclusive scan inside each warp tile + partial[w]     (4 B read + 4 B written)
// Step 3 walks the tiles in DESCENDING order: the tail of the array, read last by step 1, is still in the 126 MB L2.
// Why not a single pass: two decoupled look-back variants were built and measured first (profiles/README.md) — at B200
// bandwidth ~100 tiles start per microsecond while a descriptor round trip through L2 costs most of one, so every tile
// waits on hundreds of predecessors: 2.4 TB/s (one-warp window), 2.4 TB/s (whole-CTA window), 2.9 TB/s (two-level group
// descriptors) against 4.4-5.1 TB/s for the same kernels with the look-back switched off.  A fourth variant (16384-element
// tiles, 64-descriptor window, ticketed tiles) was validated bit-exact in round 2 and measured 159 us vs 131 us for the three-step
// form at n = 2^26 (3.4 vs 4.1 TB/s of algorithmic bytes): also retired (profiles/README.md).  A fifth, without any look-back chain
// (one persistent kernel, chunks re-read from L2 between a reduce and a scan pass, chunk-level counters only), was correct and
// took 175 us: profiles/r2_scan_stream_experiment/.
constexpr int kScanRounds = 8;                         // 128-bit accesses per lane and warp tile
constexpr int kScanWarpTile = kScanRounds * 128;       // 1024 elements per warp tile
constexpr int kScanThreads = 256;
constexpr int kScanSmall = 4096;          // largest vector one CTA scans directly (1024 threads x 4 elements)

__device__ __forceinline__ void scan_load_tile(const uint32_t *__restrict__ p, uint64_t base, uint64_t n, bool vec, int lane,
                                               uint4 (&v)[kScanRounds], bool keep) {
#pragma unroll
    for (int r = 0; r < kScanRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * 128 + (uint64_t)lane * 4;
        if (vec && i + 4 <= n) {
            v[r] = keep ? __ldg(reinterpret_cast<const uint4 *>(p + i)) : __ldcs(reinterpret_cast<const uint4 *>(p + i));
        } else {
            v[r].x = i + 0 < n ? p[i + 0] : 0u;
            v[r].y = i + 1 < n ? p[i + 1] : 0u;
            v[r].z = i + 2 < n ? p[i + 2] : 0u;
            v[r].w = i + 3 < n ? p[i + 3] : 0u;
        }
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *__restrict__ data, uint64_t n, uint64_t tiles,
                                                                   uint32_t *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * kScanThreads + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * kScanThreads) >> 5;
    const bool vec = (reinterpret_cast<uintptr_t>(data) & 15u) == 0;
    for (uint64_t w = warp; w < tiles; w += nwarps) {
        uint4 v[kScanRounds];
        scan_load_tile(data, w * kScanWarpTile, n, vec, lane, v, true);   // default caching: step 3 re-reads the tail from L2
        uint32_t s = 0;
#pragma unroll
        for (int r = 0; r < kScanRounds; ++r) s += v[r].x + v[r].y + v[r].z + v[r].w;
        s = __reduce_add_sync(0xFFFFFFFFu, s);
        if (lane == 0) partial[w] = s;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *__restrict__ data, uint64_t n, uint64_t tiles,
                                                                  const uint32_t *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * kScanThreads + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * kScanThreads) >> 5;
    const bool vec = (reinterpret_cast<uintptr_t>(data) & 15u) == 0;
    for (uint64_t k = warp; k < tiles; k += nwarps) {
        const uint64_t w = tiles - 1 - k;   // descending: most recently read tiles first
        const uint64_t base = w * kScanWarpTile;
        uint4 v[kScanRounds];
        scan_load_tile(data, base, n, vec, lane, v, false);
        const uint32_t carry = partial ? __ldg(partial + w) : 0u;
        uint32_t sum[kScanRounds], run = carry;
#pragma unroll
        for (int r = 0; r < kScanRounds; ++r) {
            sum[r] = v[r].x + v[r].y + v[r].z + v[r].w;
            uint32_t x = sum[r];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += y;
            }
            const uint32_t excl = run + x - sum[r];            // everything before this lane's four elements
            run += __shfl_sync(0xFFFFFFFFu, x, 31);
            uint4 o4;
            o4.x = excl;
            o4.y = o4.x + v[r].x;
            o4.z = o4.y + v[r].y;
            o4.w = o4.z + v[r].z;
            const uint64_t i = base + (uint64_t)r * 128 + (uint64_t)lane * 4;
            if (vec && i + 4 <= n) {
                __stcs(reinterpret_cast<uint4 *>(data + i), o4);
            } else {
                if (i + 0 < n) data[i + 0] = o4.x;
                if (i + 1 < n) data[i + 1] = o4.y;
                if (i + 2 < n) data[i + 2] = o4.z;
                if (i + 3 < n) data[i + 3] = o4.w;
            }
        }
    }
}

// One CTA scans up to 4096 elements in place: 1024 threads, four consecutive elements each.
__global__ void __launch_bounds__(1024) scan_small_kernel(uint32_t *__restrict__ data, uint32_t n) {
    __shared__ uint32_t s_w[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (uint32_t)t * 4 + q < n ? data[t * 4 + q] : 0u;
    const uint32_t s = v[0] + v[1] + v[2] + v[3];
    uint32_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const uint32_t wv = s_w[lane];
        uint32_t y2 = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, y2, o);
            if (lane >= o) y2 += y;
        }
        s_w[lane] = y2 - wv;
    }
    __syncthreads();
    uint32_t run = s_w[warp] + x - s;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if ((uint32_t)t * 4 + q < n) data[t * 4 + q] = run;
        run += v[q];
    }
}

// ---------------------------------------------------------------------------------------------
// radix sort: per 8-bit digit  tile histograms -> exclusive scan of the counts -> stable scatter
// ---------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;   // 4096 pairs per tile
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsBins = 256;

__device__ __forceinline__ uint32_t rs_count(const uint32_t *n_ptr, uint32_t len) {
    const uint32_t n = *n_ptr;
    return n < len ? n : len;
}

// counts[d * tiles + tile] = number of keys of the tile whose digit is d (digit-major, like the reference's
// counts[bin * num_wgs + wg], sort_count.wgsl).  Every tile of the launch writes its 256 counts (zeros past *n_sort), so
// the array needs no clearing.  4 B per key.
__global__ void __launch_bounds__(kRsThreads) radix_count_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ n_ptr,
                                                                 uint32_t len, uint32_t shift, uint32_t mask, uint32_t tiles,
                                                                 uint32_t *__restrict__ counts) {
    __shared__ unsigned int s_hist[kRsWarps][kRsBins];   // one private histogram per warp: no cross-warp contention
    const int t = threadIdx.x, warp = t >> 5;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) s_hist[w][t] = 0u;
    __syncthreads();
    const uint32_t tile = blockIdx.x;
    const uint32_t n = rs_count(n_ptr, len);
    const uint32_t tile_base = tile * (uint32_t)kRsTile;
    if (tile_base < n) {
        const uint32_t tile_n = min((uint32_t)kRsTile, n - tile_base);
        const uint32_t *kp = keys + tile_base;
        unsigned int *h = s_hist[warp];
        if (tile_n == (uint32_t)kRsTile && (reinterpret_cast<uintptr_t>(kp) & 15u) == 0) {
            uint4 k[kRsItems / 4];
#pragma unroll
            for (int i = 0; i < kRsItems / 4; ++i) k[i] = __ldg(reinterpret_cast<const uint4 *>(kp) + i * kRsThreads + t);
#pragma unroll
            for (int i = 0; i < kRsItems / 4; ++i) {
                atomicAdd(h + ((k[i].x >> shift) & mask), 1u);
                atomicAdd(h + ((k[i].y >> shift) & mask), 1u);
                atomicAdd(h + ((k[i].z >> shift) & mask), 1u);
                atomicAdd(h + ((k[i].w >> shift) & mask), 1u);
            }
        } else {
            for (uint32_t e = t; e < tile_n; e += kRsThreads) atomicAdd(h + ((kp[e] >> shift) & mask), 1u);
        }
    }
    __syncthreads();
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) c += s_hist[w][t];
    counts[(size_t)t * tiles + tile] = c;
}

// Stable scatter of one tile.  Tile = 4096 consecutive pairs; warp w owns the 512 consecutive pairs [w * 512, (w + 1) * 512),
// item i of lane l is pair w * 512 + i * 32 + l, so (item, lane) order is input order and every global access of a warp is
// one contiguous 128 bytes.
//   1. rank inside the warp: lanes holding the same digit find each other with one ballot per digit bit; the warp's private
//      histogram gives the count of earlier equal digits in the warp;
//   2. thread d turns the 8 warp histograms of digit d into warp offsets and the tile count; an exclusive scan of the tile
//      counts gives each digit's start inside the tile; offsets[d * tiles + tile] (the scanned counts) is where the tile's
//      first pair of digit d goes globally;
//   3. pairs are placed in shared memory in tile-sorted order and written out by consecutive threads: equal digits form
//      contiguous global runs.
// Tiles are independent (all offsets are known), so the grid walks them in DESCENDING order: the keys the count kernel
// read last are still in L2.  16 B per pair.
// The kernel is instruction-bound, not memory-bound (ncu, profiles/README.md: ~140 warp instructions per pair at 50 % issue
// utilisation in the first version), hence: full tiles take a predicate-free instantiation, 16 items per thread keep four
// CTAs resident per SM, and the digit-bit ballots are batched four items at a time so they overlap.
template <bool FULL, bool MATCH>
__device__ __forceinline__ void radix_scatter_tile(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                   uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t tile_base,
                                                   uint32_t tile_n, uint32_t shift, uint32_t mask, uint32_t goff,
                                                   uint32_t (&s_warp_hist)[kRsWarps][kRsBins], uint32_t (&s_sorted)[kRsTile],
                                                   uint32_t (&s_tile_start)[kRsBins], uint32_t (&s_gbase)[kRsBins],
                                                   uint32_t (&s_scan)[kRsWarps]) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t key[kRsItems], val[kRsItems], slot[kRsItems];
    const uint32_t wbase = (uint32_t)warp * (kRsItems * 32) + lane;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t e = wbase + i * 32;
        key[i] = (FULL || e < tile_n) ? __ldcs(keys_in + tile_base + e) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t e = wbase + i * 32;
        val[i] = (FULL || e < tile_n) ? __ldcs(vals_in + tile_base + e) : 0u;
    }
    // 1. ranks inside the warp
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t *wh = s_warp_hist[warp];
#pragma unroll
    for (int i0 = 0; i0 < kRsItems; i0 += 4) {
        uint32_t peers[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t d = (key[i0 + q] >> shift) & mask;
            uint32_t pm = 0xFFFFFFFFu;
            if (MATCH) {   // one match.any instead of eight ballots (WGB_RS_MATCH=1; measured, see profiles/README.md)
                const bool valid = FULL || wbase + (i0 + q) * 32 < tile_n;
                pm = __match_any_sync(0xFFFFFFFFu, valid ? d : 0xFFFFFFFFu);
                if (!valid) pm = 0u;
            } else {
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (d >> b) & 1u;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
                    pm &= bit ? bal : ~bal;
                }
                if (!FULL) {
                    const bool valid = wbase + (i0 + q) * 32 < tile_n;
                    pm &= __ballot_sync(0xFFFFFFFFu, valid);
                    if (!valid) pm = 0u;
                }
            }
            peers[q] = pm;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t d = (key[i0 + q] >> shift) & mask;
            const uint32_t pm = peers[q];
            uint32_t pre = 0;
            if (pm != 0u && (pm & lt_mask) == 0u) {   // lowest lane of the group
                pre = wh[d];
                wh[d] = pre + __popc(pm);
            }
            __syncwarp();
            pre = __shfl_sync(0xFFFFFFFFu, pre, pm ? __ffs(pm) - 1 : lane);
            slot[i0 + q] = pre + __popc(pm & lt_mask);
        }
    }
    __syncthreads();
    // 2. digit t: warp offsets, tile count, start inside the tile
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
        const uint32_t c = s_warp_hist[w][t];
        s_warp_hist[w][t] = count;
        count += c;
    }
    uint32_t cx = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t b = __shfl_up_sync(0xFFFFFFFFu, cx, o);
        if (lane >= o) cx += b;
    }
    if (lane == 31) s_scan[warp] = cx;
    __syncthreads();
    uint32_t coff = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w)
        if (w < warp) coff += s_scan[w];
    const uint32_t tile_start = coff + cx - count;
    s_tile_start[t] = tile_start;
    s_gbase[t] = goff - tile_start;
    __syncthreads();
    // 3. place in tile-sorted order, write out in runs
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        if (FULL || wbase + i * 32 < tile_n) {
            const uint32_t d = (key[i] >> shift) & mask;
            slot[i] += s_tile_start[d] + wh[d];
            s_sorted[slot[i]] = key[i];
        }
    }
    __syncthreads();
    uint32_t dig4[kRsItems / 4];   // digit of the key at sorted position j * 256 + t, four per register
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t e = (uint32_t)j * kRsThreads + t;
        if ((j & 3) == 0) dig4[j / 4] = 0;
        if (FULL || e < tile_n) {
            const uint32_t k = s_sorted[e];
            const uint32_t d = (k >> shift) & mask;
            dig4[j / 4] |= d << ((j & 3) * 8);
            keys_out[s_gbase[d] + e] = k;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRsItems; ++i)
        if (FULL || wbase + i * 32 < tile_n) s_sorted[slot[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t e = (uint32_t)j * kRsThreads + t;
        if (FULL || e < tile_n) vals_out[s_gbase[(dig4[j / 4] >> ((j & 3) * 8)) & 0xFFu] + e] = s_sorted[e];
    }
}

template <bool MATCH>
__global__ void __launch_bounds__(kRsThreads, 3) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                      const uint32_t *__restrict__ vals_in,
                                                                      uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                      const uint32_t *__restrict__ n_ptr, uint32_t len, uint32_t shift,
                                                                      uint32_t mask, uint32_t tiles,
                                                                      const uint32_t *__restrict__ offsets) {
    __shared__ uint32_t s_warp_hist[kRsWarps][kRsBins];
    __shared__ uint32_t s_sorted[kRsTile];
    __shared__ uint32_t s_tile_start[kRsBins];   // first slot of digit d inside the tile-sorted order
    __shared__ uint32_t s_gbase[kRsBins];        // global position of the tile's first pair of digit d, minus s_tile_start[d]
    __shared__ uint32_t s_scan[kRsWarps];
    const int t = threadIdx.x;
    const uint32_t tile = gridDim.x - 1 - blockIdx.x;
    const uint32_t n = rs_count(n_ptr, len);
    const uint32_t tile_base = tile * (uint32_t)kRsTile;
    if (tile_base >= n) return;
    const uint32_t goff = __ldg(offsets + (size_t)t * tiles + tile);
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) s_warp_hist[w][t] = 0u;
    __syncthreads();
    const uint32_t tile_n = min((uint32_t)kRsTile, n - tile_base);
    if (tile_n == (uint32_t)kRsTile)
        radix_scatter_tile<true, MATCH>(keys_in, vals_in, keys_out, vals_out, tile_base, tile_n, shift, mask, goff, s_warp_hist, s_sorted,
                                 s_tile_start, s_gbase, s_scan);
    else
        radix_scatter_tile<false, MATCH>(keys_in, vals_in, keys_out, vals_out, tile_base, tile_n, shift, mask, goff, s_warp_hist, s_sorted,
                                  s_tile_start, s_gbase, s_scan);
}

}  // namespace

// Exclusive scan of data[0, n) in place; `ws` provides room for the partials of every level (scan_workspace_elems(n)).
static size_t scan_workspace_elems(uint64_t n) {
    size_t total = 0;
    while (n > (uint64_t)kScanSmall) {
        n = (n + kScanWarpTile - 1) / kScanWarpTile;
        total += (n + 3) & ~(size_t)3;   // keep every level 16-byte aligned
    }
    return total;
}

static wgb_status scan_in_place(wgb_pass *p, uint32_t *data, uint64_t n, uint32_t *ws) {
    if (n == 0) return WGB_OK;
    wgb_ctx *ctx = p->ctx;
    if (n <= (uint64_t)kScanSmall) {
        scan_small_kernel<<<1, 1024, 0, p->stream>>>(data, (uint32_t)n);
        WGB_CUDA(cudaGetLastError());
        count_launch(ctx);
        return WGB_OK;
    }
    const uint64_t tiles = (n + kScanWarpTile - 1) / kScanWarpTile;
    const uint64_t ctas_needed = (tiles + kScanThreads / 32 - 1) / (kScanThreads / 32);
    uint64_t grid = (uint64_t)ctx->prop.multiProcessorCount * 8;   // one resident wave: 8 CTAs of 256 threads per SM
    if (grid > ctas_needed) grid = ctas_needed;
    scan_reduce_kernel<<<(unsigned)grid, kScanThreads, 0, p->stream>>>(data, n, tiles, ws);
    WGB_CUDA(cudaGetLastError());
    count_launch(ctx);
    WGB_TRY(scan_in_place(p, ws, tiles, ws + ((tiles + 3) & ~(uint64_t)3)));
    scan_apply_kernel<<<(unsigned)grid, kScanThreads, 0, p->stream>>>(data, n, tiles, ws);
    WGB_CUDA(cudaGetLastError());
    count_launch(ctx);
    return WGB_OK;
}

wgb_status launch_prefix_sum(wgb_pass *p, uint32_t *data, uint64_t n) {
    if (n == 0) return WGB_OK;
    void *w = nullptr;
    WGB_TRY(workspace_reserve(p->ctx, 4, scan_workspace_elems(n) * 4 + 16, &w));
    return scan_in_place(p, data, n, reinterpret_cast<uint32_t *>(w));
}

wgb_status launch_radix_sort(wgb_pass *p, const uint32_t *keys_in, const uint32_t *vals_in, uint32_t len, const uint32_t *n_dev,
                             uint32_t sorting_bits, uint32_t *keys_out, uint32_t *vals_out) {
    if (sorting_bits > 32) WGB_FAIL(WGB_ERR_INVALID, "Can only sort up to 32 bits");   // radix_sort/mod.rs:126
    const uint32_t total_bits = 4 * ((sorting_bits + 3) / 4);   // the reference sorts whole 4-bit digits (mod.rs:156)
    if (total_bits == 0 || len == 0) return WGB_OK;             // zero passes: outputs untouched, like the reference
    wgb_ctx *ctx = p->ctx;
    const uint32_t passes = (total_bits + 7) / 8;
    const uint32_t tiles = (len + kRsTile - 1) / kRsTile;
    // workspace: [counts 256 x tiles][scan partials][pong keys][pong values]
    const size_t n_counts = (size_t)kRsBins * tiles;
    const size_t counts_bytes = (n_counts * 4 + 255) & ~(size_t)255;
    const size_t scan_bytes = (scan_workspace_elems(n_counts) * 4 + 16 + 255) & ~(size_t)255;
    const size_t pong_bytes = passes > 1 ? (((size_t)len * 4 + 255) & ~(size_t)255) : 0;
    void *w = nullptr;
    WGB_TRY(workspace_reserve(ctx, 5, counts_bytes + scan_bytes + 2 * pong_bytes, &w));
    uint32_t *counts = reinterpret_cast<uint32_t *>(w);
    uint32_t *scan_ws = reinterpret_cast<uint32_t *>((char *)w + counts_bytes);
    uint32_t *pong_k = reinterpret_cast<uint32_t *>((char *)w + counts_bytes + scan_bytes);
    uint32_t *pong_v = reinterpret_cast<uint32_t *>((char *)w + counts_bytes + scan_bytes + pong_bytes);
    // the last pass lands in the caller's output; passes alternate between it and the workspace pair (mod.rs:163-221)
    const uint32_t *cur_k = keys_in, *cur_v = vals_in;
    for (uint32_t i = 0; i < passes; ++i) {
        const uint32_t shift = 8 * i, nb = total_bits - shift < 8 ? total_bits - shift : 8, mask = (1u << nb) - 1u;
        const bool to_out = (passes - 1 - i) % 2 == 0;
        uint32_t *dk = to_out ? keys_out : pong_k, *dv = to_out ? vals_out : pong_v;
        radix_count_kernel<<<tiles, kRsThreads, 0, p->stream>>>(cur_k, n_dev, len, shift, mask, tiles, counts);
        WGB_CUDA(cudaGetLastError());
        count_launch(ctx);
        WGB_TRY(scan_in_place(p, counts, n_counts, scan_ws));
        static const bool use_match = [] { const char *e = getenv("WGB_RS_MATCH"); return e && atoi(e) != 0; }();
        if (use_match) radix_scatter_kernel<true><<<tiles, kRsThreads, 0, p->stream>>>(cur_k, cur_v, dk, dv, n_dev, len, shift, mask, tiles, counts);
        else radix_scatter_kernel<false><<<tiles, kRsThreads, 0, p->stream>>>(cur_k, cur_v, dk, dv, n_dev, len, shift, mask, tiles, counts);
        WGB_CUDA(cudaGetLastError());
        count_launch(ctx);
        cur_k = dk;
        cur_v = dv;
    }
    return WGB_OK;
}

}  // namespace wgb
