// scan_sort.cu — the two integer primitives SURVEY.md §8(f) lists next to the linalg path, HBM-bound u32 work:
//
//   wgb_prefix_sum   replaces WgPrefixSum::dispatch (/root/reference/crates/wgrapier/src/dynamics/prefix_sum.rs:49-99 +
//                    prefix_sum.wgsl:35-147): in-place exclusive prefix sum of a u32 vector (wrapping adds).  The reference scans
//                    256-element blocks with a Blelloch tree, recurses over the block totals (one dispatch per level) and
//                    adds them back level by level: 3 passes over the data (read+write, read+write of aux, read+write) and
//                    2 * levels dispatches.  Here: ONE launch, single pass, decoupled look-back across 4096-element tiles —
//                    8 bytes of HBM traffic per element (read + write), the algorithmic minimum.
//   wgb_radix_sort   replaces RadixSort::dispatch (crates/wgparry/src/utils/radix_sort/mod.rs:111-223 + sort_*.wgsl): stable
//                    LSD sort of (u32 key, u32 value) pairs by the low 4 * ceil(sorting_bits / 4) key bits, pair count read
//                    from device memory.  The reference runs ceil(bits / 4) passes of 5 dispatches (count, reduce, scan,
//                    scan_add, scatter), each pass reading the keys twice: 20 B per pair per 4 bits.  Here: one histogram
//                    launch for all digits, then one "onesweep" launch per 8-bit digit (tile-local ranking with warp
//                    match, per-digit decoupled look-back for the global offsets, tile sorted in shared memory and written
//                    out in runs): 16 B per pair per 8 bits + 4 B once.  A stable sort by the same masked key has exactly
//                    one result, so the output is bit-identical to the reference's.
//
// Data layout: plain u32 arrays (GpuVector<u32>), element offsets from the view shape.  Look-back descriptors and the
// ping-pong buffers live in context workspace slots and are cleared with one memset node per call.
#include "common.cuh"

namespace wgb {

namespace {

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// exclusive prefix sum, single pass with decoupled look-back
// ---------------------------------------------------------------------------------------------
constexpr int kPsThreads = 256;
constexpr int kPsRounds = 4;                                // 128-bit accesses per thread and tile
constexpr int kPsTile = kPsThreads * 4 * kPsRounds;         // 4096 elements
constexpr unsigned long long kPsAggregate = 1ull << 32, kPsPrefix = 2ull << 32;

// desc[tile] = (state << 32) | value; state 0 = not yet published, 1 = tile aggregate, 2 = inclusive prefix up to the tile.
// Tiles are handed out by an atomic ticket, so a tile only ever waits for tiles whose CTAs are already running.
__global__ void __launch_bounds__(kPsThreads) prefix_sum_kernel(uint32_t *__restrict__ data, uint64_t n,
                                                                unsigned long long *__restrict__ desc,
                                                                unsigned int *__restrict__ ticket) {
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_wt[kPsRounds * (kPsThreads / 32)];   // warp totals in (round, warp) order = element order
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * kPsTile;
    const bool full = base + kPsTile <= n && (reinterpret_cast<uintptr_t>(data) & 15u) == 0;

    uint4 v[kPsRounds];
    uint32_t sum[kPsRounds], incl[kPsRounds];
#pragma unroll
    for (int r = 0; r < kPsRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * (kPsThreads * 4) + (uint64_t)t * 4;
        if (full) {
            v[r] = __ldcs(reinterpret_cast<const uint4 *>(data + i));
        } else {
            v[r].x = i + 0 < n ? data[i + 0] : 0u;
            v[r].y = i + 1 < n ? data[i + 1] : 0u;
            v[r].z = i + 2 < n ? data[i + 2] : 0u;
            v[r].w = i + 3 < n ? data[i + 3] : 0u;
        }
        sum[r] = v[r].x + v[r].y + v[r].z + v[r].w;
    }
#pragma unroll
    for (int r = 0; r < kPsRounds; ++r) {
        uint32_t x = sum[r];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        incl[r] = x;
        if (lane == 31) s_wt[r * (kPsThreads / 32) + warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
        // 32 warp totals -> exclusive offsets; the last lane ends up with the tile aggregate
        const uint32_t wt = s_wt[lane];
        uint32_t x = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        s_wt[lane] = x - wt;
        const uint32_t aggregate = __shfl_sync(0xFFFFFFFFu, x, 31);
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(desc, kPsPrefix | aggregate);
        } else {
            if (lane == 0) st_relaxed_u64(desc + tile, kPsAggregate | aggregate);
            int64_t pred = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = pred - lane;
                unsigned long long d = kPsPrefix;   // before tile 0: prefix 0
                if (idx >= 0) {
                    do {
                        d = ld_relaxed_u64(desc + idx);
                    } while ((d >> 32) == 0);
                }
                const unsigned has_prefix = __ballot_sync(0xFFFFFFFFu, (d >> 32) == 2);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;   // nearest predecessor holding a full prefix
                uint32_t x2 = lane <= first ? (uint32_t)d : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x2 += __shfl_xor_sync(0xFFFFFFFFu, x2, o);
                exclusive += x2;
                if (has_prefix) break;
                pred -= 32;
            }
            if (lane == 0) st_relaxed_u64(desc + tile, kPsPrefix | (uint32_t)(exclusive + aggregate));
        }
        if (lane == 0) s_excl = exclusive;
    }
    __syncthreads();
    const uint32_t tile_excl = s_excl;
#pragma unroll
    for (int r = 0; r < kPsRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * (kPsThreads * 4) + (uint64_t)t * 4;
        uint4 o;
        o.x = tile_excl + s_wt[r * (kPsThreads / 32) + warp] + (incl[r] - sum[r]);
        o.y = o.x + v[r].x;
        o.z = o.y + v[r].y;
        o.w = o.z + v[r].z;
        if (full) {
            __stcs(reinterpret_cast<uint4 *>(data + i), o);
        } else {
            if (i + 0 < n) data[i + 0] = o.x;
            if (i + 1 < n) data[i + 1] = o.y;
            if (i + 2 < n) data[i + 2] = o.z;
            if (i + 3 < n) data[i + 3] = o.w;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// radix sort: global digit histograms + one onesweep pass per 8-bit digit
// ---------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;   // 4096 pairs per tile
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsBins = 256;
constexpr int kRsMaxPasses = 4;
constexpr unsigned int kRsFlagAggregate = 1u << 30, kRsFlagPrefix = 2u << 30, kRsValueMask = (1u << 30) - 1u;

struct RsPasses {
    uint32_t num, shift[kRsMaxPasses], mask[kRsMaxPasses];
};

__device__ __forceinline__ uint32_t rs_count(const uint32_t *n_ptr, uint32_t len) {
    const uint32_t n = *n_ptr;
    return n < len ? n : len;
}

// hist[p][d] += number of keys whose digit p equals d (all passes in one sweep over the keys: 4 B per key)
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ n_ptr,
                                                         uint32_t len, RsPasses ps, unsigned int *__restrict__ hist) {
    __shared__ unsigned int sh[kRsMaxPasses * kRsBins];
    for (int i = threadIdx.x; i < kRsMaxPasses * kRsBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const uint32_t n = rs_count(n_ptr, len);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool vec = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0;
    if (vec) {
        const uint64_t n4 = n / 4;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const uint4 k = __ldcs(reinterpret_cast<const uint4 *>(keys) + i);
            for (uint32_t p = 0; p < ps.num; ++p) {
                atomicAdd(&sh[p * kRsBins + ((k.x >> ps.shift[p]) & ps.mask[p])], 1u);
                atomicAdd(&sh[p * kRsBins + ((k.y >> ps.shift[p]) & ps.mask[p])], 1u);
                atomicAdd(&sh[p * kRsBins + ((k.z >> ps.shift[p]) & ps.mask[p])], 1u);
                atomicAdd(&sh[p * kRsBins + ((k.w >> ps.shift[p]) & ps.mask[p])], 1u);
            }
        }
        for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint32_t k = keys[i];
            for (uint32_t p = 0; p < ps.num; ++p) atomicAdd(&sh[p * kRsBins + ((k >> ps.shift[p]) & ps.mask[p])], 1u);
        }
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint32_t k = keys[i];
            for (uint32_t p = 0; p < ps.num; ++p) atomicAdd(&sh[p * kRsBins + ((k >> ps.shift[p]) & ps.mask[p])], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)ps.num * kRsBins; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, sh[i]);
}

// One 8-bit (or narrower, last) digit.  Tile = 4096 consecutive pairs; warp w owns the 512 consecutive pairs
// [w * 512, (w + 1) * 512), item i of lane l is pair w * 512 + i * 32 + l, so (item, lane) order is input order and every
// global access of a warp is one contiguous 128 bytes.
//   1. rank inside the warp: lanes holding the same digit find each other with match.any; the warp's private histogram
//      gives the count of earlier equal digits in the warp;
//   2. thread d turns the 8 warp histograms of digit d into warp offsets and the tile count, publishes the count, and
//      walks back over the preceding tiles' descriptors until it meets an inclusive prefix (decoupled look-back);
//   3. an exclusive scan of the tile counts gives each digit's start inside the tile; pairs are placed in shared memory
//      in tile-sorted order and written out by consecutive threads: equal digits form contiguous global runs.
__global__ void __launch_bounds__(kRsThreads) radix_onesweep_kernel(const uint32_t *__restrict__ keys_in,
                                                                    const uint32_t *__restrict__ vals_in,
                                                                    uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                    const uint32_t *__restrict__ n_ptr, uint32_t len, uint32_t shift,
                                                                    uint32_t mask, const unsigned int *__restrict__ hist,
                                                                    unsigned int *__restrict__ desc, unsigned int *__restrict__ ticket) {
    __shared__ uint32_t s_warp_hist[kRsWarps][kRsBins];
    __shared__ uint32_t s_sorted[kRsTile];
    __shared__ uint32_t s_tile_start[kRsBins];   // first slot of digit d inside the tile-sorted order
    __shared__ uint32_t s_gbase[kRsBins];        // global position of the tile's first pair of digit d, minus s_tile_start[d]
    __shared__ uint32_t s_scan[kRsWarps];
    __shared__ uint32_t s_tile;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) s_warp_hist[w][t] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n = rs_count(n_ptr, len);
    const uint32_t num_tiles = (n + kRsTile - 1) / kRsTile;
    if (tile >= num_tiles) return;
    const uint32_t tile_base = tile * (uint32_t)kRsTile;
    const uint32_t tile_n = min((uint32_t)kRsTile, n - tile_base);

    uint32_t key[kRsItems], val[kRsItems];
    uint16_t slot[kRsItems];
    const uint32_t wbase = (uint32_t)warp * (kRsItems * 32) + lane;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t e = wbase + i * 32;
        key[i] = e < tile_n ? __ldcs(keys_in + tile_base + e) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t e = wbase + i * 32;
        val[i] = e < tile_n ? __ldcs(vals_in + tile_base + e) : 0u;
    }
    // 1. ranks inside the warp
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t *wh = s_warp_hist[warp];
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const bool valid = wbase + i * 32 < tile_n;
        const uint32_t d = (key[i] >> shift) & mask;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 0x100u);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (lane == leader && valid) {
            pre = wh[d];
            wh[d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xFFFFFFFFu, pre, leader);
        slot[i] = (uint16_t)(pre + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();
    // 2. digit t: warp offsets, tile count, look-back
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
        const uint32_t c = s_warp_hist[w][t];
        s_warp_hist[w][t] = count;
        count += c;
    }
    unsigned int *my_desc = desc + (size_t)tile * kRsBins + t;
    uint32_t exclusive = 0;
    if (tile == 0) {
        st_relaxed_u32(my_desc, kRsFlagPrefix | count);
    } else {
        st_relaxed_u32(my_desc, kRsFlagAggregate | count);
        const unsigned int *pd = my_desc - kRsBins;
        while (true) {
            unsigned int dsc;
            do {
                dsc = ld_relaxed_u32(pd);
            } while ((dsc >> 30) == 0);
            exclusive += dsc & kRsValueMask;
            if ((dsc >> 30) == 2) break;
            pd -= kRsBins;
        }
        st_relaxed_u32(my_desc, kRsFlagPrefix | (exclusive + count));
    }
    // global start of digit t = number of keys with a smaller digit (exclusive scan of the histogram) ...
    const uint32_t h = hist[t];
    uint32_t hx = h, cx = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, hx, o), b = __shfl_up_sync(0xFFFFFFFFu, cx, o);
        if (lane >= o) { hx += a; cx += b; }
    }
    __shared__ uint32_t s_hscan[kRsWarps];
    if (lane == 31) { s_hscan[warp] = hx; s_scan[warp] = cx; }
    __syncthreads();
    uint32_t hoff = 0, coff = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w)
        if (w < warp) { hoff += s_hscan[w]; coff += s_scan[w]; }
    const uint32_t digit_start = hoff + hx - h;    // ... over all tiles
    const uint32_t tile_start = coff + cx - count; // ... and inside this tile
    s_tile_start[t] = tile_start;
    s_gbase[t] = digit_start + exclusive - tile_start;
    __syncthreads();
    // 3. place in tile-sorted order, write out in runs
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        if (wbase + i * 32 < tile_n) {
            const uint32_t d = (key[i] >> shift) & mask;
            slot[i] = (uint16_t)(s_tile_start[d] + wh[d] + slot[i]);
            s_sorted[slot[i]] = key[i];
        }
    }
    __syncthreads();
    uint32_t gpos[kRsItems];
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t e = (uint32_t)j * kRsThreads + t;
        if (e < tile_n) {
            const uint32_t k = s_sorted[e];
            gpos[j] = s_gbase[(k >> shift) & mask] + e;
            keys_out[gpos[j]] = k;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRsItems; ++i)
        if (wbase + i * 32 < tile_n) s_sorted[slot[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t e = (uint32_t)j * kRsThreads + t;
        if (e < tile_n) vals_out[gpos[j]] = s_sorted[e];
    }
}

}  // namespace

wgb_status launch_prefix_sum(wgb_pass *p, uint32_t *data, uint64_t n) {
    if (n == 0) return WGB_OK;
    wgb_ctx *ctx = p->ctx;
    const uint64_t tiles = (n + kPsTile - 1) / kPsTile;
    void *w = nullptr;
    const size_t bytes = 16 + tiles * sizeof(unsigned long long);
    WGB_TRY(workspace_reserve(ctx, 4, bytes, &w));
    WGB_CUDA(cudaMemsetAsync(w, 0, bytes, p->stream));
    prefix_sum_kernel<<<(unsigned)tiles, kPsThreads, 0, p->stream>>>(data, n, reinterpret_cast<unsigned long long *>((char *)w + 16),
                                                                     reinterpret_cast<unsigned int *>(w));
    WGB_CUDA(cudaGetLastError());
    count_launch(ctx);
    return WGB_OK;
}

wgb_status launch_radix_sort(wgb_pass *p, const uint32_t *keys_in, const uint32_t *vals_in, uint32_t len, const uint32_t *n_dev,
                             uint32_t sorting_bits, uint32_t *keys_out, uint32_t *vals_out) {
    if (sorting_bits > 32) WGB_FAIL(WGB_ERR_INVALID, "Can only sort up to 32 bits");   // radix_sort/mod.rs:126
    const uint32_t total_bits = 4 * ((sorting_bits + 3) / 4);   // the reference sorts whole 4-bit digits (mod.rs:156)
    if (total_bits == 0 || len == 0) return WGB_OK;             // zero passes: outputs untouched, like the reference
    if (len > kRsValueMask) WGB_FAIL(WGB_ERR_UNSUPPORTED, "radix sort: more than 2^30 - 1 pairs");
    wgb_ctx *ctx = p->ctx;
    RsPasses ps{};
    ps.num = (total_bits + 7) / 8;
    for (uint32_t i = 0; i < ps.num; ++i) {
        ps.shift[i] = 8 * i;
        const uint32_t nb = total_bits - 8 * i < 8 ? total_bits - 8 * i : 8;
        ps.mask[i] = (1u << nb) - 1u;
    }
    const uint32_t tiles = (len + kRsTile - 1) / kRsTile;
    // workspace: [hist 4 x 256][tickets 4 (+pad)][descriptors passes x tiles x 256][pong keys][pong values]
    const size_t hist_bytes = kRsMaxPasses * kRsBins * 4, ticket_bytes = 64;
    const size_t desc_bytes = (size_t)ps.num * tiles * kRsBins * 4;
    const size_t clear_bytes = hist_bytes + ticket_bytes + desc_bytes;
    const size_t pong_off = (clear_bytes + 255) & ~(size_t)255;
    const size_t pong_bytes = ps.num > 1 ? (((size_t)len * 4 + 255) & ~(size_t)255) : 0;
    void *w = nullptr;
    WGB_TRY(workspace_reserve(ctx, 5, pong_off + 2 * pong_bytes, &w));
    unsigned int *hist = reinterpret_cast<unsigned int *>(w);
    unsigned int *tickets = reinterpret_cast<unsigned int *>((char *)w + hist_bytes);
    unsigned int *desc = reinterpret_cast<unsigned int *>((char *)w + hist_bytes + ticket_bytes);
    uint32_t *pong_k = reinterpret_cast<uint32_t *>((char *)w + pong_off);
    uint32_t *pong_v = reinterpret_cast<uint32_t *>((char *)w + pong_off + pong_bytes);
    WGB_CUDA(cudaMemsetAsync(w, 0, clear_bytes, p->stream));
    int hgrid = ctx->prop.multiProcessorCount * 8;
    const uint32_t hneed = (len + 1023) / 1024;
    if ((uint32_t)hgrid > hneed) hgrid = (int)hneed;
    radix_hist_kernel<<<hgrid, 256, 0, p->stream>>>(keys_in, n_dev, len, ps, hist);
    WGB_CUDA(cudaGetLastError());
    count_launch(ctx);
    // the last pass lands in the caller's output; passes alternate between it and the workspace pair (mod.rs:163-221)
    const uint32_t *cur_k = keys_in, *cur_v = vals_in;
    for (uint32_t i = 0; i < ps.num; ++i) {
        const bool to_out = (ps.num - 1 - i) % 2 == 0;
        uint32_t *dk = to_out ? keys_out : pong_k, *dv = to_out ? vals_out : pong_v;
        radix_onesweep_kernel<<<tiles, kRsThreads, 0, p->stream>>>(cur_k, cur_v, dk, dv, n_dev, len, ps.shift[i], ps.mask[i],
                                                                   hist + i * kRsBins, desc + (size_t)i * tiles * kRsBins, tickets + i);
        WGB_CUDA(cudaGetLastError());
        count_launch(ctx);
        cur_k = dk;
        cur_v = dv;
    }
    return WGB_OK;
}

}  // namespace wgb
