// gemv.cu — out = m * v and out = tr(m) * v for column-major f32 views.  HBM-bound.
//
// Reference kernels replaced (/root/reference/crates/wgebra/src/linalg/gemv.wgsl):
//   gemv          :68-90   thread owns 4 rows, serial loop over all of K
//   gemv_fast     :29-65   32 threads per 4 rows + tree reduce
//   gemv_tr       :93-115  thread owns 4 columns, serial loop over all rows
//   gemv_tr_fast  :118-154 32 threads per 4 columns + tree reduce
// and the launch logic of gemv.rs:64-137 (grid = [rows, out_ncols, out_nmats]: m is re-read once
// per output column).
//
// B200 design
//   gemv    : CTA = 8 warps owns a 128-row tile x a K range.  A warp reads one 512-byte column
//             segment per instruction (lane = 4 consecutive rows, one 128-bit load), the 8 warps
//             interleave columns, 4 columns in flight per lane; v is staged in shared memory; up to
//             4 output columns are produced per pass over m (m is read once, not once per column).
//             Cross-warp sum through shared memory in warp order.  When M is too small to fill the
//             machine the K range is split across CTAs (grid.y) and the last CTA of a row tile folds
//             the split partials in split order (deterministic; no float atomics).
//   gemv_tr : one warp per column of m (the contiguous axis), 128-bit loads, 4 in flight per lane,
//             warp-shuffle sum; rows split across CTAs when there are too few columns.
// Any offset / stride / length is valid: tiles that are not 16-byte aligned or are ragged take a
// scalar, predicated path inside the same kernels.
//
// Algorithmic bytes per launch: 4 * (M*K + K*C + M*C) per matrix.
#include "common.cuh"
#include "reduce.cuh"

namespace wgb {

static constexpr int kThreads = 256;
static constexpr int kWarps = 8;
static constexpr int kTileRows = 128;
static constexpr int kStage = 512;  // columns of v staged per step
static constexpr int kTStage = 2048;  // gemv_tr: rows of v staged per step (8 KB per output column)

struct GemvArgs {
    const float *m, *v;
    float *out;
    uint32_t M, K;      // out rows, reduction length
    uint32_t C;         // output columns
    uint32_t cchunks;   // ceil(C / NV)
    uint64_t ldm, sm;   // m column / matrix stride
    uint64_t ldv, sv;
    uint64_t ldo, so;
    uint32_t nsplit, chunk;  // reduction split count and length per split
    uint32_t z_base;
    float *partials;
    unsigned int *counters;
    int op;             // < 0: out = result (gemv.wgsl:88); else out = result (op) e — the fused OpAssign step (op_assign.wgsl:14-47)
    const float *e;     // operand of the fused step, indexed like out with its own strides
    uint64_t lde, se;
    // fused Gemv -> Reduce (wgb_gemv_reduce; C == 1, one matrix): `out` is an internal scratch vector, and the last CTA to store its
    // part of it reduces it exactly as wgb_reduce would (reduce.cuh: reduce_in_one_cta) into red_result
    int red_op;                 // < 0: none
    float *red_result;
    unsigned int *red_counter;  // CTAs that have stored their outputs (left at zero)
    uint32_t red_tiles, red_grid;
};

constexpr uint32_t kRedMaxGrid = kGemvReduceMaxGrid;   // per-CTA partials of the emulated reduce launch kept in shared memory

// Tail of both kernels in fused-reduce mode; called by every CTA that has just stored final outputs (all threads of the CTA).
__device__ __forceinline__ void fused_reduce_tail(const GemvArgs &a) {
    __shared__ float red[32];
    __shared__ float vpart[kRedMaxGrid];
    __shared__ bool red_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) red_last = atomicAdd(a.red_counter, 1u) == a.red_tiles - 1;
    __syncthreads();
    if (!red_last) return;
    __threadfence();
    float r;
    switch (a.red_op) {
    case WGB_RED_MIN: r = reduce_in_one_cta<WGB_RED_MIN>(a.out, a.M, a.red_grid, red, vpart); break;
    case WGB_RED_MAX: r = reduce_in_one_cta<WGB_RED_MAX>(a.out, a.M, a.red_grid, red, vpart); break;
    case WGB_RED_SUM: r = reduce_in_one_cta<WGB_RED_SUM>(a.out, a.M, a.red_grid, red, vpart); break;
    case WGB_RED_PROD: r = reduce_in_one_cta<WGB_RED_PROD>(a.out, a.M, a.red_grid, red, vpart); break;
    default: r = reduce_in_one_cta<WGB_RED_SQNORM>(a.out, a.M, a.red_grid, red, vpart); break;
    }
    if (threadIdx.x == 0) {
        *a.red_result = r;
        *a.red_counter = 0u;
    }
}

// The single store of an output element (t = matrix, c = output column, r = row).  With a fused op this is OpAssign's
// `a[i] = a[i] op b[i]` (op_assign.wgsl:41-47) applied to a = the GEMV result the unfused chain would have written to `out`
// first and b = the operand: same operands, same f32 operation, one HBM round trip of `out` less.
__device__ __forceinline__ void store_out(const GemvArgs &a, uint64_t t, uint64_t c, uint64_t r, float s) {
    if (a.op >= 0) {
        const float b = a.e[t * a.se + c * a.lde + r];
        switch (a.op) {
        case WGB_OP_ADD: s = s + b; break;
        case WGB_OP_SUB: s = s - b; break;
        case WGB_OP_MUL: s = s * b; break;
        case WGB_OP_DIV: s = s / b; break;
        default: break;
        }
    }
    a.out[t * a.so + c * a.ldo + r] = s;
}

__device__ __forceinline__ void fma4(float (&acc)[4], const float4 &x, float s) {
    acc[0] = fmaf(x.x, s, acc[0]);
    acc[1] = fmaf(x.y, s, acc[1]);
    acc[2] = fmaf(x.z, s, acc[2]);
    acc[3] = fmaf(x.w, s, acc[3]);
}

template <bool VEC, int NV>
__global__ void __launch_bounds__(kThreads) gemv_n_kernel(GemvArgs a) {
    __shared__ float vs[NV][kStage];
    __shared__ float red[kWarps][NV][kTileRows];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t z = a.z_base + blockIdx.z;
    const uint32_t t = z / a.cchunks, c0 = (z % a.cchunks) * NV;
    const uint32_t nv = min((uint32_t)NV, a.C - c0);
    const uint32_t row0 = blockIdx.x * kTileRows;
    const uint32_t k0 = blockIdx.y * a.chunk;
    const uint32_t k1 = min(a.K, k0 + a.chunk);
    const float *mp = a.m + (uint64_t)t * a.sm;
    const float *vp = a.v + (uint64_t)t * a.sv + (uint64_t)c0 * a.ldv;
    const bool full = VEC && (row0 + kTileRows <= a.M);

    float acc[NV][4];
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;

    for (uint32_t ks = k0; ks < k1; ks += kStage) {
        const uint32_t kn = min((uint32_t)kStage, k1 - ks);
        for (uint32_t idx = threadIdx.x; idx < kn * NV; idx += kThreads) {
            const uint32_t c = idx / kn, kk = idx - c * kn;
            vs[c][kk] = c < nv ? __ldg(vp + (uint64_t)c * a.ldv + ks + kk) : 0.f;
        }
        __syncthreads();
        if (full) {
            const float *base = mp + row0 + 4 * lane + (uint64_t)ks * a.ldm;
            uint32_t kk = w;
            for (; kk + 3 * kWarps < kn; kk += 4 * kWarps) {
                float4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = __ldcs(reinterpret_cast<const float4 *>(base + (uint64_t)(kk + u * kWarps) * a.ldm));
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int c = 0; c < NV; ++c) fma4(acc[c], x[u], vs[c][kk + u * kWarps]);
            }
            for (; kk < kn; kk += kWarps) {
                const float4 x = __ldcs(reinterpret_cast<const float4 *>(base + (uint64_t)kk * a.ldm));
#pragma unroll
                for (int c = 0; c < NV; ++c) fma4(acc[c], x, vs[c][kk]);
            }
        } else {
            // scalar, predicated: lane owns rows row0 + lane + 32 q (still one coalesced 128-byte request per q)
            for (uint32_t kk = w; kk < kn; kk += kWarps) {
                const float *colp = mp + (uint64_t)(ks + kk) * a.ldm;
                float x[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t r = row0 + lane + 32 * q;
                    x[q] = r < a.M ? __ldcs(colp + r) : 0.f;
                }
#pragma unroll
                for (int c = 0; c < NV; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[c][q] = fmaf(x[q], vs[c][kk], acc[c][q]);
            }
        }
        __syncthreads();
    }

    // cross-warp sum, fixed warp order
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) red[w][c][full ? 4 * lane + q : lane + 32 * q] = acc[c][q];
    __syncthreads();
    const bool split = a.nsplit > 1;
    const uint64_t tile_id = (uint64_t)z * gridDim.x + blockIdx.x;
    for (uint32_t idx = threadIdx.x; idx < NV * kTileRows; idx += kThreads) {
        const uint32_t c = idx / kTileRows, rl = idx % kTileRows;
        float s = 0.f;
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) s += red[ww][c][rl];
        if (!split) {
            if (row0 + rl < a.M && c < nv) store_out(a, t, c0 + c, row0 + rl, s);
        } else {
            a.partials[((tile_id * a.nsplit + blockIdx.y) * NV + c) * kTileRows + rl] = s;
        }
    }
    if (!split) {
        if (NV == 1 && a.red_op >= 0) fused_reduce_tail(a);
        return;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(a.counters + tile_id, 1u);
        is_last = ticket == a.nsplit - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (uint32_t idx = threadIdx.x; idx < NV * kTileRows; idx += kThreads) {
        const uint32_t c = idx / kTileRows, rl = idx % kTileRows;
        float s = 0.f;
        for (uint32_t y = 0; y < a.nsplit; ++y) s += __ldcg(a.partials + ((tile_id * a.nsplit + y) * NV + c) * kTileRows + rl);
        if (row0 + rl < a.M && c < nv) store_out(a, t, c0 + c, row0 + rl, s);
    }
    if (threadIdx.x == 0) a.counters[tile_id] = 0u;
    if (NV == 1 && a.red_op >= 0) fused_reduce_tail(a);
}

__device__ __forceinline__ float dot4(const float4 &x, const float4 &y, float acc) {
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc);
    return fmaf(x.w, y.w, acc);
}

// out[j] = sum_i m[i, j] * v[i]; here a.M = number of outputs (columns of m), a.K = rows of m.
template <bool VEC, int NV>
__global__ void __launch_bounds__(kThreads) gemv_t_kernel(GemvArgs a) {
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t z = a.z_base + blockIdx.z;
    const uint32_t t = z / a.cchunks, c0 = (z % a.cchunks) * NV;
    const uint32_t nv = min((uint32_t)NV, a.C - c0);
    const uint32_t j = blockIdx.x * kWarps + w;
    const bool valid = j < a.M;
    const uint32_t r0 = blockIdx.y * a.chunk;
    const uint32_t r1 = min(a.K, r0 + a.chunk);
    const float *col = a.m + (uint64_t)t * a.sm + (uint64_t)(valid ? j : 0) * a.ldm;
    const float *vp = a.v + (uint64_t)t * a.sv + (uint64_t)c0 * a.ldv;

    float acc[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] = 0.f;

    if (VEC) {
        // v is staged in shared memory kTStage rows at a time (all 8 warps = 8 columns of m consume the same rows), so the only
        // global loads in the loop are the 128-bit streaming reads of m, eight in flight per lane (4 KB per warp).
        __shared__ float4 vs[NV][kTStage / 4];
        const uint32_t nvec = (r1 - r0) >> 2;  // r0 is a multiple of 4 (chunk is)
        const float *cb = col + r0;
        const float *vb = vp + r0;
        for (uint32_t s0 = 0; s0 < nvec; s0 += kTStage / 4) {
            const uint32_t ns = min((uint32_t)kTStage / 4, nvec - s0);
            __syncthreads();   // the previous stage has been consumed
            for (uint32_t i = threadIdx.x; i < ns; i += kThreads)
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    if (c < nv) vs[c][i] = __ldg(reinterpret_cast<const float4 *>(vb + (uint64_t)c * a.ldv) + s0 + i);
            __syncthreads();
            if (valid) {
                const float4 *xb = reinterpret_cast<const float4 *>(cb) + s0;
                uint32_t i = lane;
                for (; i + 224 < ns; i += 256) {
                    float4 x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = __ldcs(xb + i + 32 * u);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int c = 0; c < NV; ++c)
                            if (c < nv) acc[c] = dot4(x[u], vs[c][i + 32 * u], acc[c]);
                }
                for (; i < ns; i += 32) {
                    const float4 x = __ldcs(xb + i);
#pragma unroll
                    for (int c = 0; c < NV; ++c)
                        if (c < nv) acc[c] = dot4(x, vs[c][i], acc[c]);
                }
            }
        }
        if (valid)
            for (uint32_t r = r0 + (nvec << 2) + lane; r < r1; r += 32) {
                const float x = col[r];
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    if (c < nv) acc[c] = fmaf(x, vp[(uint64_t)c * a.ldv + r], acc[c]);
            }
    } else if (valid) {
        {
            uint32_t r = r0 + lane;
            for (; r + 96 < r1; r += 128) {
                float x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = __ldcs(col + r + 32 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int c = 0; c < NV; ++c)
                        if (c < nv) acc[c] = fmaf(x[u], __ldg(vp + (uint64_t)c * a.ldv + r + 32 * u), acc[c]);
            }
            for (; r < r1; r += 32) {
                const float x = __ldcs(col + r);
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    if (c < nv) acc[c] = fmaf(x, __ldg(vp + (uint64_t)c * a.ldv + r), acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);

    const bool split = a.nsplit > 1;
    if (!split) {
        if (valid && lane == 0)
#pragma unroll
            for (int c = 0; c < NV; ++c)
                if (c < nv) store_out(a, t, c0 + c, j, acc[c]);
        if (NV == 1 && a.red_op >= 0) fused_reduce_tail(a);
        return;
    }
    const uint64_t tile_id = (uint64_t)z * gridDim.x + blockIdx.x;
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < NV; ++c) a.partials[((tile_id * a.nsplit + blockIdx.y) * kWarps + w) * NV + c] = acc[c];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(a.counters + tile_id, 1u);
        is_last = ticket == a.nsplit - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < kWarps * NV) {
        const uint32_t ww = threadIdx.x / NV, c = threadIdx.x % NV;
        const uint32_t jj = blockIdx.x * kWarps + ww;
        float s = 0.f;
        for (uint32_t y = 0; y < a.nsplit; ++y) s += __ldcg(a.partials + ((tile_id * a.nsplit + y) * kWarps + ww) * NV + c);
        if (jj < a.M && c < nv) store_out(a, t, c0 + c, jj, s);
    }
    if (threadIdx.x == 0) a.counters[tile_id] = 0u;
    if (NV == 1 && a.red_op >= 0) fused_reduce_tail(a);
}

template <bool TR, bool VEC, int NV>
static void launch_one(const GemvArgs &a, dim3 grid, cudaStream_t st) {
    if (TR) gemv_t_kernel<VEC, NV><<<grid, kThreads, 0, st>>>(a);
    else gemv_n_kernel<VEC, NV><<<grid, kThreads, 0, st>>>(a);
}

template <bool TR>
static void launch_sel(bool vec, int nv, const GemvArgs &a, dim3 grid, cudaStream_t st) {
    if (vec) {
        if (nv == 1) launch_one<TR, true, 1>(a, grid, st);
        else if (nv == 2) launch_one<TR, true, 2>(a, grid, st);
        else launch_one<TR, true, 4>(a, grid, st);
    } else {
        if (nv == 1) launch_one<TR, false, 1>(a, grid, st);
        else if (nv == 2) launch_one<TR, false, 2>(a, grid, st);
        else launch_one<TR, false, 4>(a, grid, st);
    }
}

// Resident CTAs per SM of the gemv_tr instantiation a launch will use (cached: the occupancy query costs microseconds).
static int gemv_t_resident(bool vec, int nv) {
    static int cache[2][3] = {};
    const int vi = vec ? 1 : 0, ni = nv == 1 ? 0 : (nv == 2 ? 1 : 2);
    if (cache[vi][ni] == 0) {
        int n = 0;
        cudaError_t e;
        if (vec) e = nv == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_t_kernel<true, 1>, kThreads, 0)
                   : nv == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_t_kernel<true, 2>, kThreads, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_t_kernel<true, 4>, kThreads, 0);
        else e = nv == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_t_kernel<false, 1>, kThreads, 0)
                 : nv == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_t_kernel<false, 2>, kThreads, 0)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_t_kernel<false, 4>, kThreads, 0);
        if (e != cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = 4;
        }
        cache[vi][ni] = n;
    }
    return cache[vi][ni];
}

__global__ void gemv_empty_k_kernel(GemvArgs a, uint32_t nmats) {   // K == 0 with a fused op: out = 0 (op) operand
    const uint64_t total = (uint64_t)a.M * a.C * nmats;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i % a.M, c = (i / a.M) % a.C, t = i / ((uint64_t)a.M * a.C);
        store_out(a, t, c, r, 0.0f);
    }
}

int reduce_grid_for(wgb_ctx *ctx, int op, uint64_t n);   // level1.cu

wgb_status launch_gemv(wgb_pass *p, bool tr, float *out, const wgb_view_shape &so, const float *m,
                       const wgb_view_shape &sm, const float *v, const wgb_view_shape &sv, int op, const float *operand,
                       const wgb_view_shape *se, int red_op, float *red_result) {
    wgb_ctx *ctx = p->ctx;
    GemvArgs a{};
    a.op = op;
    a.red_op = -1;
    if (op >= 0) {
        a.e = operand + se->offset;
        a.lde = se->stride;
        a.se = se->stride_mat;
    }
    a.m = m + sm.offset;
    a.v = v + sv.offset;
    a.out = out + so.offset;
    a.M = so.size[0];
    a.K = sv.size[0];
    a.C = so.size[1];
    const uint32_t nmats = so.size[2];
    a.ldm = sm.stride; a.sm = sm.stride_mat;
    a.ldv = sv.stride; a.sv = sv.stride_mat;
    a.ldo = so.stride; a.so = so.stride_mat;
    const int nv = a.C == 1 ? 1 : (a.C == 2 ? 2 : 4);
    a.cchunks = (a.C + nv - 1) / nv;
    const uint64_t zdim = (uint64_t)a.cchunks * nmats;

    if (a.K == 0) {
        // empty reduction: the reference's loops do not execute and it stores zeros (gemv.wgsl:74,88)
        if (op >= 0) {
            gemv_empty_k_kernel<<<ctx->prop.multiProcessorCount, 256, 0, p->stream>>>(a, nmats);
            WGB_CUDA(cudaGetLastError());
            count_launch(ctx);
            return WGB_OK;
        }
        for (uint32_t t = 0; t < nmats; ++t)
            for (uint32_t c = 0; c < a.C; ++c)
                WGB_CUDA(cudaMemsetAsync(a.out + (uint64_t)t * a.so + (uint64_t)c * a.ldo, 0, (size_t)a.M * 4, p->stream));
        return WGB_OK;
    }

    auto al16 = [](const void *ptr) { return ((uintptr_t)ptr & 15u) == 0; };
    bool vec;
    uint64_t tiles_x;
    if (!tr) {
        vec = al16(a.m) && a.ldm % 4 == 0 && (nmats == 1 || a.sm % 4 == 0);
        tiles_x = (a.M + kTileRows - 1) / kTileRows;
    } else {
        vec = al16(a.m) && a.ldm % 4 == 0 && (nmats == 1 || a.sm % 4 == 0) && al16(a.v) && (a.C == 1 || a.ldv % 4 == 0) &&
              (nmats == 1 || a.sv % 4 == 0);
        tiles_x = (a.M + kWarps - 1) / kWarps;
    }
    // Split the reduction axis when the output axis alone cannot fill the machine: aim at ~16 CTAs per SM in total.
    // (A split count tuned to a whole number of resident waves — 13 splits for 65536 x 4096 gemv_tr — measured *slower*,
    // 5742 vs 6072-6155 GB/s: more partials to fold and shorter streams per warp outweigh the fuller last wave.)
    const uint64_t target = (uint64_t)ctx->prop.multiProcessorCount * 16;
    const uint32_t min_chunk = tr ? 2048 : 256;
    uint64_t nsplit = 1;
    if (tiles_x * zdim < target) {
        nsplit = (target + tiles_x * zdim - 1) / (tiles_x * zdim);
        const uint64_t max_split = (a.K + min_chunk - 1) / min_chunk;
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit > 65535) nsplit = 65535;
    }
    if (tr && nsplit > 1) {
        // gemv_tr CTAs all stream the same number of bytes, so the launch runs in waves of (SMs x resident CTAs): among the split
        // counts up to twice the first guess, take the one whose last wave is fullest (65536 x 4096: 8 splits = 6.92 waves on 4
        // resident CTAs per SM instead of 5 splits = 4.32)
        const uint64_t resident = (uint64_t)ctx->prop.multiProcessorCount * (uint64_t)gemv_t_resident(vec, nv);
        const uint64_t max_split = (a.K + min_chunk - 1) / min_chunk;
        double best = 0.0;
        uint64_t best_n = nsplit;
        for (uint64_t n = nsplit; n <= 2 * nsplit + 1 && n <= max_split && n <= 65535; ++n) {
            const double waves = (double)(tiles_x * zdim * n) / (double)resident;
            const double eff = waves / (double)(uint64_t)(waves + 0.999999);
            if (eff > best + 0.02) {
                best = eff;
                best_n = n;
            }
        }
        nsplit = best_n;
    }
    uint32_t chunk = (uint32_t)(((uint64_t)a.K + nsplit - 1) / nsplit);
    chunk = (chunk + 31u) & ~31u;
    nsplit = ((uint64_t)a.K + chunk - 1) / chunk;
    if (nsplit > 1) {
        const uint64_t per_tile = tr ? (uint64_t)kWarps * nv : (uint64_t)nv * kTileRows;
        const uint64_t need = tiles_x * zdim * nsplit * per_tile;
        if (tiles_x * zdim > (1u << 20) || need > ((uint64_t)64 << 20)) {
            nsplit = 1;
            chunk = (a.K + 31u) & ~31u;
        } else {
            WGB_TRY(scratch_reserve(ctx, need, tiles_x * zdim + 1));   // (+ 1: the last slot is the fused-reduce ticket)
        }
    }
    a.nsplit = (uint32_t)nsplit;
    a.chunk = chunk;
    a.partials = ctx->scratch.partials;
    a.counters = ctx->scratch.counters;
    if (red_op >= 0) {
        // fused Gemv -> Reduce: only reached with C == 1, one matrix, K > 0 and a grid the tail can emulate (checked by the caller)
        a.red_op = red_op;
        a.red_result = red_result;
        a.red_counter = ctx->scratch.counters + (ctx->scratch.n_counters - 1);
        a.red_tiles = (uint32_t)tiles_x;
        a.red_grid = (uint32_t)reduce_grid_for(ctx, red_op, a.M);
    }
    if (tiles_x > 0x7fffffffull) WGB_FAIL(WGB_ERR_UNSUPPORTED, "gemv: too many row tiles");

    for (uint64_t z0 = 0; z0 < zdim; z0 += 65535) {
        const uint32_t zn = (uint32_t)((zdim - z0) < 65535 ? (zdim - z0) : 65535);
        a.z_base = (uint32_t)z0;
        dim3 grid((unsigned)tiles_x, (unsigned)nsplit, zn);
        if (tr) launch_sel<true>(vec, nv, a, grid, p->stream);
        else launch_sel<false>(vec, nv, a, grid, p->stream);
        WGB_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
    return WGB_OK;
}

}  // namespace wgb
