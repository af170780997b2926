// gemm_simt.cu — FFMA GEMM for every view the tensor-core path cannot take (strides / offsets that
// TMA rejects, tiny matrices, WGB_F32_SIMT).  Exact f32 products, f32 accumulation.
//
// Semantics follow gemm.wgsl:81-113 (out = m1 * m2) and :116-148 (out = tr(m1) * m2) of
// /root/reference/crates/wgebra/src/linalg/, for any M, N, K (the reference requires multiples of 4),
// any column / matrix stride and offset, batched over size[2] (gemm.rs:126 grid.y).  BT: m2 is N-contiguous
// (a row-major m2, shape.wgsl:49-53), the product is unchanged.
//
// Shape: 128x128 output tile per CTA, K step 8, 256 threads x (8x8) register tile, operands staged in
// shared memory K-major so the inner loop reads two float4 per operand per k; global loads are
// register-prefetched one K step ahead.
#include "common.cuh"

namespace wgb {

namespace {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4, NT = 256;

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

struct SimtArgs {
    const void *a, *b;
    void *c;
    uint32_t M, N, K;
    uint64_t lda, ldb, ldc, sa, sb, sc;
    uint32_t z_base;
    int ep_op;
    const void *e;
    uint64_t lde, se;
};

template <bool TR, bool BT, typename TIn, typename TOut>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(SimtArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const uint32_t t = g.z_base + blockIdx.z;
    const TIn *A = reinterpret_cast<const TIn *>(g.a) + (uint64_t)t * g.sa;
    const TIn *B = reinterpret_cast<const TIn *>(g.b) + (uint64_t)t * g.sb;
    TOut *C = reinterpret_cast<TOut *>(g.c) + (uint64_t)t * g.sc;
    const uint32_t m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    float ra[4], rb[4];
    auto load_tiles = [&](uint32_t k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + NT * e;
            if (!TR) {  // A is M x K, M contiguous
                const uint32_t m = m0 + (idx % BM), k = k0 + (idx / BM);
                ra[e] = (m < g.M && k < g.K) ? to_f32<TIn>(A[(uint64_t)k * g.lda + m]) : 0.f;
            } else {    // A is K x M, K contiguous
                const uint32_t k = k0 + (idx % BK), m = m0 + (idx / BK);
                ra[e] = (m < g.M && k < g.K) ? to_f32<TIn>(A[(uint64_t)m * g.lda + k]) : 0.f;
            }
            if (!BT) {  // B is K x N, K contiguous
                const uint32_t k = k0 + (idx % BK), n = n0 + (idx / BK);
                rb[e] = (n < g.N && k < g.K) ? to_f32<TIn>(B[(uint64_t)n * g.ldb + k]) : 0.f;
            } else {    // B is N x K in memory, N contiguous
                const uint32_t n = n0 + (idx % BN), k = k0 + (idx / BN);
                rb[e] = (n < g.N && k < g.K) ? to_f32<TIn>(B[(uint64_t)k * g.ldb + n]) : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + NT * e;
            if (!TR) As[buf][idx / BM][idx % BM] = ra[e];
            else As[buf][idx % BK][idx / BK] = ra[e];
            if (!BT) Bs[buf][idx % BK][idx / BK] = rb[e];
            else Bs[buf][idx / BN][idx % BN] = rb[e];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const uint32_t nk = (g.K + BK - 1) / BK;
    if (nk > 0) {
        load_tiles(0);
        store_tiles(0);
    }
    __syncthreads();
    for (uint32_t kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][ty * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + ty * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t n = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (n >= g.N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
            if (m < g.M) {
                float val = acc[i][j];
                if (g.ep_op >= 0)
                    val = epilogue_apply<TOut>(g.ep_op, val, reinterpret_cast<const TOut *>(g.e) + (uint64_t)t * g.se + (uint64_t)n * g.lde + m);
                C[(uint64_t)n * g.ldc + m] = from_f32<TOut>(val);
            }
        }
    }
}

// Small-tile variant for batches of small matrices and for mid-sized problems that would leave most SMs idle with 128 x 128
// tiles ("many small batched matmuls").  T x T output tile, (T/4)^2 threads, 4 x 4 register tile, K step 8.
template <int T, bool TR, bool BT, typename TIn, typename TOut>
__global__ void __launch_bounds__((T / 4) * (T / 4)) gemm_simt_small_kernel(SimtArgs g) {
    constexpr int NTS = (T / 4) * (T / 4);    // 256 (T = 64) or 64 (T = 32)
    constexpr int E = T * BK / NTS;           // tile elements each thread stages per operand: 2 or 4
    __shared__ __align__(16) float As[2][BK][T + PAD];
    __shared__ __align__(16) float Bs[2][BK][T + PAD];
    const uint32_t t = g.z_base + blockIdx.z;
    const TIn *A = reinterpret_cast<const TIn *>(g.a) + (uint64_t)t * g.sa;
    const TIn *B = reinterpret_cast<const TIn *>(g.b) + (uint64_t)t * g.sb;
    TOut *C = reinterpret_cast<TOut *>(g.c) + (uint64_t)t * g.sc;
    const uint32_t m0 = blockIdx.x * T, n0 = blockIdx.y * T;
    const int tid = threadIdx.x, tx = tid % (T / 4), ty = tid / (T / 4);

    float ra[E], rb[E];
    auto load_tiles = [&](uint32_t k0) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int idx = tid + NTS * e;
            if (!TR) {
                const uint32_t m = m0 + (idx % T), k = k0 + (idx / T);
                ra[e] = (m < g.M && k < g.K) ? to_f32<TIn>(A[(uint64_t)k * g.lda + m]) : 0.f;
            } else {
                const uint32_t k = k0 + (idx % BK), m = m0 + (idx / BK);
                ra[e] = (m < g.M && k < g.K) ? to_f32<TIn>(A[(uint64_t)m * g.lda + k]) : 0.f;
            }
            if (!BT) {
                const uint32_t k = k0 + (idx % BK), n = n0 + (idx / BK);
                rb[e] = (n < g.N && k < g.K) ? to_f32<TIn>(B[(uint64_t)n * g.ldb + k]) : 0.f;
            } else {
                const uint32_t n = n0 + (idx % T), k = k0 + (idx / T);
                rb[e] = (n < g.N && k < g.K) ? to_f32<TIn>(B[(uint64_t)k * g.ldb + n]) : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int idx = tid + NTS * e;
            if (!TR) As[buf][idx / T][idx % T] = ra[e];
            else As[buf][idx % BK][idx / BK] = ra[e];
            if (!BT) Bs[buf][idx % BK][idx / BK] = rb[e];
            else Bs[buf][idx / T][idx % T] = rb[e];
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const uint32_t nk = (g.K + BK - 1) / BK;
    if (nk > 0) {
        load_tiles(0);
        store_tiles(0);
    }
    __syncthreads();
    for (uint32_t kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][ty * 4]);
            const float av[4] = {a0.x, a0.y, a0.z, a0.w};
            const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t n = n0 + ty * 4 + j;
        if (n >= g.N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t m = m0 + tx * 4 + i;
            if (m < g.M) {
                float val = acc[i][j];
                if (g.ep_op >= 0)
                    val = epilogue_apply<TOut>(g.ep_op, val, reinterpret_cast<const TOut *>(g.e) + (uint64_t)t * g.se + (uint64_t)n * g.lde + m);
                C[(uint64_t)n * g.ldc + m] = from_f32<TOut>(val);
            }
        }
    }
}

template <int T, bool TR, bool BT>
void launch_small_typed(const GemmProblem &p, const SimtArgs &a, dim3 grid, cudaStream_t st) {
    constexpr int NTS = (T / 4) * (T / 4);
    if (p.in_dtype == WGB_F32 && p.out_dtype == WGB_F32) gemm_simt_small_kernel<T, TR, BT, float, float><<<grid, NTS, 0, st>>>(a);
    else if (p.in_dtype == WGB_F32) gemm_simt_small_kernel<T, TR, BT, float, __nv_bfloat16><<<grid, NTS, 0, st>>>(a);
    else if (p.out_dtype == WGB_F32) gemm_simt_small_kernel<T, TR, BT, __nv_bfloat16, float><<<grid, NTS, 0, st>>>(a);
    else gemm_simt_small_kernel<T, TR, BT, __nv_bfloat16, __nv_bfloat16><<<grid, NTS, 0, st>>>(a);
}

template <int T>
void launch_small(const GemmProblem &g, const SimtArgs &a, dim3 grid, cudaStream_t st) {
    if (g.tr) {
        if (g.b_nmajor) launch_small_typed<T, true, true>(g, a, grid, st);
        else launch_small_typed<T, true, false>(g, a, grid, st);
    } else {
        if (g.b_nmajor) launch_small_typed<T, false, true>(g, a, grid, st);
        else launch_small_typed<T, false, false>(g, a, grid, st);
    }
}

// Tile choice: modelled time = (CTAs on the busiest SM) x (padded tile flops) / (relative FFMA efficiency of the tile), with a
// penalty when an SM would hold fewer than 8 warps.  The efficiencies are measured ratios (tools/batched_probe.py).
int pick_simt_tile(const GemmProblem &g, int sms) {
    const char *forced = getenv("WGB_SIMT_TILE");
    if (forced && *forced) {
        const int f = atoi(forced);
        if (f == 32 || f == 64 || f == 128) return f;
    }
    int best = 128;
    double best_cost = 0;
    for (int T : {128, 64, 32}) {
        const double eff = T == 128 ? 1.0 : T == 64 ? 0.95 : 0.75;   // 28 / 27 / 22 TFLOP/s on full tiles (profiles/r1_batched_probe.txt)
        const uint64_t ctas = (uint64_t)((g.M + T - 1) / T) * ((g.N + T - 1) / T) * g.nmats;
        const uint64_t per_sm = (ctas + sms - 1) / sms;
        const double warps = (double)(per_sm > 16 ? 16 : per_sm) * (T == 32 ? 2 : 8);
        const double cost = (double)per_sm * T * T / (eff * (warps < 8 ? warps / 8 : 1.0));
        if (T == 128 || cost < best_cost * 0.95) {   // prefer the larger tile unless the smaller one is clearly better
            best = T;
            best_cost = cost;
        }
    }
    return best;
}

template <bool TR, bool BT>
void launch_typed(const GemmProblem &p, const SimtArgs &a, dim3 grid, cudaStream_t st) {
    if (p.in_dtype == WGB_F32 && p.out_dtype == WGB_F32) gemm_simt_kernel<TR, BT, float, float><<<grid, NT, 0, st>>>(a);
    else if (p.in_dtype == WGB_F32) gemm_simt_kernel<TR, BT, float, __nv_bfloat16><<<grid, NT, 0, st>>>(a);
    else if (p.out_dtype == WGB_F32) gemm_simt_kernel<TR, BT, __nv_bfloat16, float><<<grid, NT, 0, st>>>(a);
    else gemm_simt_kernel<TR, BT, __nv_bfloat16, __nv_bfloat16><<<grid, NT, 0, st>>>(a);
}

}  // namespace

wgb_status launch_gemm_simt(wgb_pass *p, const GemmProblem &g) {
    const size_t es = dtype_size(g.in_dtype), os = dtype_size(g.out_dtype);
    SimtArgs a{};
    a.a = (const char *)g.a + g.a_off * es;
    a.b = (const char *)g.b + g.b_off * es;
    a.c = (char *)g.c + g.c_off * os;
    a.M = g.M; a.N = g.N; a.K = g.K;
    a.lda = g.lda; a.ldb = g.ldb; a.ldc = g.ldc;
    a.sa = g.sa; a.sb = g.sb; a.sc = g.sc;
    a.ep_op = g.ep_op;
    a.e = g.e ? (const char *)g.e + g.e_off * os : nullptr;
    a.lde = g.lde; a.se = g.se;
    const int tile = pick_simt_tile(g, p->ctx->prop.multiProcessorCount);
    const uint32_t gx = (g.M + tile - 1) / tile, gy = (g.N + tile - 1) / tile;
    if (gy > 65535) WGB_FAIL(WGB_ERR_UNSUPPORTED, "gemm (SIMT path): N = %u needs more than 65535 column tiles", g.N);
    for (uint32_t z0 = 0; z0 < g.nmats; z0 += 65535) {
        a.z_base = z0;
        dim3 grid(gx, gy, g.nmats - z0 < 65535 ? g.nmats - z0 : 65535);
        if (tile == 64) launch_small<64>(g, a, grid, p->stream);
        else if (tile == 32) launch_small<32>(g, a, grid, p->stream);
        else if (g.tr) {
            if (g.b_nmajor) launch_typed<true, true>(g, a, grid, p->stream);
            else launch_typed<true, false>(g, a, grid, p->stream);
        } else {
            if (g.b_nmajor) launch_typed<false, true>(g, a, grid, p->stream);
            else launch_typed<false, false>(g, a, grid, p->stream);
        }
        WGB_CUDA(cudaGetLastError());
        count_launch(p->ctx);
    }
    p->last_gemm_path = 1;
    return WGB_OK;
}

}  // namespace wgb
