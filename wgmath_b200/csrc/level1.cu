// level1.cu — the HBM-bound vector kernels: op_assign, reduce / dot, column reduce, seeded fill.
//
// Reference kernels replaced (paths under /root/reference/crates/wgebra/src/linalg/):
//   op_assign.wgsl:41-47  one f32 per thread, 64-thread groups
//   reduce.wgsl:68-96     ONE workgroup of 128 threads for the whole vector (reduce.rs:112)
// B200 design: 128-bit loads/stores wherever the view's start is 16-byte aligned (a scalar
// head/tail handles the rest, so every offset / length is valid), 4 independent 16-byte
// requests in flight per thread, grid sized in multiples of the SM count, and for reductions
// warp-shuffle partials -> one partial per CTA -> the last CTA to finish folds the partials in
// index order (deterministic, one launch, no float atomics).
//
// Algorithmic bytes: op_assign 12 B/elem (8 for Copy), reduce 4 B/elem, dot 8 B/elem,
// column reduce 4 B/elem.
#include "common.cuh"
#include "reduce.cuh"

namespace wgb {

static constexpr int kThreads = 256;
static_assert(kThreads == kRedThreads, "reduce.cuh assumes the level-1 CTA size");

__device__ __forceinline__ float4 ld_stream4(const float *p) {
    return __ldcs(reinterpret_cast<const float4 *>(p));
}

// ---------------------------------------------------------------------------- op_assign
template <int OP>
__device__ __forceinline__ float apply_op(float a, float b) {
    if (OP == WGB_OP_ADD) return a + b;   // op_assign.wgsl:14-16
    if (OP == WGB_OP_SUB) return a - b;   // :18-20
    if (OP == WGB_OP_MUL) return a * b;   // :22-24
    if (OP == WGB_OP_DIV) return a / b;   // :26-28 (IEEE division: nvcc default -prec-div=true)
    return b;                             // :36-38 copy
}

template <int OP, bool B_ALIGNED>
__global__ void __launch_bounds__(kThreads, 4) op_assign_kernel(float *__restrict__ a, const float *__restrict__ b,
                                                             uint64_t n, uint32_t head) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    // scalar head (until `a` is 16-byte aligned) and tail
    const uint64_t nvec = (n - head) >> 2;
    const uint64_t tail0 = head + (nvec << 2);
    if (tid < head) a[tid] = apply_op<OP>(OP == WGB_OP_COPY ? 0.f : a[tid], b[tid]);
    if (tail0 + tid < n) a[tail0 + tid] = apply_op<OP>(OP == WGB_OP_COPY ? 0.f : a[tail0 + tid], b[tail0 + tid]);

    // 2 x (a, b) 16-byte requests in flight per thread; pointer-bumping keeps the loop at ~32 registers
    float4 *pa = reinterpret_cast<float4 *>(a + head) + tid;
    const float *pb = b + head + 4 * tid;
    const uint64_t step = nthreads;
    auto ldb = [](const float *q) {
        if (B_ALIGNED) return ld_stream4(q);
        return make_float4(__ldcs(q), __ldcs(q + 1), __ldcs(q + 2), __ldcs(q + 3));
    };
    auto combine = [](const float4 &x, const float4 &y) {
        float4 r;
        r.x = apply_op<OP>(x.x, y.x);
        r.y = apply_op<OP>(x.y, y.y);
        r.z = apply_op<OP>(x.z, y.z);
        r.w = apply_op<OP>(x.w, y.w);
        return r;
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    uint64_t i = tid;
    for (; i + step < nvec; i += 2 * step) {
        const float4 x0 = OP != WGB_OP_COPY ? __ldcs(pa) : zero;
        const float4 x1 = OP != WGB_OP_COPY ? __ldcs(pa + step) : zero;
        const float4 y0 = ldb(pb);
        const float4 y1 = ldb(pb + 4 * step);
        __stcs(pa, combine(x0, y0));
        __stcs(pa + step, combine(x1, y1));
        pa += 2 * step;
        pb += 8 * step;
    }
    if (i < nvec) {
        const float4 x0 = OP != WGB_OP_COPY ? __ldcs(pa) : zero;
        __stcs(pa, combine(x0, ldb(pb)));
    }
}

// Resident CTAs per SM for a kernel (queried once): grids are sized to exactly one wave of resident CTAs, each
// thread grid-strides over the rest, so there is no partial last wave and no wave hand-over bubble.
template <typename K>
static int resident_ctas(K kernel) {
    static std::mutex mu;
    static std::unordered_map<const void *, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find((const void *)kernel);
    if (it != cache.end()) return it->second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, 0) != cudaSuccess || n < 1) n = 4;
    cache[(const void *)kernel] = n;
    return n;
}
static int grid_for(wgb_ctx *ctx, uint64_t work_items_per_thread_unit, int ctas_per_sm) {
    const uint64_t max_ctas = (uint64_t)ctx->prop.multiProcessorCount * ctas_per_sm;
    uint64_t need = (work_items_per_thread_unit + kThreads - 1) / kThreads;
    if (need < 1) need = 1;
    return (int)(need < max_ctas ? need : max_ctas);
}

wgb_status launch_op_assign(wgb_pass *p, int op, float *a, const float *b, uint64_t n) {
    const uint32_t mis = (uint32_t)(((uintptr_t)a & 15u) >> 2);
    uint32_t head = mis ? 4u - mis : 0u;
    if (((uintptr_t)a & 3u) != 0) WGB_FAIL(WGB_ERR_INVALID, "op_assign: buffer is not 4-byte aligned");
    if (head > n) head = (uint32_t)n;
    const bool b_aligned = (((uintptr_t)(b + head)) & 15u) == 0;
    const uint64_t nvec = (n - head) >> 2;
    // grid: enough CTAs for >= 2 vectors per thread, capped at exactly one wave of resident CTAs
#define LAUNCH(OPV)                                                                                       \
    case OPV:                                                                                             \
        if (b_aligned) {                                                                                  \
            const int grid = grid_for(p->ctx, (nvec + 1) / 2 + 8, resident_ctas(op_assign_kernel<OPV, true>));   \
            op_assign_kernel<OPV, true><<<grid, kThreads, 0, p->stream>>>(a, b, n, head);                 \
        } else {                                                                                          \
            const int grid = grid_for(p->ctx, (nvec + 1) / 2 + 8, resident_ctas(op_assign_kernel<OPV, false>));  \
            op_assign_kernel<OPV, false><<<grid, kThreads, 0, p->stream>>>(a, b, n, head);                \
        }                                                                                                 \
        break;
    switch (op) {
        LAUNCH(WGB_OP_ADD)
        LAUNCH(WGB_OP_SUB)
        LAUNCH(WGB_OP_MUL)
        LAUNCH(WGB_OP_DIV)
        LAUNCH(WGB_OP_COPY)
        default: WGB_FAIL(WGB_ERR_INVALID, "op_assign: unknown op");
    }
#undef LAUNCH
    WGB_CUDA(cudaGetLastError());
    count_launch(p->ctx);
    return WGB_OK;
}

// ---------------------------------------------------------------------------- reduce / dot (tree: reduce.cuh)
template <int OP, bool Y_ALIGNED>
__global__ void __launch_bounds__(kThreads, 4) reduce_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                          uint64_t n, float *__restrict__ partials,
                                                          unsigned int *__restrict__ counter, float *__restrict__ result) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    float v = thread_partial<OP, Y_ALIGNED>(x, y, n, tid, nthreads);
    v = block_reduce<OP>(v, red);
    if (gridDim.x == 1) {
        if (threadIdx.x == 0) *result = v;
        return;
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        is_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fold the per-CTA partials in index order: the result does not depend on which CTA finished last
        float acc = red_init<OP>();
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) acc = red_comb<OP>(acc, __ldcg(partials + i));
        acc = block_reduce<OP>(acc, red);
        if (threadIdx.x == 0) {
            *result = acc;
            *counter = 0u;  // leave the ticket at zero for the next launch
        }
    }
}

// The grid launch_reduce uses for `n` elements (the fused Gemv -> Reduce tail reproduces exactly this launch).
int reduce_grid_for(wgb_ctx *ctx, int op, uint64_t n) {
    const uint64_t nvec = n >> 2;
    int occ;
    switch (op) {
        case WGB_RED_MIN: occ = resident_ctas(reduce_kernel<WGB_RED_MIN, true>); break;
        case WGB_RED_MAX: occ = resident_ctas(reduce_kernel<WGB_RED_MAX, true>); break;
        case WGB_RED_SUM: occ = resident_ctas(reduce_kernel<WGB_RED_SUM, true>); break;
        case WGB_RED_PROD: occ = resident_ctas(reduce_kernel<WGB_RED_PROD, true>); break;
        case WGB_RED_SQNORM: occ = resident_ctas(reduce_kernel<WGB_RED_SQNORM, true>); break;
        default: occ = resident_ctas(reduce_kernel<5, true>); break;
    }
    int grid = grid_for(ctx, (nvec + 7) / 8 + 1, occ);  // >= 8 vectors per thread before adding CTAs; one resident wave at most
    if ((size_t)grid > ctx->scratch.partials_floats) grid = (int)ctx->scratch.partials_floats;
    return grid;
}

wgb_status launch_reduce(wgb_pass *p, int op, const float *x, const float *y, uint64_t n, float *result) {
    if (((uintptr_t)x & 3u) != 0) WGB_FAIL(WGB_ERR_INVALID, "reduce: buffer is not 4-byte aligned");
    wgb_ctx *ctx = p->ctx;
    const int grid = reduce_grid_for(ctx, op, n);
    const uint32_t mis = (uint32_t)(((uintptr_t)x & 15u) >> 2);
    const uint64_t head = mis ? 4u - mis : 0u;
    const bool y_aligned = y && ((((uintptr_t)(y + head)) & 15u) == 0);
    float *partials = ctx->scratch.partials;
    unsigned int *counter = ctx->scratch.counters;  // slot 0
#define LAUNCH(OPV)                                                                                               \
    case OPV: reduce_kernel<OPV, true><<<grid, kThreads, 0, p->stream>>>(x, y, n, partials, counter, result); break;
    switch (op) {
        LAUNCH(WGB_RED_MIN)
        LAUNCH(WGB_RED_MAX)
        LAUNCH(WGB_RED_SUM)
        LAUNCH(WGB_RED_PROD)
        LAUNCH(WGB_RED_SQNORM)
        case 5:
            if (y_aligned) reduce_kernel<5, true><<<grid, kThreads, 0, p->stream>>>(x, y, n, partials, counter, result);
            else reduce_kernel<5, false><<<grid, kThreads, 0, p->stream>>>(x, y, n, partials, counter, result);
            break;
        default: WGB_FAIL(WGB_ERR_INVALID, "reduce: unknown op");
    }
#undef LAUNCH
    WGB_CUDA(cudaGetLastError());
    count_launch(ctx);
    return WGB_OK;
}

// ---------------------------------------------------------------------------- column reduce
// One CTA per column (grid-stride over columns): the reference needs one Reduce dispatch per
// GpuMatrix::column(j) (tensor.rs:574-585); here every column of every matrix is reduced in one launch.
template <int OP>
__global__ void __launch_bounds__(kThreads) reduce_columns_kernel(const float *__restrict__ m, wgb_view_shape s,
                                                                  float *__restrict__ out, uint64_t ncols_total) {
    __shared__ float red[32];
    for (uint64_t c = blockIdx.x; c < ncols_total; c += gridDim.x) {
        const uint64_t t = c / s.size[1], j = c % s.size[1];
        const float *col = m + (uint64_t)s.offset + t * (uint64_t)s.stride_mat + j * (uint64_t)s.stride;
        float v = thread_partial<OP, true>(col, nullptr, s.size[0], threadIdx.x, blockDim.x);
        v = block_reduce<OP>(v, red);
        if (threadIdx.x == 0) out[c] = v;
    }
}
// Short columns: one warp per column.
template <int OP>
__global__ void __launch_bounds__(kThreads) reduce_columns_warp_kernel(const float *__restrict__ m, wgb_view_shape s,
                                                                       float *__restrict__ out, uint64_t ncols_total) {
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < ncols_total; c += warps) {
        const uint64_t t = c / s.size[1], j = c % s.size[1];
        const float *col = m + (uint64_t)s.offset + t * (uint64_t)s.stride_mat + j * (uint64_t)s.stride;
        float v = thread_partial<OP, true>(col, nullptr, s.size[0], lane, 32);
        v = warp_reduce<OP>(v);
        if (lane == 0) out[c] = v;
    }
}

wgb_status launch_reduce_columns(wgb_pass *p, int op, const float *m, const wgb_view_shape &s, float *out) {
    const uint64_t ncols = (uint64_t)s.size[1] * s.size[2];
    const bool per_warp = s.size[0] < 2048;
    const uint64_t ctas_needed = per_warp ? (ncols * 32 + kThreads - 1) / kThreads : ncols;
    const uint64_t cap = (uint64_t)p->ctx->prop.multiProcessorCount * 8;
    const int grid = (int)(ctas_needed < cap ? ctas_needed : cap);
#define LAUNCH(OPV)                                                                                         \
    case OPV:                                                                                               \
        if (per_warp) reduce_columns_warp_kernel<OPV><<<grid, kThreads, 0, p->stream>>>(m, s, out, ncols);  \
        else reduce_columns_kernel<OPV><<<grid, kThreads, 0, p->stream>>>(m, s, out, ncols);                \
        break;
    switch (op) {
        LAUNCH(WGB_RED_MIN)
        LAUNCH(WGB_RED_MAX)
        LAUNCH(WGB_RED_SUM)
        LAUNCH(WGB_RED_PROD)
        LAUNCH(WGB_RED_SQNORM)
        default: WGB_FAIL(WGB_ERR_INVALID, "reduce_columns: unknown op");
    }
#undef LAUNCH
    WGB_CUDA(cudaGetLastError());
    count_launch(p->ctx);
    return WGB_OK;
}

// ---------------------------------------------------------------------------- seeded fill
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) fill_uniform_kernel(T *__restrict__ base, wgb_view_shape s, uint64_t seed,
                                                                uint32_t row0, uint32_t col0) {
    const uint64_t rows = s.size[0];
    const uint64_t total = rows * s.size[1] * s.size[2];
    const uint64_t key0 = seed * 0xD1342543DE82EF95ull;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = e % rows;
        const uint64_t jt = e / rows;  // j + t * ncols
        const uint64_t j = jt % s.size[1], t = jt / s.size[1];
        const uint64_t key = key0 ^ (((uint64_t)(col0 + jt) << 32) | (uint64_t)(row0 + i));
        const float v = (float)(splitmix64(key) >> 40) * 5.9604644775390625e-8f;  // 2^-24
        const uint64_t idx = (uint64_t)s.offset + t * (uint64_t)s.stride_mat + j * (uint64_t)s.stride + i;
        if (sizeof(T) == 4) reinterpret_cast<float *>(base)[idx] = v;
        else reinterpret_cast<__nv_bfloat16 *>(base)[idx] = __float2bfloat16_rn(v);
    }
}

wgb_status launch_fill_uniform(wgb_pass *p, void *base, const wgb_view_shape &s, wgb_dtype dt, uint64_t seed,
                               uint32_t row0, uint32_t col0) {
    const uint64_t total = (uint64_t)s.size[0] * s.size[1] * s.size[2];
    const int grid = grid_for(p->ctx, (total + 3) / 4, 16);
    if (dt == WGB_F32) fill_uniform_kernel<float><<<grid, kThreads, 0, p->stream>>>((float *)base, s, seed, row0, col0);
    else fill_uniform_kernel<__nv_bfloat16><<<grid, kThreads, 0, p->stream>>>((__nv_bfloat16 *)base, s, seed, row0, col0);
    WGB_CUDA(cudaGetLastError());
    count_launch(p->ctx);
    return WGB_OK;
}

}  // namespace wgb
