// reduce.cuh — the reduction tree of wgb_reduce / wgb_dot as device templates, shared by level1.cu (reduce_kernel) and gemv.cu
// (the fused Gemv -> Reduce tail, which must reproduce reduce_kernel's result bit for bit).
// Reference: /root/reference/crates/wgebra/src/linalg/reduce.rs:30-60 (init values and combine functions), reduce.wgsl:59-96.
#pragma once
#include "common.cuh"

namespace wgb {

static constexpr int kRedThreads = 256;

// COHERENT = false: streaming loads (ld.global.cs).  COHERENT = true: ld.global.cg, for data other CTAs of the SAME launch wrote.
template <bool COHERENT>
__device__ __forceinline__ float4 red_ld4(const float *p) {
    return COHERENT ? __ldcg(reinterpret_cast<const float4 *>(p)) : __ldcs(reinterpret_cast<const float4 *>(p));
}
template <bool COHERENT>
__device__ __forceinline__ float red_ld1(const float *p) {
    return COHERENT ? __ldcg(p) : *p;
}

// OP 0..4 = wgb_reduce_op, 5 = dot.
template <int OP>
__device__ __forceinline__ float red_init() {   // reduce.rs:30-38 / reduce.wgsl:32-46
    if (OP == WGB_RED_MIN) return 3.4e38f;
    if (OP == WGB_RED_MAX) return -3.4e38f;
    if (OP == WGB_RED_PROD) return 1.0f;
    return 0.0f;
}
template <int OP>
__device__ __forceinline__ float red_elem(float acc, float x, float y) {   // workspace_fn
    if (OP == WGB_RED_MIN) return fminf(acc, x);
    if (OP == WGB_RED_MAX) return fmaxf(acc, x);
    if (OP == WGB_RED_SUM) return acc + x;
    if (OP == WGB_RED_PROD) return acc * x;
    if (OP == WGB_RED_SQNORM) return fmaf(x, x, acc);
    return fmaf(x, y, acc);  // dot
}
template <int OP>
__device__ __forceinline__ float red_comb(float a, float b) {   // reduce_fn (SqNorm and dot combine by sum)
    if (OP == WGB_RED_MIN) return fminf(a, b);
    if (OP == WGB_RED_MAX) return fmaxf(a, b);
    if (OP == WGB_RED_PROD) return a * b;
    return a + b;
}

template <int OP>
__device__ __forceinline__ float warp_reduce(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_comb<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide combine; result valid in thread 0.
template <int OP>
__device__ __forceinline__ float block_reduce(float v, float *smem /* >= 32 floats */) {
    v = warp_reduce<OP>(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? smem[lane] : red_init<OP>();
        v = warp_reduce<OP>(v);
    }
    __syncthreads();
    return v;
}

// Accumulate x[0..n) (and y for dot) into a per-thread partial, cooperatively over `nthreads` threads.
template <int OP, bool Y_ALIGNED, bool COHERENT = false>
__device__ __forceinline__ float thread_partial(const float *__restrict__ x, const float *__restrict__ y, uint64_t n,
                                                uint64_t tid, uint64_t nthreads) {
    constexpr bool DOT = OP == 5;
    const uint32_t mis = (uint32_t)(((uintptr_t)x & 15u) >> 2);
    uint64_t head = mis ? 4u - mis : 0u;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) >> 2;
    const uint64_t tail0 = head + (nvec << 2);
    float acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = red_init<OP>();
    if (tid < head) acc[0] = red_elem<OP>(acc[0], red_ld1<COHERENT>(x + tid), DOT ? y[tid] : 0.f);
    if (tail0 + tid < n) acc[1] = red_elem<OP>(acc[1], red_ld1<COHERENT>(x + tail0 + tid), DOT ? y[tail0 + tid] : 0.f);
    // pointer-bumping main loop: U independent 16-byte requests per stream in flight per thread
    constexpr int U = DOT ? 2 : 4;
    const float *px = x + head + 4 * tid;
    const float *py = DOT ? y + head + 4 * tid : nullptr;
    const uint64_t step4 = 4 * nthreads;
    auto ldy = [](const float *q) {
        if (Y_ALIGNED) return red_ld4<false>(q);
        return make_float4(__ldcs(q), __ldcs(q + 1), __ldcs(q + 2), __ldcs(q + 3));
    };
    auto fold = [&](const float4 &a, const float4 &b) {
        acc[0] = red_elem<OP>(acc[0], a.x, b.x);
        acc[1] = red_elem<OP>(acc[1], a.y, b.y);
        acc[2] = red_elem<OP>(acc[2], a.z, b.z);
        acc[3] = red_elem<OP>(acc[3], a.w, b.w);
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    uint64_t i = tid;
    for (; i + (U - 1) * nthreads < nvec; i += U * nthreads) {
        float4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = red_ld4<COHERENT>(px + u * step4);
            b[u] = DOT ? ldy(py + u * step4) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) fold(a[u], b[u]);
        px += U * step4;
        if (DOT) py += U * step4;
    }
    for (; i < nvec; i += nthreads) {
        fold(red_ld4<COHERENT>(px), DOT ? ldy(py) : zero);
        px += step4;
        if (DOT) py += step4;
    }
    return red_comb<OP>(red_comb<OP>(acc[0], acc[1]), red_comb<OP>(acc[2], acc[3]));
}


// What reduce_kernel<OP, true> launched with `vgrid` CTAs of kRedThreads threads computes, evaluated by ONE CTA of kRedThreads
// threads: virtual CTA after virtual CTA the same per-thread partials, the same block tree, then the same fold of the per-CTA
// partials in index order.  `x` may have been written by other CTAs of the running launch (coherent loads).  red: >= 32 floats of
// shared memory, vpart: >= vgrid floats of shared memory.  The result is valid in thread 0.
template <int OP>
__device__ __forceinline__ float reduce_in_one_cta(const float *x, uint64_t n, uint32_t vgrid, float *red, float *vpart) {
    float v = red_init<OP>();
    for (uint32_t vb = 0; vb < vgrid; ++vb) {
        v = thread_partial<OP, true, true>(x, nullptr, n, (uint64_t)vb * kRedThreads + threadIdx.x, (uint64_t)vgrid * kRedThreads);
        v = block_reduce<OP>(v, red);
        if (vgrid == 1) return v;
        if (threadIdx.x == 0) vpart[vb] = v;
    }
    __syncthreads();
    float acc = red_init<OP>();
    for (uint32_t i = threadIdx.x; i < vgrid; i += kRedThreads) acc = red_comb<OP>(acc, vpart[i]);
    return block_reduce<OP>(acc, red);
}

}  // namespace wgb
