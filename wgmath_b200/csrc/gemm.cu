// gemm.cu — path selection for wgb_gemm / wgb_gemm_ex (include/wgb200.h).
//
// The reference picks one of four WGSL pipelines by GemmVariant (gemm.rs:102-115); all four compute the
// same product, so here the variant only selects tr / non-tr and the *kernel family* is chosen from the
// operand types, the requested f32 mode and whether the views satisfy TMA's alignment rules:
//   bf16 operands, TMA-eligible          -> tcgen05 kind::f16 (bf16)            path 2
//   f32 operands, WGB_F32_TF32           -> tcgen05 kind::tf32, single pass      path 3
//   f32 operands, WGB_F32_AUTO / 3XTF32  -> tcgen05 kind::tf32, 3-pass hi/lo     path 4   (parity-gated default)
//   anything else (odd strides / offsets, tiny problems, WGB_F32_SIMT, K == 0) -> FFMA kernel, path 1
#include "common.cuh"

namespace wgb {

bool gemm_tc_direct_f32_ok(const GemmProblem &g);   // gemm_tc.cu

bool gemm_fused_eligible(const GemmProblem &g, wgb_f32_mode mode) {
    const uint64_t work = (uint64_t)g.M * g.N * (uint64_t)g.K;
    if (work < (uint64_t)96 * 96 * 96 || g.K == 0 || mode == WGB_F32_SIMT || !gemm_tc_eligible(g)) return false;
    if (g.in_dtype == WGB_F32 && (g.nmats > 65535 || (mode == WGB_F32_TF32 && !gemm_tc_direct_f32_ok(g)))) return false;
    return true;
}

wgb_status gemm_dispatch(wgb_pass *p, const GemmProblem &g, wgb_f32_mode mode) {
    const uint64_t work = (uint64_t)g.M * g.N * (uint64_t)g.K;
    // Below ~64^3 per matrix a tensor-core launch (tensor-map encode + 2 helper kernels for 3xTF32) costs more
    // than the FFMA kernel's whole run time.
    const bool tiny = work < (uint64_t)96 * 96 * 96 || g.K == 0;
    // f32 matrices of at most 128 x 128 (batches of small matrices): one 3xTF32 tile per matrix plus the two split kernels
    // measured slower than the FFMA tiles (128 x 128 x 128 x 1024: 215 vs 165 us; profiles/r1_batched_probe.txt)
    const bool small_f32 = g.in_dtype == WGB_F32 && mode == WGB_F32_AUTO && g.M <= 128 && g.N <= 128 && g.K <= 1024 && !g.fused;
    const bool want_tc = !tiny && !small_f32 && mode != WGB_F32_SIMT && gemm_tc_eligible(g);
    if (want_tc) {
        int path = 0;
        wgb_status s = launch_gemm_tc(p, g, mode, &path);
        if (s == WGB_OK) p->last_gemm_path = path;
        return s;
    }
    if (g.fused && g.fused->nranks > 1)
        WGB_FAIL(WGB_ERR_UNSUPPORTED, "fused all-gather needs the tensor-core GEMM path (TMA-eligible views, problem >= 96^3)");
    return launch_gemm_simt(p, g);
}


// ------------------------------------------------------------------------------------------------------------
// GEMM with the caller's reduction fused in (wgb_gemm_reduce): result[j] = reduce_i (m1 * m2)[i, j]
// (axis 1, "column reduction": what one Reduce dispatch per GpuMatrix::column(j) of the product gives, reduce.rs:100-113 +
// tensor.rs:574-585) or result[i] = reduce_j (m1 * m2)[i, j] (axis 2).  The product is never stored: the tensor-core epilogue
// leaves per-32-row (per-32-column) partial results and a fold kernel combines them in index order — deterministic, equal to
// the two-dispatch chain up to f32 rounding (the chain's own tree spans the whole column, so the bits can differ).
// ------------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ float red_init_rt(int op) {
    return op == WGB_RED_MIN ? 3.4e38f : op == WGB_RED_MAX ? -3.4e38f : op == WGB_RED_PROD ? 1.0f : 0.0f;
}
__device__ __forceinline__ float red_comb_rt(int op, float a, float b) {
    return op == WGB_RED_MIN ? fminf(a, b) : op == WGB_RED_MAX ? fmaxf(a, b) : op == WGB_RED_PROD ? a * b : a + b;
}
// result[i] = fold over p of partials[p * n_out + i], p ascending
__global__ void __launch_bounds__(256) gemm_reduce_fold_kernel(const float *__restrict__ partials, uint32_t count, uint32_t n_out, int op,
                                                               float *__restrict__ result) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    float acc = red_init_rt(op);
    for (uint32_t p = 0; p < count; ++p) acc = red_comb_rt(op, acc, __ldcg(partials + (uint64_t)p * n_out + i));
    result[i] = acc;
}
// fallback for products the tensor-core path does not take (tiny / unaligned views): reduce a stored column-major product
__global__ void __launch_bounds__(256) reduce_axis_kernel(const float *__restrict__ c, uint32_t M, uint32_t N, int axis, int op,
                                                          float *__restrict__ result) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_out = axis == 1 ? N : M, n_red = axis == 1 ? M : N;
    if (i >= n_out) return;
    float acc = red_init_rt(op);
    for (uint32_t r = 0; r < n_red; ++r) {
        const float x = axis == 1 ? c[(uint64_t)i * M + r] : c[(uint64_t)r * M + i];
        acc = red_comb_rt(op, acc, op == WGB_RED_SQNORM ? x * x : x);
    }
    result[i] = acc;
}
}  // namespace

wgb_status gemm_reduce_dispatch(wgb_pass *p, GemmProblem g, wgb_f32_mode mode, int axis, int op, float *result) {
    wgb_ctx *ctx = p->ctx;
    const uint32_t n_out = axis == 1 ? g.N : g.M, n_red = axis == 1 ? g.M : g.N;
    const uint64_t work = (uint64_t)g.M * g.N * (uint64_t)g.K;
    const bool tc = work >= (uint64_t)96 * 96 * 96 && g.K > 0 && mode != WGB_F32_SIMT && gemm_tc_eligible(g) &&
                    !(g.in_dtype == WGB_F32 && mode == WGB_F32_TF32 && !gemm_tc_direct_f32_ok(g)) && g.nmats == 1;
    if (tc) {
        const uint32_t count = (n_red + 31) / 32;
        void *w = nullptr;
        WGB_TRY(workspace_reserve(ctx, 7, (size_t)count * n_out * sizeof(float), &w));
        g.red_axis = axis;
        g.red_op = op;
        g.red_partials = (float *)w;
        g.out_dtype = WGB_F32;
        g.c = w;   // never written (the reduction replaces the store); any valid pointer keeps the launcher's checks happy
        g.c_off = 0; g.ldc = g.M; g.sc = (uint64_t)g.M * g.N;
        int path = 0;
        WGB_TRY(launch_gemm_tc(p, g, mode, &path));
        p->last_gemm_path = path;
        gemm_reduce_fold_kernel<<<(n_out + 255) / 256, 256, 0, p->stream>>>((const float *)w, count, n_out, op, result);
        WGB_CUDA(cudaGetLastError());
        count_launch(ctx);
        return WGB_OK;
    }
    // the FFMA kernel has no fused reduction: store the (small) product in a scratch matrix, then reduce it
    void *w = nullptr;
    WGB_TRY(workspace_reserve(ctx, 7, (size_t)g.M * g.N * sizeof(float) + 16, &w));
    g.out_dtype = WGB_F32;
    g.c = w;
    g.c_off = 0; g.ldc = g.M; g.sc = (uint64_t)g.M * g.N;
    WGB_TRY(gemm_dispatch(p, g, mode));
    reduce_axis_kernel<<<(n_out + 255) / 256, 256, 0, p->stream>>>((const float *)w, g.M, g.N, axis, op, result);
    WGB_CUDA(cudaGetLastError());
    count_launch(ctx);
    return WGB_OK;
}

}  // namespace wgb
