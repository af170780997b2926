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

}  // namespace wgb
