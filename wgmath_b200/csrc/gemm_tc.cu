#include "common.cuh"
namespace wgb {
void tmap_cache_destroy(wgb_ctx *) {}
bool gemm_tc_eligible(const GemmProblem &) { return false; }
wgb_status launch_gemm_tc(wgb_pass *, const GemmProblem &, wgb_f32_mode, int *) { WGB_FAIL(WGB_ERR_UNSUPPORTED, "tc stub"); }
}
