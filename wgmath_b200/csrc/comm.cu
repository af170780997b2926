#include "common.cuh"
namespace wgb { void comm_destroy(wgb_ctx *) {} }
extern "C" {
wgb_status wgb_comm_get_unique_id(void *) { WGB_FAIL(WGB_ERR_UNSUPPORTED, "comm stub"); }
wgb_status wgb_comm_init_rank(wgb_ctx *, int, int, const void *) { WGB_FAIL(WGB_ERR_UNSUPPORTED, "comm stub"); }
wgb_status wgb_comm_destroy(wgb_ctx *) { return WGB_OK; }
wgb_status wgb_gemm_row_sharded(wgb_pass *, wgb_gemm_variant, wgb_buffer *, const wgb_buffer *, const wgb_view_shape *, const wgb_buffer *, const wgb_view_shape *, wgb_dtype, wgb_dtype, wgb_f32_mode, int) { WGB_FAIL(WGB_ERR_UNSUPPORTED, "comm stub"); }
}
