// gemm_tc_inst_bf16_bf16out_3.cu — explicit instantiation of the tcgen05 GEMM launcher for one operand family
// (KIND = 0 [0 bf16, 1 tf32], A MN-major = true, B MN-major = false, passes = 1, output = __nv_bfloat16); see gemm_tc_kernel.cuh.
#include "gemm_tc_kernel.cuh"

namespace wgb {
namespace tc {
template wgb_status launch_sel<0, true, false, 1, __nv_bfloat16>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
}  // namespace tc
}  // namespace wgb
