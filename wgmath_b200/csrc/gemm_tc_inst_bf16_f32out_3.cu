// gemm_tc_inst_bf16_f32out_3.cu — explicit instantiation of the tcgen05 GEMM launcher for one operand family
// (KIND = 0 [0 bf16, 1 tf32], A MN-major = true, B MN-major = false, passes = 1, output = float); see gemm_tc_kernel.cuh.
#include "gemm_tc_kernel.cuh"

namespace wgb {
namespace tc {
template wgb_status launch_sel<0, true, false, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
}  // namespace tc
}  // namespace wgb
