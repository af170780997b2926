// gemm_tc_inst_tf32_3.cu — explicit instantiation of the tcgen05 GEMM launcher for one operand family
// (KIND = 1 [0 bf16, 1 tf32], A MN-major = true, B MN-major = false, passes = 3, output = float); see gemm_tc_kernel.cuh.
#include "gemm_tc_kernel.cuh"

namespace wgb {
namespace tc {
template wgb_status launch_sel<1, true, false, 3, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
}  // namespace tc
}  // namespace wgb
