// common.cuh — internal definitions shared by the translation units of libwgebra_b200.so.
// Handles, error plumbing, view-shape validation, launch accounting and the split-reduction
// scratch that the bandwidth-bound kernels use.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/wgb200.h"

namespace wgb {

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define WGB_FAIL(status, ...)          \
    do {                               \
        ::wgb::set_error(__VA_ARGS__); \
        return (status);               \
    } while (0)

#define WGB_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::wgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                      \
            return _e == cudaErrorMemoryAllocation ? WGB_ERR_OOM : WGB_ERR_CUDA;             \
        }                                                                                    \
    } while (0)

#define WGB_TRY(expr)                      \
    do {                                   \
        wgb_status _s = (expr);            \
        if (_s != WGB_OK) return _s;       \
    } while (0)

// ---------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------
struct Scratch {
    float *partials = nullptr;      // split-reduction partial results
    size_t partials_floats = 0;
    unsigned int *counters = nullptr;  // last-arriver tickets; always left at zero
    size_t n_counters = 0;
};

struct Workspace {  // growable device workspace (3xTF32 operand splits, sharded GEMM panels)
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct CommState;  // comm.cu

struct HostGemmState {   // wgb_gemm_host / wgb_gemm_host_enqueue: two alternating device operand slots
    uint64_t calls = 0;
    bool pending[2] = {false, false};
    cudaEvent_t done[2] = {nullptr, nullptr};   // recorded on the download stream after a slot's last panel has left
    std::vector<cudaEvent_t> evs;               // per-panel upload / product events
};

}  // namespace wgb

struct wgb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // the in-order queue
    cudaStream_t comm_stream = nullptr;  // side stream for the all-gather of C / the downloads of wgb_gemm_host
    cudaStream_t h2d_stream = nullptr;   // upload stream of wgb_gemm_host (created on first use)
    cudaDeviceProp prop{};
    std::atomic<uint64_t> launches{0};
    std::atomic<int> refs{1};            // the caller's handle + one per live buffer / event / pass / graph / peer group (abi.cu)
    wgb::Scratch scratch;
    wgb::Workspace ws[8];   // 0/1: 3xTF32 operand splits, 2: split-K partials, 3: host GEMM slots, 4: scan, 5: sort, 6: diagnostics,
                            // 7: the product vector of a fused Gemv -> Reduce
    uint64_t ws_generation = 0;   // bumped whenever a workspace / scratch block is reallocated (captured graphs hold the old pointers)
    wgb::CommState *comm = nullptr;
    wgb::HostGemmState host_gemm;
    std::mutex mu;
    void *tmap_cache = nullptr;  // gemm_tc.cu
    unsigned long long *tc_trace = nullptr;  // diagnostics: per-cluster timeline of the last tcgen05 GEMM (wgb_debug_tc_trace)
};

struct wgb_pass {
    wgb_ctx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    wgb_event *end_ts = nullptr;
    int last_gemm_path = 0;
    int last_tc[WGB_TC_CONFIG_WORDS] = {};   // configuration of the last tcgen05 GEMM launch (wgb_pass_last_gemm_config)
    uint64_t nvtx_range = 0;   // nvtxRangeId_t of the pass label
};

struct wgb_buffer {
    wgb_ctx *ctx = nullptr;
    void *ptr = nullptr;  // device pointer, or pinned host pointer when host_pinned
    size_t bytes = 0;
    bool host_pinned = false;
    bool owned = true;
    uint32_t usage = 0;
};

struct wgb_event {
    wgb_ctx *ctx = nullptr;
    cudaEvent_t ev = nullptr;
};

namespace wgb {

// Context lifetime: wgpu handles are reference counted (Arc<Device> inside every Buffer), so dropping the device before the
// buffers is legal in the reference.  Same here: wgb_ctx_destroy drops the caller's reference, every child object holds one, and
// the context is torn down when the last of them goes (host bindings whose finalisers run in arbitrary order — Python's cycle
// collector at interpreter exit — rely on this).
inline void ctx_retain(wgb_ctx *ctx) { ctx->refs.fetch_add(1, std::memory_order_relaxed); }
void ctx_release(wgb_ctx *ctx);

inline void count_launch(wgb_ctx *ctx, uint64_t n = 1) { ctx->launches.fetch_add(n, std::memory_order_relaxed); }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline size_t dtype_size(wgb_dtype d) { return d == WGB_BF16 ? 2 : 4; }

// Largest element index (exclusive) touched by a matrix / cube view: offset + extent.
inline uint64_t view_extent(const wgb_view_shape &s) {
    if (s.size[0] == 0 || s.size[1] == 0 || s.size[2] == 0) return 0;
    return (uint64_t)s.offset + (uint64_t)(s.size[2] - 1) * s.stride_mat + (uint64_t)(s.size[1] - 1) * s.stride +
           s.size[0];
}
inline uint64_t vector_extent(const wgb_view_shape &s) { return s.size[0] ? (uint64_t)s.offset + s.size[0] : 0; }

wgb_status check_view(const wgb_buffer *b, const wgb_view_shape &s, size_t elem_size, const char *what,
                      bool vector_only = false);

// Grow-only device workspace slot; contents are undefined after a grow.
wgb_status workspace_reserve(wgb_ctx *ctx, int slot, size_t bytes, void **out);
wgb_status scratch_reserve(wgb_ctx *ctx, size_t partial_floats, size_t counters);
// One of the two alternating device operand slots of the host-operand GEMMs (abi.cu): slots sit at stable offsets (0 and half of
// workspace 3), so products of different sizes never overlap a slot that is still in flight; growing drains all three streams.
// The upload stream is made to wait for the slot's previous user.
wgb_status host_gemm_slot(wgb_ctx *ctx, size_t slot_bytes, char **base, int *slot);
// SMs the persistent GEMM leaves free for the NCCL CTAs of an in-flight exchange (0 without a communicator).
int comm_sm_margin(const wgb_ctx *ctx);

// ---- kernels launched from the dispatch layer (one per .cu) --------------------------------
wgb_status launch_op_assign(wgb_pass *p, int op, float *a, const float *b, uint64_t n);
wgb_status launch_reduce(wgb_pass *p, int op, const float *x, const float *y, uint64_t n, float *result);
wgb_status launch_reduce_columns(wgb_pass *p, int op, const float *m, const wgb_view_shape &ms, float *out);
wgb_status launch_fill_uniform(wgb_pass *p, void *base, const wgb_view_shape &s, wgb_dtype dt, uint64_t seed,
                               uint32_t row0, uint32_t col0);
wgb_status launch_prefix_sum(wgb_pass *p, uint32_t *data, uint64_t n);
wgb_status launch_radix_sort(wgb_pass *p, const uint32_t *keys_in, const uint32_t *vals_in, uint32_t len, const uint32_t *n_dev,
                             uint32_t sorting_bits, uint32_t *keys_out, uint32_t *vals_out);
wgb_status launch_gemv(wgb_pass *p, bool tr, float *out, const wgb_view_shape &so, const float *m,
                       const wgb_view_shape &sm, const float *v, const wgb_view_shape &sv, int op = -1,
                       const float *operand = nullptr, const wgb_view_shape *operand_shape = nullptr, int red_op = -1,
                       float *red_result = nullptr);
int reduce_grid_for(wgb_ctx *ctx, int op, uint64_t n);   // the grid wgb_reduce launches for n elements (level1.cu)
constexpr uint32_t kGemvReduceMaxGrid = 2048;            // largest such grid the fused Gemv -> Reduce tail emulates (gemv.cu)

// Fused all-gather of the GEMM output over peer (NVLink-mapped) memory: the epilogue stores every output element
// into the gathered buffer of every rank, then the last CTA publishes a completion flag to each peer.
constexpr int kMaxPeers = 8;
struct FusedGather {
    int nranks = 1, rank = 0;
    void *peer_c[kMaxPeers] = {};            // gathered buffer of rank q (peer mapping; [rank] is the local one)
    void *mc_c = nullptr;                    // the same buffer through the group's NVSwitch multicast mapping (null: none)
    unsigned int *ready_local = nullptr;     // ready_local[q]: rank q may be written into for epoch >= value
    unsigned int *done_remote[kMaxPeers] = {};  // done array of rank q (we write entry [rank])
    unsigned int *cta_counter = nullptr;     // local: CTAs finished (left at zero)
    unsigned int epoch = 0;
    long long timeout = 0;                   // flag-wait bound in SM cycles (0: unbounded)
};

struct GemmProblem {
    const FusedGather *fused = nullptr;   // non-null: write C into every rank's gathered buffer (comm.cu)
    bool tr;               // K is the contiguous axis of m1 (out = tr(m1) * m2 on column-major views)
    bool b_nmajor = false; // N is the contiguous axis of m2 (a row-major m2); false: K contiguous (column-major m2)
    uint32_t M, N, K, nmats;
    const void *a;         // m1 base (element 0 of the buffer)
    const void *b;         // m2 base
    void *c;               // out base
    uint64_t a_off, b_off, c_off;     // element offsets
    uint64_t lda, ldb, ldc;           // column strides, elements
    uint64_t sa, sb, sc;              // matrix strides, elements
    wgb_dtype in_dtype, out_dtype;
    // fused element-wise epilogue: out = (m1 * m2) (ep_op) e, e a view of out's element type (ep_op < 0: none)
    int ep_op = -1;
    const void *e = nullptr;
    uint64_t e_off = 0, lde = 0, se = 0;
    // fused reduction of the product (wgb_gemm_reduce): c is not written; red_axis 1 = one value per column, 2 = one per row
    int red_axis = 0, red_op = 0;
    float *red_partials = nullptr;   // [ceil(M / 32)][N] (axis 1) or [ceil(N / 32)][M] (axis 2), filled by the epilogue
};
// out-of-line helper shared by both GEMM kernels
template <typename T>
__device__ __forceinline__ float epilogue_apply(int op, float v, const T *e);
template <>
__device__ __forceinline__ float epilogue_apply<float>(int op, float v, const float *e) {
    const float x = *e;
    return op == WGB_OP_ADD ? v + x : op == WGB_OP_SUB ? v - x : op == WGB_OP_MUL ? v * x : v / x;
}
template <>
__device__ __forceinline__ float epilogue_apply<__nv_bfloat16>(int op, float v, const __nv_bfloat16 *e) {
    const float x = __bfloat162float(*e);
    return op == WGB_OP_ADD ? v + x : op == WGB_OP_SUB ? v - x : op == WGB_OP_MUL ? v * x : v / x;
}
wgb_status launch_gemm_simt(wgb_pass *p, const GemmProblem &g);
// tcgen05 path. `passes`: 1 (bf16 or single-pass tf32) or 3 (3xTF32; operands pre-split).
bool gemm_tc_eligible(const GemmProblem &g);
wgb_status launch_gemm_tc(wgb_pass *p, const GemmProblem &g, wgb_f32_mode mode, int *path_out);
// true when gemm_dispatch will run the peer-storing tensor-core kernel for this problem (checked before a fused call signals its peers)
bool gemm_fused_eligible(const GemmProblem &g, wgb_f32_mode mode);

// column-panel range restriction used by the sharded GEMM (compute only n in [n_begin, n_end))
wgb_status gemm_dispatch(wgb_pass *p, const GemmProblem &g, wgb_f32_mode mode);
// the product reduced along one axis without being stored (gemm.cu): axis 1 = one value per column, 2 = one per row
wgb_status gemm_reduce_dispatch(wgb_pass *p, GemmProblem g, wgb_f32_mode mode, int axis, int op, float *result);
bool gemm_tc_direct_f32_ok(const GemmProblem &g);

}  // namespace wgb
