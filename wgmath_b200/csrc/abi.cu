// abi.cu — the extern "C" surface of libwgebra_b200.so (include/wgb200.h): contexts, passes,
// buffers, events, and the validation / dispatch layer in front of the kernels.
//
// Reference behaviour mirrored here (paths under /root/reference/crates/):
//   wgcore/src/gpu.rs:15-58            GpuInstance::new           -> wgb_ctx_create
//   wgcore/src/kernel.rs:7-27          compute_pass               -> wgb_pass_begin / _end
//   wgcore/src/kernel.rs:103-148       zero-sized binding / empty grid => dispatch skipped
//   wgcore/src/tensor.rs:112-186       TensorBuilder::build*      -> wgb_buffer_create*
//   wgcore/src/tensor.rs:227-265       copy_from / copy_from_view -> wgb_buffer_copy
//   wgcore/src/tensor.rs:300-384       read*                      -> wgb_buffer_read
//   wgebra/src/linalg/gemm.rs:78-96    dimension asserts          -> WGB_ERR_DIM_MISMATCH
//   wgebra/src/linalg/gemv.rs:77-124   dimension asserts, TrFast fallback
//   wgebra/src/linalg/op_assign.rs:82-86, reduce.rs:100-113
#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler injects the NVTX library

#include "common.cuh"

namespace wgb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

wgb_status check_view(const wgb_buffer *b, const wgb_view_shape &s, size_t elem_size, const char *what,
                      bool vector_only) {
    const uint64_t ext = vector_only ? vector_extent(s) : view_extent(s);
    if (ext * elem_size > b->bytes)
        WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "%s: view reaches element %llu but the buffer holds %llu", what,
                 (unsigned long long)ext, (unsigned long long)(b->bytes / elem_size));
    return WGB_OK;
}

wgb_status workspace_reserve(wgb_ctx *ctx, int slot, size_t bytes, void **out) {
    Workspace &w = ctx->ws[slot];
    if (w.bytes < bytes) {
        // Growing is rare (first call at a new size); synchronise so no in-flight kernel loses its operands.
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (w.ptr) WGB_CUDA(cudaFree(w.ptr));
        w.ptr = nullptr;
        w.bytes = 0;
        ++ctx->ws_generation;
        size_t want = bytes + (bytes >> 3);
        cudaError_t e = cudaMalloc(&w.ptr, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            want = bytes;
            WGB_CUDA(cudaMalloc(&w.ptr, want));
        }
        w.bytes = want;
    }
    *out = w.ptr;
    return WGB_OK;
}

wgb_status scratch_reserve(wgb_ctx *ctx, size_t partial_floats, size_t counters) {
    Scratch &s = ctx->scratch;
    if (s.partials_floats < partial_floats) {
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (s.partials) WGB_CUDA(cudaFree(s.partials));
        ++ctx->ws_generation;
        s.partials = nullptr;
        s.partials_floats = 0;
        WGB_CUDA(cudaMalloc(&s.partials, partial_floats * sizeof(float)));
        s.partials_floats = partial_floats;
    }
    if (s.n_counters < counters) {
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (s.counters) WGB_CUDA(cudaFree(s.counters));
        ++ctx->ws_generation;
        s.counters = nullptr;
        s.n_counters = 0;
        WGB_CUDA(cudaMalloc(&s.counters, counters * sizeof(unsigned int)));
        WGB_CUDA(cudaMemset(s.counters, 0, counters * sizeof(unsigned int)));
        s.n_counters = counters;
    }
    return WGB_OK;
}

wgb_status host_gemm_slot(wgb_ctx *ctx, size_t slot_bytes, char **base, int *slot_out) {
    HostGemmState &hs = ctx->host_gemm;
    slot_bytes = (slot_bytes + 255) & ~(size_t)255;
    if (ctx->ws[3].bytes / 2 < slot_bytes + 256) {   // growing frees the slots: drain everything that may still use them
        WGB_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));
        WGB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
        hs.pending[0] = hs.pending[1] = false;
        void *w = nullptr;
        WGB_TRY(workspace_reserve(ctx, 3, 2 * (slot_bytes + 256), &w));
    }
    const size_t half = (ctx->ws[3].bytes / 2) & ~(size_t)255;
    const int slot = (int)(hs.calls++ & 1u);
    // the product that used this slot two calls ago must have left it (its download is the last user)
    if (hs.pending[slot]) WGB_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, hs.done[slot], 0));
    *base = (char *)ctx->ws[3].ptr + (size_t)slot * half;
    *slot_out = slot;
    return WGB_OK;
}

void tmap_cache_destroy(wgb_ctx *ctx);  // gemm_tc.cu
void comm_destroy(wgb_ctx *ctx);        // comm.cu

}  // namespace wgb

using namespace wgb;

extern "C" {

int wgb_abi_version(void) { return WGB200_ABI_VERSION; }
const char *wgb_last_error_string(void) { return wgb::g_err; }

// --------------------------------------------------------------------------- context
wgb_status wgb_ctx_create(int device_ordinal, wgb_ctx **out) {
    if (!out) WGB_FAIL(WGB_ERR_INVALID, "wgb_ctx_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        WGB_FAIL(WGB_ERR_NO_DEVICE, "no CUDA device visible (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device_ordinal < 0 || device_ordinal >= ndev)
        WGB_FAIL(WGB_ERR_INVALID, "device ordinal %d out of range [0, %d)", device_ordinal, ndev);
    wgb_ctx *ctx = new wgb_ctx();
    ctx->device = device_ordinal;
    if (cudaGetDeviceProperties(&ctx->prop, device_ordinal) != cudaSuccess) {
        delete ctx;
        WGB_FAIL(WGB_ERR_CUDA, "cudaGetDeviceProperties failed");
    }
    if (ctx->prop.major != 10) {
        int maj = ctx->prop.major, min = ctx->prop.minor;
        delete ctx;
        WGB_FAIL(WGB_ERR_NO_DEVICE, "device %d is sm_%d%d; libwgebra_b200 is built for sm_100a only", device_ordinal,
                 maj, min);
    }
    DeviceGuard g(device_ordinal);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        WGB_FAIL(WGB_ERR_CUDA, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    wgb_status s = scratch_reserve(ctx, (size_t)1 << 20, (size_t)1 << 16);
    if (s != WGB_OK) {
        delete ctx;
        return s;
    }
    *out = ctx;
    return WGB_OK;
}

}  // extern "C"

static void ctx_teardown(wgb_ctx *ctx);
namespace wgb {
void ctx_release(wgb_ctx *ctx) {
    if (ctx->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) ctx_teardown(ctx);
}
}  // namespace wgb

static void ctx_teardown(wgb_ctx *ctx) {
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->comm_stream);
    comm_destroy(ctx);
    tmap_cache_destroy(ctx);
    if (ctx->tc_trace) cudaFree(ctx->tc_trace);
    if (ctx->scratch.partials) cudaFree(ctx->scratch.partials);
    if (ctx->scratch.counters) cudaFree(ctx->scratch.counters);
    for (auto &w : ctx->ws)
        if (w.ptr) cudaFree(w.ptr);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->comm_stream);
    if (ctx->h2d_stream) {
        cudaStreamSynchronize(ctx->h2d_stream);
        cudaStreamDestroy(ctx->h2d_stream);
    }
    for (auto e : ctx->host_gemm.done)
        if (e) cudaEventDestroy(e);
    for (auto e : ctx->host_gemm.evs) cudaEventDestroy(e);
    delete ctx;
}

extern "C" {

wgb_status wgb_ctx_destroy(wgb_ctx *ctx) {
    if (!ctx) return WGB_OK;
    {   // queued work is drained now; the resources go with the last child object (common.cuh: context lifetime)
        DeviceGuard g(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->comm_stream);
    }
    wgb::ctx_release(ctx);
    return WGB_OK;
}

wgb_status wgb_debug_tc_trace(wgb_ctx *ctx, int enable, unsigned long long *out, size_t max_records, size_t *n_records) {
    if (!ctx) WGB_FAIL(WGB_ERR_INVALID, "wgb_debug_tc_trace: null context");
    DeviceGuard g(ctx->device);
    constexpr size_t kClusters = 256, kWords = 8;
    if (enable && !ctx->tc_trace) {
        WGB_CUDA(cudaMalloc(&ctx->tc_trace, kClusters * kWords * sizeof(unsigned long long)));
        WGB_CUDA(cudaMemset(ctx->tc_trace, 0, kClusters * kWords * sizeof(unsigned long long)));
    }
    if (out && ctx->tc_trace) {
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));
        const size_t n = max_records < kClusters ? max_records : kClusters;
        WGB_CUDA(cudaMemcpy(out, ctx->tc_trace, n * kWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (n_records) *n_records = n;
    } else if (n_records) {
        *n_records = 0;
    }
    if (!enable && ctx->tc_trace) {
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));
        WGB_CUDA(cudaFree(ctx->tc_trace));
        ctx->tc_trace = nullptr;
    }
    return WGB_OK;
}

wgb_status wgb_ctx_sync(wgb_ctx *ctx) {
    if (!ctx) WGB_FAIL(WGB_ERR_INVALID, "wgb_ctx_sync: null context");
    DeviceGuard g(ctx->device);
    WGB_CUDA(cudaStreamSynchronize(ctx->stream));
    WGB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    if (ctx->h2d_stream) WGB_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
    ctx->host_gemm.pending[0] = ctx->host_gemm.pending[1] = false;
    return WGB_OK;
}

wgb_status wgb_ctx_device_info(wgb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem,
                               char *name, size_t name_len) {
    if (!ctx) WGB_FAIL(WGB_ERR_INVALID, "null context");
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (total_mem) *total_mem = ctx->prop.totalGlobalMem;
    if (name && name_len) {
        strncpy(name, ctx->prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return WGB_OK;
}

wgb_status wgb_ctx_launch_count(wgb_ctx *ctx, uint64_t *count) {
    if (!ctx || !count) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    *count = ctx->launches.load();
    return WGB_OK;
}

wgb_status wgb_ctx_stream(wgb_ctx *ctx, void **cuda_stream) {
    if (!ctx || !cuda_stream) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    *cuda_stream = (void *)ctx->stream;
    return WGB_OK;
}

// --------------------------------------------------------------------------- passes
wgb_status wgb_pass_begin(wgb_ctx *ctx, const char *label, wgb_event *begin_ts, wgb_event *end_ts, wgb_pass **out) {
    if (!ctx || !out) WGB_FAIL(WGB_ERR_INVALID, "wgb_pass_begin: null argument");
    wgb_pass *p = new wgb_pass();
    p->ctx = ctx;
    p->stream = ctx->stream;
    p->end_ts = end_ts;
    // the reference labels every compute pass (kernel.rs:15-26, ComputePassDescriptor::label); here the label names an NVTX range
    // around the pass's recording, which is what ncu / nsys show next to the kernels
    p->nvtx_range = nvtxRangeStartA(label ? label : "compute_pass");
    ctx_retain(ctx);
    if (begin_ts) {
        DeviceGuard g(ctx->device);
        cudaError_t e = cudaEventRecord(begin_ts->ev, p->stream);
        if (e != cudaSuccess) {
            nvtxRangeEnd(p->nvtx_range);
            delete p;
            ctx_release(ctx);
            WGB_FAIL(WGB_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
        }
    }
    *out = p;
    return WGB_OK;
}

wgb_status wgb_pass_end(wgb_pass *pass) {
    if (!pass) return WGB_OK;
    wgb_status st = WGB_OK;
    if (pass->end_ts) {
        DeviceGuard g(pass->ctx->device);
        if (cudaEventRecord(pass->end_ts->ev, pass->stream) != cudaSuccess) {
            set_error("cudaEventRecord failed at pass end");
            st = WGB_ERR_CUDA;
        }
    }
    nvtxRangeEnd(pass->nvtx_range);
    wgb_ctx *ctx = pass->ctx;
    delete pass;
    ctx_release(ctx);
    return st;
}

wgb_status wgb_submit(wgb_ctx *ctx) {
    if (!ctx) WGB_FAIL(WGB_ERR_INVALID, "null context");
    // Kernels were enqueued at dispatch time; cudaStreamQuery nudges the driver to flush (WDDM-style batching
    // does not exist on Linux, so this is effectively free).
    DeviceGuard g(ctx->device);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return WGB_OK;  // recording
    cudaError_t e = cudaStreamQuery(ctx->stream);
    if (e != cudaSuccess && e != cudaErrorNotReady) WGB_FAIL(WGB_ERR_CUDA, "queue error: %s", cudaGetErrorString(e));
    return WGB_OK;
}

wgb_status wgb_pass_last_gemm_path(const wgb_pass *pass, int *path) {
    if (!pass || !path) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    *path = pass->last_gemm_path;
    return WGB_OK;
}

wgb_status wgb_pass_last_gemm_config(const wgb_pass *pass, int *config) {
    if (!pass || !config) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    for (int i = 0; i < WGB_TC_CONFIG_WORDS; ++i) config[i] = pass->last_gemm_path >= 2 ? pass->last_tc[i] : 0;
    return WGB_OK;
}

// --------------------------------------------------------------------------- graphs
}  // extern "C"
struct wgb_graph {
    wgb_ctx *ctx = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches_per_replay = 0;
    uint64_t launches_at_begin = 0;
    uint64_t ws_generation = 0;   // the context's workspace generation the recorded kernels were given pointers from
};
static thread_local uint64_t g_capture_launch_mark = 0;
extern "C" {

wgb_status wgb_graph_capture_begin(wgb_ctx *ctx) {
    if (!ctx) WGB_FAIL(WGB_ERR_INVALID, "null context");
    DeviceGuard g(ctx->device);
    WGB_CUDA(cudaStreamSynchronize(ctx->stream));
    g_capture_launch_mark = ctx->launches.load();
    WGB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    return WGB_OK;
}

wgb_status wgb_graph_capture_end(wgb_ctx *ctx, wgb_graph **out) {
    if (!ctx || !out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(ctx->device);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e != cudaSuccess || !graph) {
        (void)cudaGetLastError();
        WGB_FAIL(WGB_ERR_CUDA, "graph capture failed (%s): a dispatch synchronised or allocated during capture — warm the "
                               "workspaces by running the sequence once before capturing", cudaGetErrorString(e));
    }
    wgb_graph *gr = new wgb_graph();
    gr->ctx = ctx;
    gr->graph = graph;
    gr->launches_per_replay = ctx->launches.load() - g_capture_launch_mark;
    gr->ws_generation = ctx->ws_generation;
    ctx->launches.store(g_capture_launch_mark);   // recorded, not executed
    e = cudaGraphInstantiate(&gr->exec, graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        delete gr;
        WGB_FAIL(WGB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    }
    ctx_retain(ctx);
    *out = gr;
    return WGB_OK;
}

wgb_status wgb_graph_launch(wgb_graph *graph) {
    if (!graph) WGB_FAIL(WGB_ERR_INVALID, "null graph");
    DeviceGuard g(graph->ctx->device);
    // The recorded kernels carry raw pointers into the context's workspaces (3xTF32 operand copies, split partials, scan / sort
    // scratch).  A later eager call with a larger problem reallocates them: replaying would touch freed memory.
    if (graph->ws_generation != graph->ctx->ws_generation)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_graph_launch: a workspace of the context was reallocated after this graph was recorded "
                                  "(a larger problem ran since); record the graph again");
    WGB_CUDA(cudaGraphLaunch(graph->exec, graph->ctx->stream));
    count_launch(graph->ctx, graph->launches_per_replay);
    return WGB_OK;
}

wgb_status wgb_graph_destroy(wgb_graph *graph) {
    if (!graph) return WGB_OK;
    DeviceGuard g(graph->ctx->device);
    cudaStreamSynchronize(graph->ctx->stream);
    if (graph->exec) cudaGraphExecDestroy(graph->exec);
    if (graph->graph) cudaGraphDestroy(graph->graph);
    wgb_ctx *ctx = graph->ctx;
    delete graph;
    ctx_release(ctx);
    return WGB_OK;
}

// --------------------------------------------------------------------------- buffers
wgb_status wgb_buffer_create(wgb_ctx *ctx, size_t bytes, uint32_t usage, wgb_buffer **out) {
    if (!ctx || !out) WGB_FAIL(WGB_ERR_INVALID, "wgb_buffer_create: null argument");
    DeviceGuard g(ctx->device);
    wgb_buffer *b = new wgb_buffer();
    b->ctx = ctx;
    b->bytes = bytes;
    b->usage = usage;
    b->host_pinned = (usage & (WGB_USAGE_MAP_READ | WGB_USAGE_MAP_WRITE)) != 0;
    if (bytes) {
        cudaError_t e = b->host_pinned ? cudaHostAlloc(&b->ptr, bytes, cudaHostAllocDefault) : cudaMalloc(&b->ptr, bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            delete b;
            WGB_FAIL(WGB_ERR_OOM, "allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    ctx_retain(ctx);
    *out = b;
    return WGB_OK;
}

wgb_status wgb_buffer_create_init(wgb_ctx *ctx, const void *host_data, size_t bytes, uint32_t usage,
                                  wgb_buffer **out) {
    if (bytes && !host_data) WGB_FAIL(WGB_ERR_INVALID, "wgb_buffer_create_init: null data");
    WGB_TRY(wgb_buffer_create(ctx, bytes, usage, out));
    if (bytes) {
        DeviceGuard g(ctx->device);
        wgb_buffer *b = *out;
        if (b->host_pinned) {
            memcpy(b->ptr, host_data, bytes);
        } else {
            // create_buffer_init semantics: the data is captured at creation, so the copy is complete on return.
            cudaError_t e = cudaMemcpyAsync(b->ptr, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                wgb_buffer_destroy(b);
                *out = nullptr;
                WGB_FAIL(WGB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
            }
        }
    }
    return WGB_OK;
}

wgb_status wgb_buffer_wrap(wgb_ctx *ctx, void *device_ptr, size_t bytes, wgb_buffer **out) {
    if (!ctx || !out || (!device_ptr && bytes)) WGB_FAIL(WGB_ERR_INVALID, "wgb_buffer_wrap: null argument");
    wgb_buffer *b = new wgb_buffer();
    b->ctx = ctx;
    b->ptr = device_ptr;
    b->bytes = bytes;
    b->owned = false;
    b->usage = WGB_USAGE_STORAGE | WGB_USAGE_COPY_SRC | WGB_USAGE_COPY_DST;
    ctx_retain(ctx);
    *out = b;
    return WGB_OK;
}

wgb_status wgb_buffer_destroy(wgb_buffer *buf) {
    if (!buf) return WGB_OK;
    if (buf->owned && buf->ptr) {
        DeviceGuard g(buf->ctx->device);
        // wgpu keeps a destroyed buffer alive until queued work that uses it has finished.
        cudaStreamSynchronize(buf->ctx->stream);
        if (buf->host_pinned) cudaFreeHost(buf->ptr);
        else cudaFree(buf->ptr);
    }
    wgb_ctx *ctx = buf->ctx;
    delete buf;
    ctx_release(ctx);
    return WGB_OK;
}

wgb_status wgb_buffer_size(const wgb_buffer *buf, size_t *bytes) {
    if (!buf || !bytes) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    *bytes = buf->bytes;
    return WGB_OK;
}

wgb_status wgb_buffer_device_ptr(const wgb_buffer *buf, void **ptr) {
    if (!buf || !ptr) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    *ptr = buf->ptr;
    return WGB_OK;
}

wgb_status wgb_buffer_write(wgb_ctx *ctx, wgb_buffer *dst, size_t dst_off, const void *host_src, size_t bytes) {
    if (!ctx || !dst || (!host_src && bytes)) WGB_FAIL(WGB_ERR_INVALID, "wgb_buffer_write: null argument");
    if (dst_off + bytes > dst->bytes) WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_buffer_write: range exceeds the buffer");
    if (!bytes) return WGB_OK;
    DeviceGuard g(ctx->device);
    WGB_CUDA(cudaMemcpyAsync((char *)dst->ptr + dst_off, host_src, bytes,
                             dst->host_pinned ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice, ctx->stream));
    return WGB_OK;
}

wgb_status wgb_buffer_copy(wgb_ctx *ctx, wgb_pass *pass, wgb_buffer *dst, size_t dst_off, const wgb_buffer *src,
                           size_t src_off, size_t bytes) {
    if (!ctx || !dst || !src) WGB_FAIL(WGB_ERR_INVALID, "wgb_buffer_copy: null argument");
    if (dst_off + bytes > dst->bytes || src_off + bytes > src->bytes)
        WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_buffer_copy: range exceeds a buffer");
    if (!bytes) return WGB_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = pass ? pass->stream : ctx->stream;
    WGB_CUDA(cudaMemcpyAsync((char *)dst->ptr + dst_off, (const char *)src->ptr + src_off, bytes, cudaMemcpyDefault, st));
    return WGB_OK;
}

wgb_status wgb_buffer_read(wgb_ctx *ctx, const wgb_buffer *src, size_t src_off, void *host_dst, size_t bytes) {
    if (!ctx || !src || (!host_dst && bytes)) WGB_FAIL(WGB_ERR_INVALID, "wgb_buffer_read: null argument");
    if (src_off + bytes > src->bytes) WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_buffer_read: range exceeds the buffer");
    DeviceGuard g(ctx->device);
    if (src->host_pinned) {
        WGB_CUDA(cudaStreamSynchronize(ctx->stream));  // device.poll(wait)
        if (bytes) memcpy(host_dst, (const char *)src->ptr + src_off, bytes);
        return WGB_OK;
    }
    if (bytes)
        WGB_CUDA(cudaMemcpyAsync(host_dst, (const char *)src->ptr + src_off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    WGB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WGB_OK;
}

wgb_status wgb_host_alloc(size_t bytes, void **out) {
    if (!out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!bytes) return WGB_OK;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        WGB_FAIL(WGB_ERR_OOM, "pinned allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return WGB_OK;
}

wgb_status wgb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
    return WGB_OK;
}

// --------------------------------------------------------------------------- events
wgb_status wgb_event_create(wgb_ctx *ctx, wgb_event **out) {
    if (!ctx || !out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(ctx->device);
    wgb_event *ev = new wgb_event();
    ev->ctx = ctx;
    cudaError_t e = cudaEventCreate(&ev->ev);
    if (e != cudaSuccess) {
        delete ev;
        WGB_FAIL(WGB_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e));
    }
    ctx_retain(ctx);
    *out = ev;
    return WGB_OK;
}

wgb_status wgb_event_destroy(wgb_event *ev) {
    if (!ev) return WGB_OK;
    {
        DeviceGuard g(ev->ctx->device);
        cudaEventDestroy(ev->ev);
    }
    wgb_ctx *ctx = ev->ctx;
    delete ev;
    ctx_release(ctx);
    return WGB_OK;
}

wgb_status wgb_event_record(wgb_event *ev, wgb_pass *pass) {
    if (!ev) WGB_FAIL(WGB_ERR_INVALID, "null event");
    DeviceGuard g(ev->ctx->device);
    WGB_CUDA(cudaEventRecord(ev->ev, pass ? pass->stream : ev->ctx->stream));
    return WGB_OK;
}

wgb_status wgb_event_elapsed_ms(wgb_event *begin, wgb_event *end, float *ms) {
    if (!begin || !end || !ms) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(end->ctx->device);
    WGB_CUDA(cudaEventSynchronize(end->ev));
    WGB_CUDA(cudaEventElapsedTime(ms, begin->ev, end->ev));
    return WGB_OK;
}

// --------------------------------------------------------------------------- operators
static bool any_zero_buffer(std::initializer_list<const wgb_buffer *> bufs) {
    for (auto b : bufs)
        if (b->bytes == 0) return true;  // kernel.rs:111-113: zero-sized binding => not queueable
    return false;
}

wgb_status wgb_op_assign(wgb_pass *pass, wgb_op_assign_variant op, wgb_buffer *a, const wgb_view_shape *sa,
                         const wgb_buffer *b, const wgb_view_shape *sb) {
    if (!pass || !a || !b || !sa || !sb) WGB_FAIL(WGB_ERR_INVALID, "wgb_op_assign: null argument");
    if ((int)op < 0 || (int)op > WGB_OP_COPY) WGB_FAIL(WGB_ERR_INVALID, "wgb_op_assign: unknown op %d", (int)op);
    if (sa->size[0] != sb->size[0])  // op_assign.rs:82-86
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Op-assign: dimension mismatch. (%u vs %u)", sa->size[0], sb->size[0]);
    if (any_zero_buffer({a, b}) || sa->size[0] == 0) return WGB_OK;
    WGB_TRY(check_view(a, *sa, 4, "op_assign a", true));
    WGB_TRY(check_view(b, *sb, 4, "op_assign b", true));
    DeviceGuard g(pass->ctx->device);
    return launch_op_assign(pass, (int)op, (float *)a->ptr + sa->offset, (const float *)b->ptr + sb->offset, sa->size[0]);
}

wgb_status wgb_prefix_sum(wgb_pass *pass, wgb_buffer *data, const wgb_view_shape *ds) {
    if (!pass || !data || !ds) WGB_FAIL(WGB_ERR_INVALID, "wgb_prefix_sum: null argument");
    if (any_zero_buffer({data}) || ds->size[0] == 0) return WGB_OK;
    WGB_TRY(check_view(data, *ds, 4, "prefix_sum data", true));
    DeviceGuard g(pass->ctx->device);
    return launch_prefix_sum(pass, (uint32_t *)data->ptr + ds->offset, ds->size[0]);
}

wgb_status wgb_radix_sort(wgb_pass *pass, const wgb_buffer *ik, const wgb_view_shape *iks, const wgb_buffer *iv,
                          const wgb_view_shape *ivs, const wgb_buffer *n_sort, uint32_t sorting_bits, wgb_buffer *ok,
                          const wgb_view_shape *oks, wgb_buffer *ov, const wgb_view_shape *ovs) {
    if (!pass || !ik || !iks || !iv || !ivs || !n_sort || !ok || !oks || !ov || !ovs)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_radix_sort: null argument");
    if (iks->size[0] != ivs->size[0])   // radix_sort/mod.rs:121-125
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Input keys and values must have the same number of elements (%u vs %u)", iks->size[0],
                 ivs->size[0]);
    if (sorting_bits > 32) WGB_FAIL(WGB_ERR_INVALID, "Can only sort up to 32 bits");   // mod.rs:126
    if (oks->size[0] < iks->size[0] || ovs->size[0] < iks->size[0])
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "radix sort: outputs (%u keys, %u values) shorter than the %u input pairs", oks->size[0],
                 ovs->size[0], iks->size[0]);
    if (any_zero_buffer({ik, iv, n_sort, ok, ov}) || iks->size[0] == 0) return WGB_OK;
    if (n_sort->bytes < 4) WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_radix_sort: n_sort buffer smaller than one u32");
    WGB_TRY(check_view(ik, *iks, 4, "radix_sort input_keys", true));
    WGB_TRY(check_view(iv, *ivs, 4, "radix_sort input_values", true));
    WGB_TRY(check_view(ok, *oks, 4, "radix_sort output_keys", true));
    WGB_TRY(check_view(ov, *ovs, 4, "radix_sort output_values", true));
    const uint32_t *kin = (const uint32_t *)ik->ptr + iks->offset, *vin = (const uint32_t *)iv->ptr + ivs->offset;
    uint32_t *kout = (uint32_t *)ok->ptr + oks->offset, *vout = (uint32_t *)ov->ptr + ovs->offset;
    const size_t len = iks->size[0];
    auto overlap = [len](const uint32_t *a, const uint32_t *b) { return a < b + len && b < a + len; };
    if (overlap(kin, kout) || overlap(kin, vout) || overlap(vin, kout) || overlap(vin, vout) || overlap(kout, vout))
        WGB_FAIL(WGB_ERR_INVALID, "wgb_radix_sort: output views overlap the inputs or each other");
    DeviceGuard g(pass->ctx->device);
    return launch_radix_sort(pass, kin, vin, iks->size[0], (const uint32_t *)n_sort->ptr, sorting_bits, kout, vout);
}

wgb_status wgb_reduce(wgb_pass *pass, wgb_reduce_op op, const wgb_buffer *value, const wgb_view_shape *vs,
                      wgb_buffer *result) {
    if (!pass || !value || !vs || !result) WGB_FAIL(WGB_ERR_INVALID, "wgb_reduce: null argument");
    if ((int)op < 0 || (int)op > WGB_RED_SQNORM) WGB_FAIL(WGB_ERR_INVALID, "wgb_reduce: unknown op %d", (int)op);
    if (any_zero_buffer({value, result})) return WGB_OK;
    if (result->bytes < 4) WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_reduce: result buffer smaller than one f32");
    WGB_TRY(check_view(value, *vs, 4, "reduce value", true));
    DeviceGuard g(pass->ctx->device);
    return launch_reduce(pass, (int)op, (const float *)value->ptr + vs->offset, nullptr, vs->size[0], (float *)result->ptr);
}

wgb_status wgb_dot(wgb_pass *pass, const wgb_buffer *a, const wgb_view_shape *sa, const wgb_buffer *b,
                   const wgb_view_shape *sb, wgb_buffer *result) {
    if (!pass || !a || !b || !sa || !sb || !result) WGB_FAIL(WGB_ERR_INVALID, "wgb_dot: null argument");
    if (sa->size[0] != sb->size[0]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Dot: dimension mismatch. (%u vs %u)", sa->size[0], sb->size[0]);
    if (any_zero_buffer({a, b, result})) return WGB_OK;
    if (result->bytes < 4) WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_dot: result buffer smaller than one f32");
    WGB_TRY(check_view(a, *sa, 4, "dot a", true));
    WGB_TRY(check_view(b, *sb, 4, "dot b", true));
    DeviceGuard g(pass->ctx->device);
    return launch_reduce(pass, 5 /* dot */, (const float *)a->ptr + sa->offset, (const float *)b->ptr + sb->offset,
                         sa->size[0], (float *)result->ptr);
}

wgb_status wgb_reduce_columns(wgb_pass *pass, wgb_reduce_op op, const wgb_buffer *m, const wgb_view_shape *ms,
                              wgb_buffer *out, const wgb_view_shape *os) {
    if (!pass || !m || !ms || !out || !os) WGB_FAIL(WGB_ERR_INVALID, "wgb_reduce_columns: null argument");
    if ((int)op < 0 || (int)op > WGB_RED_SQNORM) WGB_FAIL(WGB_ERR_INVALID, "wgb_reduce_columns: unknown op %d", (int)op);
    if ((uint64_t)ms->size[1] * ms->size[2] != os->size[0])
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Reduce-columns: dimension mismatch. (%u x %u columns vs %u outputs)", ms->size[1],
                 ms->size[2], os->size[0]);
    if (any_zero_buffer({m, out}) || os->size[0] == 0) return WGB_OK;
    WGB_TRY(check_view(m, *ms, 4, "reduce_columns m"));
    WGB_TRY(check_view(out, *os, 4, "reduce_columns out", true));
    DeviceGuard g(pass->ctx->device);
    return launch_reduce_columns(pass, (int)op, (const float *)m->ptr, *ms, (float *)out->ptr + os->offset);
}

wgb_status wgb_fill_uniform(wgb_pass *pass, wgb_buffer *buf, const wgb_view_shape *s, wgb_dtype dt, uint64_t seed,
                            uint32_t row0, uint32_t col0) {
    if (!pass || !buf || !s) WGB_FAIL(WGB_ERR_INVALID, "wgb_fill_uniform: null argument");
    if (buf->bytes == 0 || view_extent(*s) == 0) return WGB_OK;
    WGB_TRY(check_view(buf, *s, dtype_size(dt), "fill_uniform"));
    DeviceGuard g(pass->ctx->device);
    return launch_fill_uniform(pass, buf->ptr, *s, dt, seed, row0, col0);
}

// A row-major [r x c] view (shape.wgsl:49-53: element (i, j) at offset + i * stride + j) addresses the same memory as the
// column-major [c x r] view with the same stride / stride_mat / offset, i.e. its transpose.
static wgb_view_shape transposed_shape(const wgb_view_shape &s) {
    wgb_view_shape t = s;
    t.size[0] = s.size[1];
    t.size[1] = s.size[0];
    return t;
}
static bool bad_ordering(wgb_ordering o) { return (int)o != WGB_COLUMN_MAJOR && (int)o != WGB_ROW_MAJOR; }

wgb_status wgb_gemv(wgb_pass *pass, wgb_gemv_variant variant, wgb_buffer *out, const wgb_view_shape *so,
                    const wgb_buffer *m, const wgb_view_shape *sm, const wgb_buffer *v, const wgb_view_shape *sv) {
    return wgb_gemv_ord(pass, variant, out, so, m, sm, WGB_COLUMN_MAJOR, v, sv);
}

wgb_status wgb_gemv_ord(wgb_pass *pass, wgb_gemv_variant variant, wgb_buffer *out, const wgb_view_shape *so, const wgb_buffer *m,
                        const wgb_view_shape *sm, wgb_ordering m_ord, const wgb_buffer *v, const wgb_view_shape *sv) {
    return wgb_gemv_op(pass, variant, out, so, m, sm, m_ord, v, sv, -1, nullptr, nullptr);
}

wgb_status wgb_gemv_op(wgb_pass *pass, wgb_gemv_variant variant, wgb_buffer *out, const wgb_view_shape *so, const wgb_buffer *m,
                       const wgb_view_shape *sm, wgb_ordering m_ord, const wgb_buffer *v, const wgb_view_shape *sv, int op,
                       const wgb_buffer *e, const wgb_view_shape *se) {
    if (!pass || !out || !so || !m || !sm || !v || !sv) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv: null argument");
    if ((int)variant < 0 || (int)variant > WGB_GEMV_TR_FAST) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv: unknown variant %d", (int)variant);
    if (bad_ordering(m_ord)) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv: unknown ordering %d", (int)m_ord);
    if (op >= 0) {
        if (!e || !se) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_op: null operand");
        if (op >= WGB_OP_COPY) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_op: op %d (Copy would discard the product)", op);
        if (se->size[0] != so->size[0] || se->size[1] < so->size[1] || se->size[2] < so->size[2])   // op_assign.rs:82-86
            WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Op-assign: dimension mismatch. (%u vs %u)", so->size[0], se->size[0]);
    }
    bool tr = variant == WGB_GEMV_TR || variant == WGB_GEMV_TR_FAST;
    const uint32_t m_rows = tr ? sm->size[1] : sm->size[0];
    const uint32_t m_cols = tr ? sm->size[0] : sm->size[1];
    if (m_cols != sv->size[0])  // gemv.rs:89
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemv: dimension mismatch. (matrix cols %u vs vector rows %u)", m_cols, sv->size[0]);
    if (m_rows != so->size[0])  // gemv.rs:90
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemv: dimension mismatch. (matrix rows %u vs out rows %u)", m_rows, so->size[0]);
    // gemv.rs:99-104: GemvTrFast silently becomes GemvTr when m.nrows % 128 != 0; gemv.rs:122: the *_fast
    // variants assert out_nrows % 4 == 0.  Here all variants share one kernel pair that is valid for every
    // shape, so the fallback is moot; the assert is kept so callers see the reference's panic.
    const bool fast = variant == WGB_GEMV_FAST || (variant == WGB_GEMV_TR_FAST && sm->size[0] % 128u == 0);
    if (fast && so->size[0] % 4u != 0)
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemv: the fast variants require out rows %% 4 == 0 (got %u)", so->size[0]);
    // The reference reads v / m with the *output's* column and batch counts (grid = [.., out_ncols, out_nmats],
    // gemv.rs:136) and never checks them against v / m: make the implied requirement explicit.
    if (sv->size[1] < so->size[1] || sv->size[2] < so->size[2] || sm->size[2] < so->size[2])
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemv: out has %u columns x %u matrices but v has %u x %u and m has %u matrices",
                 so->size[1], so->size[2], sv->size[1], sv->size[2], sm->size[2]);
    if (any_zero_buffer({out, m, v}) || view_extent(*so) == 0) return WGB_OK;
    wgb_view_shape sv_used = *sv, sm_used = *sm;
    if (m_ord == WGB_ROW_MAJOR) {   // the same memory read as the column-major transpose
        sm_used = transposed_shape(*sm);
        tr = !tr;
    }
    sv_used.size[1] = so->size[1];
    sv_used.size[2] = so->size[2];
    sm_used.size[2] = so->size[2];
    WGB_TRY(check_view(out, *so, 4, "gemv out"));
    WGB_TRY(check_view(m, sm_used, 4, "gemv m"));
    WGB_TRY(check_view(v, sv_used, 4, "gemv v"));
    wgb_view_shape se_used{};
    if (op >= 0) {
        if (e->bytes == 0) return WGB_OK;
        se_used = *se;
        se_used.size[1] = so->size[1];
        se_used.size[2] = so->size[2];
        WGB_TRY(check_view(e, se_used, 4, "gemv operand"));
        // the operand may BE the output view (out = m*v + out: each element is read, then written, by one thread); any other
        // overlap would race
        const char *ob = (const char *)out->ptr + 4ull * so->offset, *eb = (const char *)e->ptr + 4ull * se_used.offset;
        const bool same = ob == eb && so->stride == se_used.stride && (so->size[2] == 1 || so->stride_mat == se_used.stride_mat);
        const uint64_t on = 4 * (view_extent(*so) - so->offset), en = 4 * (view_extent(se_used) - se_used.offset);
        if (!same && ob < eb + en && eb < ob + on)
            WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_op: the operand overlaps the output without being the same view");
    }
    DeviceGuard g(pass->ctx->device);
    return launch_gemv(pass, tr, (float *)out->ptr, *so, (const float *)m->ptr, sm_used, (const float *)v->ptr, sv_used, op,
                       op >= 0 ? (const float *)e->ptr : nullptr, op >= 0 ? &se_used : nullptr);
}

// Gemv -> Reduce in one launch (the reference chain is Gemv::dispatch then Reduce::dispatch on the product
// vector, gemv.rs:64-137 + reduce.rs:100-113).  The product never reaches a caller buffer: it goes to a context scratch vector and
// the last CTA of the GEMV reduces it with exactly the tree wgb_reduce would use, so the scalar is bit-identical to the two-dispatch
// chain (with a 16-byte aligned `out`).
wgb_status wgb_gemv_reduce(wgb_pass *pass, wgb_gemv_variant variant, wgb_reduce_op rop, wgb_buffer *result, const wgb_buffer *m,
                           const wgb_view_shape *sm, wgb_ordering m_ord, const wgb_buffer *v, const wgb_view_shape *sv) {
    if (!pass || !result || !m || !sm || !v || !sv) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_reduce: null argument");
    if ((int)variant < 0 || (int)variant > WGB_GEMV_TR_FAST) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_reduce: unknown variant %d", (int)variant);
    if ((int)rop < 0 || (int)rop > WGB_RED_SQNORM) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_reduce: unknown reduce op %d", (int)rop);
    if (bad_ordering(m_ord)) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemv_reduce: unknown ordering %d", (int)m_ord);
    if (sm->size[2] != 1 || sv->size[1] != 1 || sv->size[2] != 1)
        WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemv_reduce: one matrix and one vector (the reference reduces a GpuVectorView)");
    bool tr = variant == WGB_GEMV_TR || variant == WGB_GEMV_TR_FAST;
    const uint32_t m_rows = tr ? sm->size[1] : sm->size[0];
    const uint32_t m_cols = tr ? sm->size[0] : sm->size[1];
    if (m_cols != sv->size[0])  // gemv.rs:89
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemv: dimension mismatch. (matrix cols %u vs vector rows %u)", m_cols, sv->size[0]);
    if (any_zero_buffer({result, m, v}) || m_rows == 0) return WGB_OK;
    if (result->bytes < 4) WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_gemv_reduce: result buffer smaller than one f32");
    wgb_view_shape sm_used = *sm;
    if (m_ord == WGB_ROW_MAJOR) {
        sm_used = transposed_shape(*sm);
        tr = !tr;
    }
    WGB_TRY(check_view(m, sm_used, 4, "gemv m"));
    WGB_TRY(check_view(v, *sv, 4, "gemv v", true));
    wgb_ctx *ctx = pass->ctx;
    DeviceGuard g(ctx->device);
    void *w = nullptr;
    WGB_TRY(workspace_reserve(ctx, 7, (size_t)m_rows * 4 + 16, &w));
    const wgb_view_shape so{{m_rows, 1, 1}, m_rows, m_rows, 0};
    const bool fuse = m_cols > 0 && (uint32_t)reduce_grid_for(ctx, (int)rop, m_rows) <= kGemvReduceMaxGrid;
    WGB_TRY(launch_gemv(pass, tr, (float *)w, so, (const float *)m->ptr, sm_used, (const float *)v->ptr, *sv, -1, nullptr, nullptr,
                        fuse ? (int)rop : -1, (float *)result->ptr));
    if (!fuse) return launch_reduce(pass, (int)rop, (const float *)w, nullptr, m_rows, (float *)result->ptr);   // same tree, second launch
    return WGB_OK;
}

static wgb_status gemm_common(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *so,
                              const wgb_buffer *m1, const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2,
                              wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode mode, int ep_op, const wgb_buffer *e,
                              const wgb_view_shape *se, wgb_ordering out_ord = WGB_COLUMN_MAJOR,
                              wgb_ordering m1_ord = WGB_COLUMN_MAJOR, wgb_ordering m2_ord = WGB_COLUMN_MAJOR) {
    if (!pass || !out || !so || !m1 || !s1 || !m2 || !s2) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm: null argument");
    if ((int)variant < 0 || (int)variant > WGB_GEMM_TR_FAST) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm: unknown variant %d", (int)variant);
    if ((in_dtype != WGB_F32 && in_dtype != WGB_BF16) || (out_dtype != WGB_F32 && out_dtype != WGB_BF16))
        WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm: unsupported dtype");
    if ((int)mode < 0 || (int)mode > WGB_F32_SIMT) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm: unknown f32 mode %d", (int)mode);
    if (bad_ordering(out_ord) || bad_ordering(m1_ord) || bad_ordering(m2_ord)) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm: unknown ordering");
    const bool ro = out_ord == WGB_ROW_MAJOR, r1 = m1_ord == WGB_ROW_MAJOR, r2 = m2_ord == WGB_ROW_MAJOR;
    const bool tr = variant == WGB_GEMM_TR || variant == WGB_GEMM_TR_FAST;
    const uint32_t m_rows = tr ? s1->size[1] : s1->size[0];
    const uint32_t m_cols = tr ? s1->size[0] : s1->size[1];
    // gemm.rs:91-95, same order and message
    if (m_cols != s2->size[0]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (m1 cols %u vs m2 rows %u)", m_cols, s2->size[0]);
    if (m_rows != so->size[0]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (m1 rows %u vs out rows %u)", m_rows, so->size[0]);
    if (so->size[1] != s2->size[1]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (out cols %u vs m2 cols %u)", so->size[1], s2->size[1]);
    if (so->size[2] != s1->size[2]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (out mats %u vs m1 mats %u)", so->size[2], s1->size[2]);
    if (so->size[2] != s2->size[2]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (out mats %u vs m2 mats %u)", so->size[2], s2->size[2]);
    pass->last_gemm_path = 0;
    if (any_zero_buffer({out, m1, m2}) || view_extent(*so) == 0) return WGB_OK;
    WGB_TRY(check_view(out, ro ? transposed_shape(*so) : *so, dtype_size(out_dtype), "gemm out"));
    WGB_TRY(check_view(m1, r1 ? transposed_shape(*s1) : *s1, dtype_size(in_dtype), "gemm m1"));
    WGB_TRY(check_view(m2, r2 ? transposed_shape(*s2) : *s2, dtype_size(in_dtype), "gemm m2"));
    // Canonical device problem: C (column-major) = A * B with "K contiguous in A's memory" (tr) and "N contiguous in B's
    // memory" (b_nmajor) as the only layout facts.  A row-major out is computed as out^T = m2^T * op(m1)^T.
    const bool a_k = tr != r1;   // K is the contiguous axis of m1's memory
    const bool b_n = r2;         // N is the contiguous axis of m2's memory
    GemmProblem g{};
    g.K = m_cols;
    g.nmats = so->size[2];
    g.c = out->ptr;
    g.c_off = so->offset;
    g.ldc = so->stride;
    g.sc = so->stride_mat;
    if (!ro) {
        g.tr = a_k;
        g.b_nmajor = b_n;
        g.M = so->size[0];
        g.N = so->size[1];
        g.a = m1->ptr; g.a_off = s1->offset; g.lda = s1->stride; g.sa = s1->stride_mat;
        g.b = m2->ptr; g.b_off = s2->offset; g.ldb = s2->stride; g.sb = s2->stride_mat;
    } else {
        g.tr = !b_n;
        g.b_nmajor = !a_k;
        g.M = so->size[1];
        g.N = so->size[0];
        g.a = m2->ptr; g.a_off = s2->offset; g.lda = s2->stride; g.sa = s2->stride_mat;
        g.b = m1->ptr; g.b_off = s1->offset; g.ldb = s1->stride; g.sb = s1->stride_mat;
    }
    g.in_dtype = in_dtype;
    g.out_dtype = out_dtype;
    if (ep_op >= 0) {
        if (!e || !se) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_op: null operand");
        if (ep_op == WGB_OP_COPY || ep_op > WGB_OP_COPY) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_op: op must be Add, Sub, Mul or Div");
        if (se->size[0] != so->size[0] || se->size[1] != so->size[1] || se->size[2] != so->size[2])
            WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Op-assign: dimension mismatch. (operand %u x %u x %u vs out %u x %u x %u)", se->size[0],
                     se->size[1], se->size[2], so->size[0], so->size[1], so->size[2]);
        if (e->bytes == 0) return WGB_OK;
        WGB_TRY(check_view(e, ro ? transposed_shape(*se) : *se, dtype_size(out_dtype), "gemm_op operand"));   // e is ordered like out
        g.ep_op = ep_op;
        g.e = e->ptr;
        g.e_off = se->offset;
        g.lde = se->stride;
        g.se = se->stride_mat;
    }
    DeviceGuard dg(pass->ctx->device);
    return gemm_dispatch(pass, g, mode);
}

}  // extern "C"  (gemm_common is internal)
extern "C" {

wgb_status wgb_gemm_ex(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *so,
                       const wgb_buffer *m1, const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2,
                       wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode mode) {
    return gemm_common(pass, variant, out, so, m1, s1, m2, s2, in_dtype, out_dtype, mode, -1, nullptr, nullptr);
}

wgb_status wgb_gemm_op(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *so, const wgb_buffer *m1,
                       const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2, wgb_dtype in_dtype,
                       wgb_dtype out_dtype, wgb_f32_mode mode, wgb_op_assign_variant op, const wgb_buffer *operand,
                       const wgb_view_shape *operand_shape) {
    if ((int)op < 0) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_op: unknown op");
    return gemm_common(pass, variant, out, so, m1, s1, m2, s2, in_dtype, out_dtype, mode, (int)op, operand, operand_shape);
}


// Gemm -> Reduce along one axis in one pass over the operands: the product itself is never stored.
wgb_status wgb_gemm_reduce(wgb_pass *pass, wgb_gemm_variant variant, int axis, wgb_reduce_op rop, wgb_buffer *result,
                           const wgb_view_shape *rs, const wgb_buffer *m1, const wgb_view_shape *s1, const wgb_buffer *m2,
                           const wgb_view_shape *s2, wgb_dtype in_dtype, wgb_f32_mode mode) {
    if (!pass || !result || !rs || !m1 || !s1 || !m2 || !s2) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_reduce: null argument");
    if ((int)variant < 0 || (int)variant > WGB_GEMM_TR_FAST) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_reduce: unknown variant %d", (int)variant);
    if (axis != 1 && axis != 2) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_reduce: axis must be 1 (one value per column) or 2 (one per row)");
    if ((int)rop < 0 || (int)rop > WGB_RED_SQNORM) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_reduce: unknown reduce op %d", (int)rop);
    if (in_dtype != WGB_F32 && in_dtype != WGB_BF16) WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm_reduce: unsupported dtype");
    if ((int)mode < 0 || (int)mode > WGB_F32_SIMT) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_reduce: unknown f32 mode %d", (int)mode);
    if (s1->size[2] != 1 || s2->size[2] != 1) WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm_reduce: one matrix per operand");
    const bool tr = variant == WGB_GEMM_TR || variant == WGB_GEMM_TR_FAST;
    const uint32_t M = tr ? s1->size[1] : s1->size[0], K = tr ? s1->size[0] : s1->size[1], N = s2->size[1];
    if (K != s2->size[0]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (m1 cols %u vs m2 rows %u)", K, s2->size[0]);
    const uint32_t n_out = axis == 1 ? N : M;
    if (rs->size[0] != n_out)
        WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm-reduce: dimension mismatch. (result has %u elements, the product has %u %s)", rs->size[0], n_out,
                 axis == 1 ? "columns" : "rows");
    pass->last_gemm_path = 0;
    if (any_zero_buffer({result, m1, m2}) || M == 0 || N == 0) return WGB_OK;
    WGB_TRY(check_view(result, *rs, 4, "gemm_reduce result", true));
    WGB_TRY(check_view(m1, *s1, dtype_size(in_dtype), "gemm m1"));
    WGB_TRY(check_view(m2, *s2, dtype_size(in_dtype), "gemm m2"));
    GemmProblem g{};
    g.tr = tr;
    g.M = M; g.N = N; g.K = K; g.nmats = 1;
    g.a = m1->ptr; g.a_off = s1->offset; g.lda = s1->stride; g.sa = s1->stride_mat;
    g.b = m2->ptr; g.b_off = s2->offset; g.ldb = s2->stride; g.sb = s2->stride_mat;
    g.in_dtype = in_dtype;
    g.out_dtype = WGB_F32;
    DeviceGuard dg(pass->ctx->device);
    return gemm_reduce_dispatch(pass, g, mode, axis, (int)rop, (float *)result->ptr + rs->offset);
}

wgb_status wgb_gemm_ord(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *so, wgb_ordering out_ord,
                        const wgb_buffer *m1, const wgb_view_shape *s1, wgb_ordering m1_ord, const wgb_buffer *m2,
                        const wgb_view_shape *s2, wgb_ordering m2_ord, wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode mode,
                        int op, const wgb_buffer *operand, const wgb_view_shape *operand_shape) {
    return gemm_common(pass, variant, out, so, m1, s1, m2, s2, in_dtype, out_dtype, mode, op < 0 ? -1 : op, operand, operand_shape,
                       out_ord, m1_ord, m2_ord);
}

// Host-buffer GEMM.  Three in-order streams: uploads (A, then the B panels back to back, so the host->device link never
// waits for a kernel), the queue (panel GEMMs, each gated on its B panel), downloads (panel j of C leaves while panel j+1 of B
// arrives).  Two device operand slots alternate between calls, so with the enqueue form the download of product i runs under
// the upload of product i+1: the link is busy in both directions and the steady-state cost of a product is its link time.
// Measured on B200 (PCIe gen5 x16, bf16 4096^3: 64 MiB up at 54.6 GB/s = 1.23 ms, 32 MiB down = 0.60 ms): a download that runs
// under an upload gets about half the link rate, so one blocking call cannot hide C under B (1.58 ms whatever the panel count);
// back-to-back enqueued products are bounded by the 1.28 ms both directions need together.
static wgb_status gemm_host_enqueue(wgb_ctx *ctx, wgb_gemm_variant variant, uint32_t M, uint32_t N, uint32_t K, void *out_host,
                                    const void *m1_host, const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype,
                                    wgb_f32_mode mode, int n_panels, uint32_t default_panels) {
    if (!ctx || !out_host || !m1_host || !m2_host) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_host: null argument");
    if ((int)variant < 0 || (int)variant > WGB_GEMM_TR_FAST) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_host: unknown variant");
    if (M == 0 || N == 0) return WGB_OK;
    const bool tr = variant == WGB_GEMM_TR || variant == WGB_GEMM_TR_FAST;
    const size_t es = dtype_size(in_dtype), os = dtype_size(out_dtype);
    const size_t a_bytes = (size_t)M * K * es, b_bytes = (size_t)K * N * es, c_bytes = (size_t)M * N * os;
    if ((uint64_t)M * K > 0xFFFFFFFFull || (uint64_t)K * N > 0xFFFFFFFFull || (uint64_t)M * N > 0xFFFFFFFFull)
        WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm_host: operands exceed u32 element indexing");
    DeviceGuard dg(ctx->device);
    HostGemmState &hs = ctx->host_gemm;
    if (!ctx->h2d_stream) WGB_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    for (auto &e : hs.done)
        if (!e) WGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const size_t b_off = (a_bytes + 255) & ~(size_t)255, c_off = (b_off + b_bytes + 255) & ~(size_t)255;
    const size_t slot_bytes = (c_off + c_bytes + 255) & ~(size_t)255;
    char *w = nullptr;
    int slot = 0;
    WGB_TRY(host_gemm_slot(ctx, slot_bytes, &w, &slot));
    char *dA = w, *dB = dA + b_off, *dC = dA + c_off;
    // column panels: whole 256-column tiles
    // every extra copy costs ~10-20 us of link time (measured: enqueued products 1.35 / 1.41 / 1.52 / 1.66 ms at 2 / 4 / 8 / 16
    // panels, blocking calls 1.65 / 1.58 / 1.61 / 1.66 ms): few panels when products overlap each other, more when one call
    // has to overlap with itself
    uint32_t np = n_panels > 0 ? (uint32_t)n_panels : default_panels;
    uint32_t width = (N + np - 1) / np;
    width = (width + 255u) & ~255u;
    np = (N + width - 1) / width;
    while (hs.evs.size() < 2 * (size_t)np) {
        cudaEvent_t e;
        WGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hs.evs.push_back(e);
    }
    wgb_pass pass;
    pass.ctx = ctx;
    pass.stream = ctx->stream;
    if (K > 0) WGB_CUDA(cudaMemcpyAsync(dA, m1_host, a_bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    for (uint32_t j = 0; j < np; ++j) {
        const uint32_t n0 = j * width, nc = (N - n0) < width ? (N - n0) : width;
        if (K > 0)
            WGB_CUDA(cudaMemcpyAsync(dB + (size_t)n0 * K * es, (const char *)m2_host + (size_t)n0 * K * es, (size_t)nc * K * es,
                                     cudaMemcpyHostToDevice, ctx->h2d_stream));
        WGB_CUDA(cudaEventRecord(hs.evs[j], ctx->h2d_stream));
        WGB_CUDA(cudaStreamWaitEvent(ctx->stream, hs.evs[j], 0));
        GemmProblem g{};
        g.tr = tr;
        g.M = M; g.N = nc; g.K = K; g.nmats = 1;
        g.a = dA; g.b = dB; g.c = dC;
        g.a_off = 0; g.b_off = (uint64_t)n0 * K; g.c_off = (uint64_t)n0 * M;
        g.lda = tr ? K : M; g.ldb = K; g.ldc = M;
        g.sa = (uint64_t)M * K; g.sb = (uint64_t)K * N; g.sc = (uint64_t)M * N;
        g.in_dtype = in_dtype; g.out_dtype = out_dtype;
        WGB_TRY(gemm_dispatch(&pass, g, mode));
        WGB_CUDA(cudaEventRecord(hs.evs[np + j], ctx->stream));
        WGB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, hs.evs[np + j], 0));
        WGB_CUDA(cudaMemcpyAsync((char *)out_host + (size_t)n0 * M * os, dC + (size_t)n0 * M * os, (size_t)nc * M * os,
                                 cudaMemcpyDeviceToHost, ctx->comm_stream));
    }
    WGB_CUDA(cudaEventRecord(hs.done[slot], ctx->comm_stream));
    hs.pending[slot] = true;
    return WGB_OK;
}

wgb_status wgb_gemm_host_enqueue(wgb_ctx *ctx, wgb_gemm_variant variant, uint32_t M, uint32_t N, uint32_t K, void *out_host,
                                 const void *m1_host, const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype,
                                 wgb_f32_mode mode, int n_panels) {
    return gemm_host_enqueue(ctx, variant, M, N, K, out_host, m1_host, m2_host, in_dtype, out_dtype, mode, n_panels, 2u);
}

wgb_status wgb_gemm_host_flush(wgb_ctx *ctx) {
    if (!ctx) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_host_flush: null context");
    DeviceGuard dg(ctx->device);
    HostGemmState &hs = ctx->host_gemm;
    for (int s = 0; s < 2; ++s)
        if (hs.pending[s]) WGB_CUDA(cudaStreamWaitEvent(ctx->stream, hs.done[s], 0));
    return WGB_OK;
}

wgb_status wgb_gemm_host(wgb_ctx *ctx, wgb_gemm_variant variant, uint32_t M, uint32_t N, uint32_t K, void *out_host,
                         const void *m1_host, const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode mode,
                         int n_panels) {
    WGB_TRY(gemm_host_enqueue(ctx, variant, M, N, K, out_host, m1_host, m2_host, in_dtype, out_dtype, mode, n_panels, 4u));
    if (M == 0 || N == 0) return WGB_OK;
    DeviceGuard dg(ctx->device);
    WGB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    WGB_CUDA(cudaStreamSynchronize(ctx->stream));
    WGB_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
    ctx->host_gemm.pending[0] = ctx->host_gemm.pending[1] = false;
    return WGB_OK;
}

wgb_status wgb_gemm(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *so,
                    const wgb_buffer *m1, const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2) {
    return wgb_gemm_ex(pass, variant, out, so, m1, s1, m2, s2, WGB_F32, WGB_F32, WGB_F32_AUTO);
}

}  // extern "C"
