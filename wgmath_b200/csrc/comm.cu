// comm.cu — the one multi-GPU element of the path: a GEMM sharded by output-row block across the GPUs of
// one box, followed by the all-gather of C over NVLink / NVSwitch (BASELINE configs[4]).
// The reference has no multi-device code at all (one wgpu::Device, one Queue: wgcore/src/gpu.rs:7-12).
//
// One process per GPU; rank p owns rows [p*M/P, (p+1)*M/P) of m1 and of the product, m2 is replicated.
// Layout choice (SURVEY.md §8(e) option (i)): every rank computes its block into a *contiguous* column-major
// [M/P x N] panel; the gathered result is P panels back to back, i.e. exactly a reference GpuCube view
// size = [M/P, N, P], stride = M/P, stride_mat = (M/P)*N (tensor.rs:465-481) — no extra pass over C.
//
// Overlap: N is cut into column chunks.  Chunk c's GEMM runs on the queue stream; as soon as it finishes
// (event), the exchange of chunk c (grouped ncclSend/ncclRecv — an all-gather with arbitrary placement, so
// the chunk lands directly inside each panel) runs on the context's comm stream while chunk c+1 computes.
// The GEMM leaves `WGB_COMM_SM_MARGIN` SMs free so the NCCL CTAs can be co-resident instead of queueing
// behind the persistent GEMM CTAs.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the host framework already loaded, else the
// system one), so libwgebra_b200.so itself has no link-time NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace wgb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;   // optional
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("WGB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
        SYM(GetUniqueId, "ncclGetUniqueId");
        SYM(CommInitRank, "ncclCommInitRank");
        SYM(CommInitRankConfig, "ncclCommInitRankConfig");
        SYM(CommDestroy, "ncclCommDestroy");
        SYM(GroupStart, "ncclGroupStart");
        SYM(GroupEnd, "ncclGroupEnd");
        SYM(Send, "ncclSend");
        SYM(Recv, "ncclRecv");
        SYM(AllGather, "ncclAllGather");
        SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
    });
    return api;
}

struct CommState {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    int sm_margin = 0;
    bool overlapping = false;   // inside wgb_gemm_row_sharded: GEMMs leave sm_margin SMs to the concurrent NCCL exchange
    std::vector<cudaEvent_t> events;
    cudaEvent_t done = nullptr;
};

#define WGB_NCCL(expr)                                                                                          \
    do {                                                                                                        \
        ncclResult_t _r = (expr);                                                                               \
        if (_r != ncclSuccess)                                                                                  \
            WGB_FAIL(WGB_ERR_NCCL, "%s failed: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(_r) : "?"); \
    } while (0)

void comm_destroy(wgb_ctx *ctx) {
    CommState *c = ctx->comm;
    if (!c) return;
    for (auto e : c->events) cudaEventDestroy(e);
    if (c->done) cudaEventDestroy(c->done);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
    ctx->comm = nullptr;
}

int comm_sm_margin(const wgb_ctx *ctx) { return ctx->comm && ctx->comm->nranks > 1 && ctx->comm->overlapping ? ctx->comm->sm_margin : 0; }

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace wgb

using namespace wgb;

extern "C" {

wgb_status wgb_comm_get_unique_id(void *id_out) {
    static_assert(sizeof(ncclUniqueId) == WGB_COMM_ID_BYTES, "ncclUniqueId size changed");
    if (!id_out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    if (!nccl().ok) WGB_FAIL(WGB_ERR_NCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    ncclUniqueId id;
    WGB_NCCL(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return WGB_OK;
}

wgb_status wgb_comm_init_rank(wgb_ctx *ctx, int nranks, int rank, const void *id_bytes) {
    if (!ctx || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) WGB_FAIL(WGB_ERR_INVALID, "wgb_comm_init_rank: bad argument");
    if (!nccl().ok) WGB_FAIL(WGB_ERR_NCCL, "libnccl.so.2 could not be loaded");
    if (ctx->comm) comm_destroy(ctx);
    DeviceGuard g(ctx->device);
    CommState *c = new CommState();
    c->nranks = nranks;
    c->rank = rank;
    const int max_ctas = env_int("WGB_COMM_MAX_CTAS", 8);
    c->sm_margin = env_int("WGB_COMM_SM_MARGIN", max_ctas);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r;
    if (nccl().CommInitRankConfig) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.maxCTAs = max_ctas;   // the exchange needs a sliver of NVLink bandwidth; keep its SM footprint small
        cfg.minCTAs = 1;
        r = nccl().CommInitRankConfig(&c->comm, nranks, id, rank, &cfg);
    } else {
        r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    }
    if (r != ncclSuccess) {
        delete c;
        WGB_FAIL(WGB_ERR_NCCL, "ncclCommInitRank failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    }
    cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
    ctx->comm = c;
    return WGB_OK;
}

wgb_status wgb_comm_destroy(wgb_ctx *ctx) {
    if (ctx) {
        DeviceGuard g(ctx->device);
        cudaStreamSynchronize(ctx->comm_stream);
        comm_destroy(ctx);
    }
    return WGB_OK;
}

wgb_status wgb_gemm_row_sharded(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_buffer *m1,
                                const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2, wgb_dtype in_dtype,
                                wgb_dtype out_dtype, wgb_f32_mode mode, int n_chunks) {
    if (!pass || !out || !m1 || !s1 || !m2 || !s2) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded: null argument");
    wgb_ctx *ctx = pass->ctx;
    CommState *cs = ctx->comm;
    const int P = cs ? cs->nranks : 1, rank = cs ? cs->rank : 0;
    const bool tr = variant == WGB_GEMM_TR || variant == WGB_GEMM_TR_FAST;
    if (s1->size[2] != 1 || s2->size[2] != 1) WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm_row_sharded: batched operands are not supported");
    const uint32_t Mloc = tr ? s1->size[1] : s1->size[0];
    const uint32_t K = tr ? s1->size[0] : s1->size[1];
    const uint32_t N = s2->size[1];
    if (K != s2->size[0]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (m1 cols %u vs m2 rows %u)", K, s2->size[0]);
    const size_t os = dtype_size(out_dtype), es = dtype_size(in_dtype);
    const uint64_t panel = (uint64_t)Mloc * N;
    if (panel * P * os > out->bytes)
        WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_gemm_row_sharded: gathered output needs %llu bytes, buffer has %zu",
                 (unsigned long long)(panel * P * os), out->bytes);
    if (panel > 0xFFFFFFFFull) WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm_row_sharded: panel exceeds u32 indexing");
    WGB_TRY(check_view(m1, *s1, es, "sharded gemm m1"));
    WGB_TRY(check_view(m2, *s2, es, "sharded gemm m2"));
    if (Mloc == 0 || N == 0) return WGB_OK;
    DeviceGuard dg(ctx->device);

    // column chunks: multiples of 256 columns (whole BLOCK_N tiles), default 8 chunks when exchanging
    uint32_t nch = n_chunks > 0 ? (uint32_t)n_chunks : (P > 1 ? 8u : 1u);
    uint32_t width = (N + nch - 1) / nch;
    width = (width + 255u) & ~255u;
    nch = (N + width - 1) / width;
    if (cs)
        while (cs->events.size() < nch) {
            cudaEvent_t e;
            WGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            cs->events.push_back(e);
        }

    struct Overlap {   // scoped: only the GEMMs of this call run beside NCCL CTAs
        CommState *c;
        explicit Overlap(CommState *c_) : c(c_) { if (c) c->overlapping = true; }
        ~Overlap() { if (c) c->overlapping = false; }
    } overlap(P > 1 ? cs : nullptr);
    for (uint32_t c = 0; c < nch; ++c) {
        const uint32_t n0 = c * width, nc = (N - n0) < width ? (N - n0) : width;
        GemmProblem g{};
        g.tr = tr;
        g.M = Mloc; g.N = nc; g.K = K; g.nmats = 1;
        g.a = m1->ptr; g.b = m2->ptr; g.c = out->ptr;
        g.a_off = s1->offset;
        g.b_off = (uint64_t)s2->offset + (uint64_t)n0 * s2->stride;
        g.c_off = (uint64_t)rank * panel + (uint64_t)n0 * Mloc;
        g.lda = s1->stride; g.ldb = s2->stride; g.ldc = Mloc;
        g.sa = s1->stride_mat; g.sb = s2->stride_mat; g.sc = panel;
        g.in_dtype = in_dtype; g.out_dtype = out_dtype;
        WGB_TRY(gemm_dispatch(pass, g, mode));
        if (P > 1) {
            WGB_CUDA(cudaEventRecord(cs->events[c], pass->stream));
            WGB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, cs->events[c], 0));
            const size_t bytes = (size_t)nc * Mloc * os;
            char *base = (char *)out->ptr;
            const size_t chunk_off = (size_t)n0 * Mloc * os;
            WGB_NCCL(nccl().GroupStart());
            for (int d = 1; d < P; ++d) {
                const int to = (rank + d) % P, from = (rank - d + P) % P;
                WGB_NCCL(nccl().Send(base + (size_t)rank * panel * os + chunk_off, bytes, ncclInt8, to, cs->comm, ctx->comm_stream));
                WGB_NCCL(nccl().Recv(base + (size_t)from * panel * os + chunk_off, bytes, ncclInt8, from, cs->comm, ctx->comm_stream));
            }
            WGB_NCCL(nccl().GroupEnd());
            count_launch(ctx);
        }
    }
    if (P > 1) {
        // later work on the queue sees the gathered panels
        WGB_CUDA(cudaEventRecord(cs->done, ctx->comm_stream));
        WGB_CUDA(cudaStreamWaitEvent(pass->stream, cs->done, 0));
    }
    return WGB_OK;
}

// ------------------------------------------------------------------------------------------------------------
// fused GEMM + all-gather over peer memory
//
// Protocol (all counters are monotonic call numbers, "epochs", of one group; buffer of epoch e = e mod depth):
//   ready[q] (lives on every rank r, written by rank q): rank q's buffers for epochs <= value may be overwritten by its peers.
//            Rank q publishes it at the start of each of its calls.  depth 1: value = this call's epoch (a peer can only start
//            storing step e once q itself has reached step e).  depth >= 2: value = epoch + 1 — the buffer the *next* step will
//            use was last read by consumers that q's in-order queue has already passed, so peers may run one step ahead.
//   done[q]  (lives on every rank r, written by the last CTA of rank q's GEMM): rank q's panel of epoch `value` is complete in
//            r's buffer.  wgb_peer_gather_wait(e) = all done[q] >= e.
//   Validity of the result of call e: depth 1 / 2 -> until call e + 1 is issued; depth 3 -> until call e + 2 is issued, which is
//   what lets the wait (and the consumer) trail the GEMMs by one call so that a slow rank does not stall the others every step.
// ------------------------------------------------------------------------------------------------------------
}  // extern "C"

struct wgb_peer_gather {
    wgb_ctx *ctx = nullptr;
    int nranks = 1, rank = 0, depth = 1;
    size_t data_bytes = 0;          // one gathered buffer (256-byte multiple)
    char *local = nullptr;          // [depth][data_bytes][signal block]: cudaMalloc, or the caller's symmetric allocation
    char *peer[wgb::kMaxPeers] = {};
    char *mc = nullptr;             // multicast mapping of the same allocation (external groups only; null: none)
    bool external = false;          // memory and mappings belong to the caller (wgb_peer_gather_create_external)
    bool ipc_opened[wgb::kMaxPeers] = {};
    bool connected = false;
    unsigned int epoch = 0;
    long long timeout_cycles = 0;   // bounded flag waits (0: unbounded)
    wgb_buffer view[3];             // non-owning wgb_buffers over the local gathered buffers
    size_t signal_off() const { return (size_t)depth * data_bytes; }
    char *buffer_of(unsigned int e) const { return local + (size_t)(e % (unsigned)depth) * data_bytes; }
};

namespace wgb {
// signal block layout (u32 words): [0..8) ready[q], [16..24) done[q], [32] cta counter
constexpr size_t kSignalBytes = 4096;
constexpr int kReadyOff = 0, kDoneOff = 16, kCtaOff = 32;

__global__ void signal_ready_kernel(FusedGather f, unsigned int value) {
    const int q = threadIdx.x;
    if (q < f.nranks && q != f.rank) {
        __threadfence_system();
        // done_remote[q] points at rank q's done array; its ready array sits kDoneOff words before it
        unsigned int *ready_q = f.done_remote[q] - kDoneOff + kReadyOff;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ready_q + f.rank), "r"(value) : "memory");
    }
}
__global__ void wait_done_kernel(const unsigned int *done_local, int nranks, int rank, unsigned int epoch, long long timeout) {
    const int q = threadIdx.x;
    if (q < nranks && q != rank) {
        const long long t0 = clock64();
        unsigned int v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(done_local + q) : "memory");
            if (timeout > 0 && clock64() - t0 > timeout) __trap();
        } while ((int)(v - epoch) < 0);
    }
    __threadfence_system();
}

// Flag waits are bounded so that a lost peer surfaces as a CUDA error instead of a hung GPU; WGB_PEER_TIMEOUT_MS (default 60 s,
// 0 = wait forever) must cover the longest legitimate skew between ranks (first-call module loads, uploads ahead of a step).
static long long peer_timeout_cycles(const wgb_ctx *ctx) {
    const long long ms = env_int("WGB_PEER_TIMEOUT_MS", 60000);
    return ms <= 0 ? 0 : ms * (long long)ctx->prop.clockRate;   // clockRate is in kHz = cycles per ms
}
}  // namespace wgb

extern "C" {

wgb_status wgb_peer_gather_create_ex(wgb_ctx *ctx, int nranks, int rank, size_t gathered_bytes, int depth, wgb_peer_gather **out) {
    if (!ctx || !out || nranks < 1 || nranks > wgb::kMaxPeers || rank < 0 || rank >= nranks)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_create: bad argument (at most %d ranks)", wgb::kMaxPeers);
    if (depth < 1 || depth > 3) WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_create: depth must be 1, 2 or 3");
    DeviceGuard g(ctx->device);
    wgb_peer_gather *pg = new wgb_peer_gather();
    pg->ctx = ctx;
    pg->nranks = nranks;
    pg->rank = rank;
    pg->depth = depth;
    pg->data_bytes = (gathered_bytes + 255) & ~(size_t)255;
    pg->timeout_cycles = wgb::peer_timeout_cycles(ctx);
    {   // Load the flag kernels now.  With lazy module loading the first launch of a kernel may synchronise the context; if that
        // first launch came while a GEMM of this process spins on a flag that only a later launch can set (several ranks driven
        // by one host thread), the process would deadlock.
        cudaFuncAttributes fa;
        (void)cudaFuncGetAttributes(&fa, wgb::signal_ready_kernel);
        (void)cudaFuncGetAttributes(&fa, wgb::wait_done_kernel);
    }
    cudaError_t e = cudaMalloc((void **)&pg->local, pg->signal_off() + wgb::kSignalBytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        delete pg;
        WGB_FAIL(WGB_ERR_OOM, "peer gather: allocation of %d x %zu bytes failed: %s", depth, gathered_bytes, cudaGetErrorString(e));
    }
    cudaMemset(pg->local + pg->signal_off(), 0, wgb::kSignalBytes);
    cudaDeviceSynchronize();
    pg->peer[rank] = pg->local;
    pg->connected = nranks == 1;
    for (int d = 0; d < depth; ++d) {
        pg->view[d].ctx = ctx;
        pg->view[d].ptr = pg->local + (size_t)d * pg->data_bytes;
        pg->view[d].bytes = gathered_bytes;
        pg->view[d].owned = false;
        pg->view[d].usage = WGB_USAGE_STORAGE | WGB_USAGE_COPY_SRC | WGB_USAGE_COPY_DST;
    }
    wgb::ctx_retain(ctx);
    *out = pg;
    return WGB_OK;
}

size_t wgb_peer_gather_region_bytes(size_t gathered_bytes, int depth) {
    return (size_t)(depth < 1 ? 1 : depth) * ((gathered_bytes + 255) & ~(size_t)255) + wgb::kSignalBytes;
}

// A group over memory the caller allocated symmetrically on every rank (same size, same layout) and mapped both ways: peer_bases[q]
// = this process's unicast mapping of rank q's region, multicast_base = the region through an NVSwitch multicast object bound on
// all ranks (or null).  Any allocator that can do this works (cuMem* + cuMulticast* directly; torch's symmetric memory in the
// Python mirror).  The region must hold wgb_peer_gather_region_bytes() bytes; nothing is freed or unmapped by the library.
wgb_status wgb_peer_gather_create_external(wgb_ctx *ctx, int nranks, int rank, size_t gathered_bytes, int depth, void *const *peer_bases,
                                           void *multicast_base, wgb_peer_gather **out) {
    if (!ctx || !out || !peer_bases || nranks < 1 || nranks > wgb::kMaxPeers || rank < 0 || rank >= nranks)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_create_external: bad argument (at most %d ranks)", wgb::kMaxPeers);
    if (depth < 1 || depth > 3) WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_create_external: depth must be 1, 2 or 3");
    for (int q = 0; q < nranks; ++q)
        if (!peer_bases[q] || ((uintptr_t)peer_bases[q] & 255u)) WGB_FAIL(WGB_ERR_INVALID, "peer base %d is null or not 256-byte aligned", q);
    if ((uintptr_t)multicast_base & 255u) WGB_FAIL(WGB_ERR_INVALID, "multicast base is not 256-byte aligned");
    DeviceGuard g(ctx->device);
    wgb_peer_gather *pg = new wgb_peer_gather();
    pg->ctx = ctx;
    pg->nranks = nranks;
    pg->rank = rank;
    pg->depth = depth;
    pg->external = true;
    pg->data_bytes = (gathered_bytes + 255) & ~(size_t)255;
    pg->timeout_cycles = wgb::peer_timeout_cycles(ctx);
    cudaFuncAttributes fa;
    (void)cudaFuncGetAttributes(&fa, wgb::signal_ready_kernel);
    (void)cudaFuncGetAttributes(&fa, wgb::wait_done_kernel);
    pg->local = (char *)peer_bases[rank];
    for (int q = 0; q < nranks; ++q) pg->peer[q] = (char *)peer_bases[q];
    pg->mc = (char *)multicast_base;
    cudaError_t e = cudaMemset(pg->local + pg->signal_off(), 0, wgb::kSignalBytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        delete pg;
        WGB_FAIL(WGB_ERR_CUDA, "wgb_peer_gather_create_external: cannot clear the flag block: %s", cudaGetErrorString(e));
    }
    pg->connected = true;   // (the caller synchronises the ranks after creation: a peer's first flag must not be cleared by a late memset)
    for (int d = 0; d < depth; ++d) {
        pg->view[d].ctx = ctx;
        pg->view[d].ptr = pg->local + (size_t)d * pg->data_bytes;
        pg->view[d].bytes = gathered_bytes;
        pg->view[d].owned = false;
        pg->view[d].usage = WGB_USAGE_STORAGE | WGB_USAGE_COPY_SRC | WGB_USAGE_COPY_DST;
    }
    wgb::ctx_retain(ctx);
    *out = pg;
    return WGB_OK;
}

wgb_status wgb_peer_gather_create(wgb_ctx *ctx, int nranks, int rank, size_t gathered_bytes, wgb_peer_gather **out) {
    return wgb_peer_gather_create_ex(ctx, nranks, rank, gathered_bytes, 1, out);
}

wgb_status wgb_peer_gather_export(wgb_peer_gather *pg, void *handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == WGB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size changed");
    if (!pg || !handle_out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(pg->ctx->device);
    cudaIpcMemHandle_t h;
    WGB_CUDA(cudaIpcGetMemHandle(&h, pg->local));
    memcpy(handle_out, &h, sizeof h);
    return WGB_OK;
}

wgb_status wgb_peer_gather_connect(wgb_peer_gather *pg, const void *handles) {
    if (!pg || !handles) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(pg->ctx->device);
    for (int q = 0; q < pg->nranks; ++q) {
        if (q == pg->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)q * WGB_IPC_HANDLE_BYTES, sizeof h);
        void *ptr = nullptr;
        WGB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        pg->peer[q] = (char *)ptr;
        pg->ipc_opened[q] = true;
    }
    pg->connected = true;
    return WGB_OK;
}

// Same-process form of connect: the groups of all ranks live in this process (one context per rank; the ranks may share one
// device, which is how the whole protocol is exercised on a single-GPU box, or sit on several devices driven by one process).
wgb_status wgb_peer_gather_connect_local(wgb_peer_gather *pg, wgb_peer_gather *const *groups) {
    if (!pg || !groups) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(pg->ctx->device);
    for (int q = 0; q < pg->nranks; ++q) {
        wgb_peer_gather *o = groups[q];
        if (!o || o->nranks != pg->nranks || o->rank != q || o->depth != pg->depth || o->data_bytes != pg->data_bytes)
            WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_connect_local: group %d does not match (ranks, rank, depth or size)", q);
        if (q == pg->rank) {
            if (o != pg) WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_connect_local: groups[rank] must be the group itself");
            continue;
        }
        if (o->ctx->device != pg->ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                WGB_FAIL(WGB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", o->ctx->device, cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
        pg->peer[q] = o->local;
        pg->ipc_opened[q] = false;
    }
    pg->connected = true;
    return WGB_OK;
}

wgb_status wgb_peer_gather_buffer_at(wgb_peer_gather *pg, int calls_back, wgb_buffer **out) {
    if (!pg || !out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    if (calls_back < 0 || calls_back >= pg->depth || (unsigned)calls_back > pg->epoch)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_buffer_at: the group keeps %d result(s); %d calls back is gone", pg->depth, calls_back);
    *out = &pg->view[(pg->epoch - (unsigned)calls_back) % (unsigned)pg->depth];
    return WGB_OK;
}

wgb_status wgb_peer_gather_buffer(wgb_peer_gather *pg, wgb_buffer **out) { return wgb_peer_gather_buffer_at(pg, 0, out); }

wgb_status wgb_peer_gather_wait(wgb_pass *pass, wgb_peer_gather *pg, int calls_back) {
    if (!pass || !pg) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    if (calls_back < 0 || (unsigned)calls_back > pg->epoch) WGB_FAIL(WGB_ERR_INVALID, "wgb_peer_gather_wait: no such call");
    if (pg->nranks == 1 || pg->epoch == 0) return WGB_OK;
    DeviceGuard dg(pg->ctx->device);
    unsigned int *sig = reinterpret_cast<unsigned int *>(pg->local + pg->signal_off());
    wgb::wait_done_kernel<<<1, 32, 0, pass->stream>>>(sig + wgb::kDoneOff, pg->nranks, pg->rank, pg->epoch - (unsigned)calls_back,
                                                      pg->timeout_cycles);
    count_launch(pg->ctx);
    WGB_CUDA(cudaGetLastError());
    return WGB_OK;
}

}  // extern "C"
namespace wgb {
// ------------------------------------------------------------------------------------------------------------
// Link micro-benchmark (diagnostics): how fast can the SMs of one GPU push a buffer to `ndst` destinations —
//   mode 0  st.global.v4 (16 B per lane, 512 B per warp instruction) to unicast addresses (local or peer mappings)
//   mode 1  multimem.st.global.v4.f32 to ONE multicast address (NVSwitch replicates to every GPU of the group)
//   mode 2  cp.async.bulk shared -> global (TMA bulk stores, 16 KiB each) to unicast addresses
// Each CTA streams its slice of `src` once per destination.  Used to size the fused all-gather (profiles/).
// ------------------------------------------------------------------------------------------------------------
struct LinkDsts {
    char *p[kMaxPeers];
};
__global__ void __launch_bounds__(256) link_stream_kernel(int mode, LinkDsts dsts, int ndst, const char *__restrict__ src, size_t bytes) {
    extern __shared__ __align__(128) unsigned char link_smem[];
    const size_t chunk = 16384;
    const size_t nchunks = bytes / chunk;
    if (mode == 2) {
        // fill two staging buffers once (the payload does not matter for the rate), then keep two bulk groups in flight
        for (int i = threadIdx.x; i < (int)(2 * chunk / 16); i += blockDim.x)
            reinterpret_cast<uint4 *>(link_smem)[i] = reinterpret_cast<const uint4 *>(src)[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t k = 0;
            for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++k) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(link_smem + (k & 1u) * chunk);
                for (int d = 0; d < ndst; ++d)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dsts.p[(d + blockIdx.x) % ndst] + c * chunk),
                                 "r"(s), "r"((uint32_t)chunk)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    const size_t n16 = bytes / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src) + i);
        if (mode == 1) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dsts.p[0] + i * 16), "f"(__uint_as_float(v.x)),
                         "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                         : "memory");
        } else {
            for (int d = 0; d < ndst; ++d) {
                uint4 *q = reinterpret_cast<uint4 *>(dsts.p[(d + blockIdx.x) % ndst]) + i;
                asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(q), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
        }
    }
}
}  // namespace wgb
extern "C" {

wgb_status wgb_debug_link_stream(wgb_ctx *ctx, int mode, void *const *dsts, int ndst, const void *src, size_t bytes, int ctas, int iters,
                                 float *ms_per_iter) {
    if (!ctx || !dsts || !src || !ms_per_iter || ndst < 1 || ndst > wgb::kMaxPeers || mode < 0 || mode > 2 || iters < 1)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_debug_link_stream: bad argument");
    if (bytes % 32768 != 0) WGB_FAIL(WGB_ERR_INVALID, "wgb_debug_link_stream: bytes must be a multiple of 32 KiB");
    DeviceGuard g(ctx->device);
    wgb::LinkDsts d{};
    for (int i = 0; i < ndst; ++i) d.p[i] = (char *)dsts[i];
    if (ctas <= 0) ctas = ctx->prop.multiProcessorCount * 2;
    static bool attr_set = false;
    if (!attr_set) {
        WGB_CUDA(cudaFuncSetAttribute(wgb::link_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        attr_set = true;
    }
    cudaEvent_t e0, e1;
    WGB_CUDA(cudaEventCreate(&e0));
    WGB_CUDA(cudaEventCreate(&e1));
    wgb::link_stream_kernel<<<ctas, 256, 32768, ctx->stream>>>(mode, d, ndst, (const char *)src, bytes);   // warm-up
    WGB_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; ++i) wgb::link_stream_kernel<<<ctas, 256, 32768, ctx->stream>>>(mode, d, ndst, (const char *)src, bytes);
    WGB_CUDA(cudaEventRecord(e1, ctx->stream));
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) WGB_FAIL(WGB_ERR_CUDA, "wgb_debug_link_stream: %s", cudaGetErrorString(e));
    count_launch(ctx, iters + 1);
    *ms_per_iter = ms / iters;
    return WGB_OK;
}

// Diagnostics: a snapshot of this rank's flag block taken on a private stream, so it can be read while the queues are busy
// (or stuck): out[0..8) = ready[q], out[8..16) = done[q], out[16] = CTA counter, out[17] = this rank's call count (host side).
wgb_status wgb_peer_gather_debug_flags(wgb_peer_gather *pg, unsigned int *out /* 18 words */) {
    if (!pg || !out) WGB_FAIL(WGB_ERR_INVALID, "null argument");
    DeviceGuard g(pg->ctx->device);
    cudaStream_t s = nullptr;
    WGB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    unsigned int host[64] = {};
    cudaError_t e = cudaMemcpyAsync(host, pg->local + pg->signal_off(), sizeof host, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) WGB_FAIL(WGB_ERR_CUDA, "flag snapshot failed: %s", cudaGetErrorString(e));
    for (int q = 0; q < 8; ++q) {
        out[q] = host[wgb::kReadyOff + q];
        out[8 + q] = host[wgb::kDoneOff + q];
    }
    out[16] = host[wgb::kCtaOff];
    out[17] = pg->epoch;
    return WGB_OK;
}

// First half of tearing a group down across processes: drain this rank's queues and unmap the peers' buffers.  CUDA requires
// every importer to have closed its mapping before the exporter frees the memory, so the ranks call this, meet at a barrier of
// their launcher, and only then call wgb_peer_gather_destroy.
wgb_status wgb_peer_gather_disconnect(wgb_peer_gather *pg) {
    if (!pg) return WGB_OK;
    DeviceGuard g(pg->ctx->device);
    // the host-operand form leaves downloads of the gathered buffer on the side streams
    cudaStreamSynchronize(pg->ctx->stream);
    cudaStreamSynchronize(pg->ctx->comm_stream);
    if (pg->ctx->h2d_stream) cudaStreamSynchronize(pg->ctx->h2d_stream);
    for (int q = 0; q < pg->nranks; ++q)
        if (q != pg->rank && pg->peer[q]) {
            if (pg->ipc_opened[q]) cudaIpcCloseMemHandle(pg->peer[q]);
            pg->peer[q] = nullptr;
            pg->ipc_opened[q] = false;
        }
    pg->connected = pg->nranks == 1;
    return WGB_OK;
}

wgb_status wgb_peer_gather_destroy(wgb_peer_gather *pg) {
    if (!pg) return WGB_OK;
    wgb_peer_gather_disconnect(pg);
    DeviceGuard g(pg->ctx->device);
    if (pg->local && !pg->external) cudaFree(pg->local);
    wgb_ctx *ctx = pg->ctx;
    delete pg;
    wgb::ctx_release(ctx);
    return WGB_OK;
}

wgb_status wgb_gemm_row_sharded_fused_ex(wgb_pass *pass, wgb_gemm_variant variant, wgb_peer_gather *pg, const wgb_buffer *m1,
                                         const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2,
                                         wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode mode, uint32_t flags) {
    if (!pass || !pg || !m1 || !s1 || !m2 || !s2) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded_fused: null argument");
    if (!pg->connected) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded_fused: peer group is not connected");
    if (flags & ~(uint32_t)WGB_GATHER_NO_WAIT) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded_fused: unknown flags 0x%x", flags);
    wgb_ctx *ctx = pass->ctx;
    const int P = pg->nranks, rank = pg->rank;
    const bool tr = variant == WGB_GEMM_TR || variant == WGB_GEMM_TR_FAST;
    if (s1->size[2] != 1 || s2->size[2] != 1) WGB_FAIL(WGB_ERR_UNSUPPORTED, "wgb_gemm_row_sharded_fused: batched operands are not supported");
    const uint32_t Mloc = tr ? s1->size[1] : s1->size[0];
    const uint32_t K = tr ? s1->size[0] : s1->size[1];
    const uint32_t N = s2->size[1];
    if (K != s2->size[0]) WGB_FAIL(WGB_ERR_DIM_MISMATCH, "Gemm: dimension mismatch. (m1 cols %u vs m2 rows %u)", K, s2->size[0]);
    const size_t os = dtype_size(out_dtype), es = dtype_size(in_dtype);
    const uint64_t panel = (uint64_t)Mloc * N;
    if (panel * P * os > pg->view[0].bytes)
        WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_gemm_row_sharded_fused: gathered output needs %llu bytes, group holds %zu",
                 (unsigned long long)(panel * P * os), pg->view[0].bytes);
    WGB_TRY(check_view(m1, *s1, es, "sharded gemm m1"));
    WGB_TRY(check_view(m2, *s2, es, "sharded gemm m2"));
    if (Mloc == 0 || N == 0) return WGB_OK;
    DeviceGuard dg(ctx->device);

    const unsigned int epoch = pg->epoch + 1;
    const size_t buf_off = (size_t)(epoch % (unsigned)pg->depth) * pg->data_bytes;
    FusedGather f;
    f.nranks = P;
    f.rank = rank;
    f.epoch = epoch;
    f.timeout = pg->timeout_cycles;
    f.mc_c = pg->mc ? pg->mc + buf_off : nullptr;
    for (int q = 0; q < P; ++q) {
        f.peer_c[q] = pg->peer[q] + buf_off;
        f.done_remote[q] = reinterpret_cast<unsigned int *>(pg->peer[q] + pg->signal_off()) + wgb::kDoneOff;
    }
    unsigned int *sig = reinterpret_cast<unsigned int *>(pg->local + pg->signal_off());
    f.ready_local = sig + wgb::kReadyOff;
    f.cta_counter = sig + wgb::kCtaOff;

    GemmProblem g{};
    g.fused = &f;
    g.tr = tr;
    g.M = Mloc; g.N = N; g.K = K; g.nmats = 1;
    g.a = m1->ptr; g.b = m2->ptr; g.c = pg->local + buf_off;
    g.a_off = s1->offset; g.b_off = s2->offset;
    g.c_off = (uint64_t)rank * panel;
    g.lda = s1->stride; g.ldb = s2->stride; g.ldc = Mloc;
    g.sa = s1->stride_mat; g.sb = s2->stride_mat; g.sc = panel;
    g.in_dtype = in_dtype; g.out_dtype = out_dtype;
    // A rank that cannot run the peer-storing kernel must fail *before* it tells its peers that it has reached this step:
    // every later call of every rank would otherwise disagree about the epoch.
    if (P > 1 && !gemm_fused_eligible(g, mode))
        WGB_FAIL(WGB_ERR_UNSUPPORTED, "fused all-gather needs the tensor-core GEMM path (16-byte aligned operand views, problem >= 96^3, "
                                      "f32 mode other than SIMT)");
    pg->epoch = epoch;
    if (P > 1) {
        // "my buffers for epochs <= value may be overwritten": see the protocol note above
        wgb::signal_ready_kernel<<<1, 32, 0, pass->stream>>>(f, pg->depth >= 2 ? epoch + 1 : epoch);
        count_launch(ctx);
    }
    WGB_TRY(gemm_dispatch(pass, g, mode));
    if (P > 1 && !(flags & WGB_GATHER_NO_WAIT)) WGB_TRY(wgb_peer_gather_wait(pass, pg, 0));   // all panels have landed here
    return WGB_OK;
}

wgb_status wgb_gemm_row_sharded_fused(wgb_pass *pass, wgb_gemm_variant variant, wgb_peer_gather *pg, const wgb_buffer *m1,
                                      const wgb_view_shape *s1, const wgb_buffer *m2, const wgb_view_shape *s2, wgb_dtype in_dtype,
                                      wgb_dtype out_dtype, wgb_f32_mode mode) {
    return wgb_gemm_row_sharded_fused_ex(pass, variant, pg, m1, s1, m2, s2, in_dtype, out_dtype, mode, 0);
}

// Host-buffer form of the fused sharded GEMM, enqueued (the N > 1 counterpart of wgb_gemm_host_enqueue, abi.cu): this rank's A
// block and B are uploaded on the upload stream into one of two alternating device slots, the fused GEMM + all-gather runs on
// the queue once they have landed, and the result leaves on the download stream — this rank's own [M_local x N] panel only
// (download_all == 0: the ranks of one box assemble C in host memory, every byte crosses a host link once) or the whole gathered
// cube.  The upload of product i + 1 therefore runs under the GEMM and the download of product i.
// B crosses the host links once per box when the context has a communicator of the same ranks (wgb_comm_init_rank): every rank
// uploads only its 1/P column slice and the slices are all-gathered in place over NVLink (ncclAllGather on the queue, between
// the GEMMs: the persistent GEMM leaves no SM for a concurrent collective, and the step is host-link bound either way).
// WGB_SHARD_B_UPLOAD=0 restores the whole-B upload on every rank.
// Ordering specific to the shared gathered buffers: the queue waits for the download of the product whose buffer this call (or,
// through the ready flag, a peer running one step ahead) is about to overwrite.
wgb_status wgb_gemm_row_sharded_fused_host_enqueue(wgb_ctx *ctx, wgb_gemm_variant variant, wgb_peer_gather *pg, uint32_t M_local,
                                                   uint32_t N, uint32_t K, void *out_host, const void *m1_local_host,
                                                   const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode mode,
                                                   int download_all) {
    if (!ctx || !pg || !out_host || !m1_local_host || !m2_host)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded_fused_host_enqueue: null argument");
    if (pg->ctx != ctx) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded_fused_host_enqueue: the peer group belongs to another context");
    if ((int)variant < 0 || (int)variant > WGB_GEMM_TR_FAST) WGB_FAIL(WGB_ERR_INVALID, "unknown variant");
    if (M_local == 0 || N == 0 || K == 0) WGB_FAIL(WGB_ERR_INVALID, "wgb_gemm_row_sharded_fused_host_enqueue: empty operands");
    if ((uint64_t)M_local * K > 0xFFFFFFFFull || (uint64_t)K * N > 0xFFFFFFFFull || (uint64_t)M_local * N * pg->nranks > 0xFFFFFFFFull)
        WGB_FAIL(WGB_ERR_UNSUPPORTED, "operands exceed u32 element indexing");
    const bool tr = variant == WGB_GEMM_TR || variant == WGB_GEMM_TR_FAST;
    const size_t es = dtype_size(in_dtype), os = dtype_size(out_dtype);
    const size_t a_bytes = (size_t)M_local * K * es, b_bytes = (size_t)K * N * es, panel_bytes = (size_t)M_local * N * os;
    DeviceGuard dg(ctx->device);
    HostGemmState &hs = ctx->host_gemm;
    if (!ctx->h2d_stream) WGB_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    for (auto &e : hs.done)
        if (!e) WGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    while (hs.evs.size() < 2) {
        cudaEvent_t e;
        WGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hs.evs.push_back(e);
    }
    const size_t b_off = (a_bytes + 255) & ~(size_t)255, slot_bytes = (b_off + b_bytes + 255) & ~(size_t)255;
    char *w = nullptr;
    int slot = 0;
    WGB_TRY(host_gemm_slot(ctx, slot_bytes, &w, &slot));
    char *dA = w, *dB = dA + b_off;
    WGB_CUDA(cudaMemcpyAsync(dA, m1_local_host, a_bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    const int P = pg->nranks;
    const bool split_b = env_int("WGB_SHARD_B_UPLOAD", 1) != 0 && P > 1 && ctx->comm && ctx->comm->comm && ctx->comm->nranks == P &&
                         ctx->comm->rank == pg->rank && nccl().ok && nccl().AllGather && N % (uint32_t)P == 0;
    const size_t slice = b_bytes / (size_t)P;
    if (split_b)
        WGB_CUDA(cudaMemcpyAsync(dB + (size_t)pg->rank * slice, (const char *)m2_host + (size_t)pg->rank * slice, slice,
                                 cudaMemcpyHostToDevice, ctx->h2d_stream));
    else
        WGB_CUDA(cudaMemcpyAsync(dB, m2_host, b_bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    WGB_CUDA(cudaEventRecord(hs.evs[0], ctx->h2d_stream));
    WGB_CUDA(cudaStreamWaitEvent(ctx->stream, hs.evs[0], 0));
    if (split_b) {
        WGB_NCCL(nccl().AllGather(dB + (size_t)pg->rank * slice, dB, slice, ncclInt8, ctx->comm->comm, ctx->stream));
        count_launch(ctx);
    }
    // the download that last read the buffer about to be handed out (depth 3: two products back, i.e. this slot's own)
    const int reader = pg->depth == 3 ? slot : slot ^ 1;
    if (hs.pending[reader]) WGB_CUDA(cudaStreamWaitEvent(ctx->stream, hs.done[reader], 0));
    wgb_buffer bA, bB;
    bA.ctx = bB.ctx = ctx;
    bA.owned = bB.owned = false;
    bA.ptr = dA; bA.bytes = a_bytes;
    bB.ptr = dB; bB.bytes = b_bytes;
    const wgb_view_shape s1 = tr ? wgb_view_shape{{K, M_local, 1}, K, K * M_local, 0} : wgb_view_shape{{M_local, K, 1}, M_local, M_local * K, 0};
    const wgb_view_shape s2 = wgb_view_shape{{K, N, 1}, K, K * N, 0};
    wgb_pass pass;
    pass.ctx = ctx;
    pass.stream = ctx->stream;
    WGB_TRY(wgb_gemm_row_sharded_fused(&pass, variant, pg, &bA, &s1, &bB, &s2, in_dtype, out_dtype, mode));
    WGB_CUDA(cudaEventRecord(hs.evs[1], ctx->stream));
    WGB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, hs.evs[1], 0));
    const char *result = pg->buffer_of(pg->epoch);
    if (download_all)
        WGB_CUDA(cudaMemcpyAsync(out_host, result, panel_bytes * pg->nranks, cudaMemcpyDeviceToHost, ctx->comm_stream));
    else
        WGB_CUDA(cudaMemcpyAsync(out_host, result + (size_t)pg->rank * panel_bytes, panel_bytes, cudaMemcpyDeviceToHost, ctx->comm_stream));
    WGB_CUDA(cudaEventRecord(hs.done[slot], ctx->comm_stream));
    hs.pending[slot] = true;
    return WGB_OK;
}

}  // extern "C"
