// gemm_tc_kernel.cuh — the tcgen05 GEMM kernel template and its launcher (see gemm_tc.cu for the design notes).
// Included by gemm_tc.cu (host side: tensor maps, tail plan, dispatch) and by the gemm_tc_inst_*.cu translation units,
// which explicitly instantiate launch_sel<> for one operand family each so the ~44 kernel variants compile in parallel.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace wgb {
namespace tc {


// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
        "r"(rank)
        : "memory");
}
// Bounded wait: a broken pipeline traps (sticky CUDA error reported to the caller) instead of hanging the GPU.
// The first probe carries no clock read: in steady state the barrier has usually completed already, and the waits sit in the
// single-warp TMA / MMA issue loops whose instruction count per k-block bounds the tensor-pipe rate.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// cluster-scope acquire: the barrier is also arrived on (release.cluster) by the other CTA of the pair, whose shared-memory
// writes the waiter's MMAs go on to read
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, long long limit) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity))
        if (limit > 0 && clock64() - t0 > limit) __trap();
}
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, long long limit) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (limit > 0 && clock64() - t0 > limit) __trap();
}
// `limit` (SM cycles, 0 = unbounded) is only read on the slow path: a plain GEMM gives up after a few seconds, a GEMM whose
// epilogue waits for other GPUs (fused all-gather) must outlast the peer timeout.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, long long limit) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, limit);
}
// One lane of a converged warp (elect.sync): the form the compiler turns into a single predicated UTCHMMA / UTMALDG instead of
// the vote-and-retry loop it wraps around uniform-datapath instructions in code it must assume divergent.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred)::"memory");
    return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// cta_group::2 form: data lands in this CTA's smem, complete_tx is signalled on the *leader* CTA's barrier
// (peer bit of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(dst),
        "l"(tm), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Bulk tensor store smem -> global (SASS: UTMASTG).  The box is clipped against the tensor's extents, so ragged tiles need no
// predication; the destination may be a peer GPU's memory mapped over NVLink.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups still have shared-memory reads outstanding
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// all of this thread's bulk groups have completed (their global writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA) that reads them next
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// 16 bytes to a multicast address: the NVSwitch writes them at the same offset of every GPU bound to the multicast object
__device__ __forceinline__ void multimem_st16(void *mc_addr, const uint4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]
template <int KIND, int CG>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) {
        if (CG == 1)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                         "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                         "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                         : "memory");
    } else {
        if (CG == 1)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                         "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                         "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                         : "memory");
    }
}
// Arrive on an mbarrier when all previously issued MMAs have completed (implies fence::before_thread_sync).
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"((uint16_t)3)
                     : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// descriptors (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ---------------------------------------------------------------------------------------------
// layout_type: 2 = SWIZZLE_128B (16-byte swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms: required for
// MN-major 32-bit operands, matching TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
// The descriptor as two 32-bit halves: only the start-address field of the low word changes inside the main loop, so the
// issue loop adds a 16-byte-unit offset to a precomputed low word instead of rebuilding 64-bit fields per MMA.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint64_t pack_desc(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

constexpr uint32_t make_idesc(int kind, bool a_mn_major, bool b_mn_major, int m, int n) {
    const uint32_t fmt = kind == 0 ? 1u /* BF16 */ : 2u /* TF32 */;
    return (1u << 4)                      // accumulator format F32
           | (fmt << 7) | (fmt << 10)     // A / B element format
           | ((a_mn_major ? 1u : 0u) << 15)  // A major-ness (0 = K-major)
           | ((b_mn_major ? 1u : 0u) << 16)  // B major-ness (0 = K-major)
           | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcArgs {
    void *c;
    uint64_t ldc, sc;
    uint32_t M, N, K, nmats;
    uint32_t tiles_m, tiles_n, num_kb, total_tiles;
    // tail split-K: tiles [full_tiles, total_tiles) are each cut into `split` K ranges of `kb_per_split` k-blocks, so the
    // last, partially filled wave of tiles still occupies every SM.  Work units = full_tiles + (total - full) * split.
    uint32_t full_tiles, split, kb_per_split, total_units;
    // tail N-split (default): tiles [full_tiles, total_tiles) are instead cut into `nsplit` column strips of `tail_bn`
    // columns, each a complete (narrower) output tile: the last wave fills the machine and needs no fix-up.
    uint32_t nsplit, tail_bn, idesc_tail;
    uint32_t epi_tma;        // 1: the epilogue stages 128 x 32 blocks in shared memory and leaves with TMA bulk stores (dst maps)
                             // 2: same staging, the block leaves with multimem.st to mc_dst (NVSwitch multicast: one store
                             //    instruction lands in the gathered buffer of every rank, this one included)
    char *mc_dst;            // multicast address of element 0 of this rank's output view (epi_tma == 2)
    int ep_op;               // fused element-wise epilogue (-1: none): out = acc (op) e
    const void *ep;          // operand view base (element 0 of the view), element type = TOut
    uint64_t ep_ld, ep_sm;
    // fused reduction of the product (wgb_gemm_reduce): the product is not stored; every 32-column block of a tile leaves as
    // partial results instead — red_axis 1: per column over the 32 rows of each epilogue warp -> red_partials[row / 32][N];
    // red_axis 2: per row over the block's 32 columns -> red_partials[col / 32][M].  A fold kernel combines them in index order.
    uint32_t red_axis;       // 0: none
    int red_op;              // wgb_reduce_op
    float *red_partials;
    uint32_t fused_split;    // 3xTF32: the kernel splits the raw f32 tiles itself (FS instantiation) instead of reading hi / lo copies
    uint32_t debug_skip;     // diagnostics only (WGB_TC_DEBUG_SKIP): bit 0 = do not load A tiles, bit 1 = do not load B tiles
    unsigned long long *trace;   // diagnostics (wgb_debug_tc_trace): 8 words per cluster, null = off
    float *ws;               // [tail tile][split][cta rank][BN][128] f32 partial accumulators
    unsigned int *counters;  // [tail tile][cta rank] arrival tickets (left at zero)
    // fused all-gather over peer memory (npeers == 1: plain GEMM, dst[0] == c)
    uint32_t npeers, my_rank, epoch, handshake;
    uint32_t first_dst;                       // destination the store loops start with: rank + 1, so that at any moment the ranks of
                                              // a box aim at different peers (no hot ingress port), the local copy last
    uint32_t store_mask;                      // diagnostics (WGB_FUSED_DEBUG_STORE_MASK): bit d clear = skip destination d
    long long peer_timeout;                   // bound of the ready-flag wait in SM cycles (0: unbounded)
    long long mbar_timeout;                   // bound of every mbarrier wait in SM cycles (0: unbounded)
    char *dst[kMaxPeers];                     // where this rank's panel lives in rank d's gathered buffer
    unsigned int *ready_local;                // ready_local[q] >= epoch: rank q's buffer may be overwritten
    unsigned int *done_remote[kMaxPeers];     // rank q's done array; entry [my_rank] <- epoch when all stores are out
    unsigned int *cta_counter;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Output tensor maps of the TMA-store epilogue: one per destination (the local buffer; with the fused all-gather also this
// rank's panel inside every peer's gathered buffer).  dims = {M, N, nmats}, box = {128 rows, kEpiCols columns}, no swizzle.
struct TcDstMaps {
    CUtensorMap m[kMaxPeers];
};
constexpr int kEpiCols = 32;   // columns per staged block = one tcgen05.ld.32x32b.x32

struct WorkUnit {
    uint32_t tile, kb0, kb1, split_idx;   // split_idx == 0xFFFFFFFF: whole K range
    uint32_t n_off, bn;                   // column strip inside the tile (n_off = 0, bn = BLOCK_N for a whole tile)
    bool narrow;
};
__device__ __forceinline__ WorkUnit decode_unit(uint32_t u, const TcArgs &a, uint32_t block_n) {
    WorkUnit w;
    w.n_off = 0; w.bn = block_n; w.narrow = false;
    if (u < a.full_tiles) {
        w.tile = u; w.kb0 = 0; w.kb1 = a.num_kb; w.split_idx = 0xFFFFFFFFu;
    } else if (a.nsplit > 1) {
        const uint32_t v = u - a.full_tiles;
        w.tile = a.full_tiles + v / a.nsplit;
        w.kb0 = 0; w.kb1 = a.num_kb; w.split_idx = 0xFFFFFFFFu;
        w.n_off = (v % a.nsplit) * a.tail_bn;
        w.bn = a.tail_bn;
        w.narrow = true;
    } else {
        const uint32_t v = u - a.full_tiles;
        w.tile = a.full_tiles + v / a.split;
        w.split_idx = v % a.split;
        w.kb0 = w.split_idx * a.kb_per_split;
        w.kb1 = min(a.num_kb, w.kb0 + a.kb_per_split);
    }
    return w;
}

constexpr int kBlockM = 128;        // rows per CTA
constexpr int kRowBytes = 128;      // bytes of K (K-major) or of M (MN-major) per smem row = swizzle span
constexpr int kATileBytes = kBlockM * kRowBytes;  // 16 KiB, both major-nesses
constexpr int kNumThreads = 256;
constexpr int kSuperM = 8;          // m-tiles per rasterisation group
constexpr int kSmemLimit = 227 * 1024;
// Tensor-core accumulation into TMEM truncates (measured on B200: relative bias ~ -1.1e-8 per k for 3xTF32 on U[0,1)
// data, i.e. -2.2e-5 at K = 2048).  The parity-gated 3xTF32 path therefore accumulates at most kChunkKb k-blocks
// (256 k) per TMEM chain and the epilogue adds the chunks in f32 round-to-nearest into a running sum kept in TMEM.
constexpr int kChunkKb = 8;
// FS kernels: lo = x - trunc(x) is never negative, so the two cross terms are biased the same way as the main one and the
// truncation bias per chain triples; chains of half the length keep the result where the pre-split form has it.
constexpr int kChunkKbFs = 4;

template <int BN, int PASSES, int CG>
struct TcCfg {
    static constexpr int SETS = PASSES == 3 ? 2 : 1;                  // hi / lo operand copies
    static constexpr int B_ROWS = BN / CG;                            // rows of the B tile held by one CTA
    static constexpr int B_TILE_BYTES = B_ROWS * kRowBytes;
    static constexpr int STAGE_BYTES = SETS * (kATileBytes + B_TILE_BYTES);
    static constexpr int BAR_BYTES = 1024;
    static constexpr int STAGES_RAW = (kSmemLimit - BAR_BYTES - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
    // 3xTF32 keeps a third TMEM region: the running f32 sum of K chunks (see CHUNK_KB)
    static constexpr int TMEM_COLS = PASSES == 3 ? 512 : 2 * BN;      // two accumulator stages (+ running sum); power of 2
    static_assert(PASSES != 3 || 3 * BN <= 512, "3xTF32 needs two accumulator stages and a running sum in 512 TMEM columns");
    static_assert(STAGES >= 2, "pipeline too shallow");
};

// TMA-store epilogue: two staging blocks of [kEpiCols columns][128 rows] of TOut behind the barriers.  The ring gives up a
// stage where the two do not fit beside it (BLOCK_N 256 x CTA pairs: 7 -> 6 stages); the launch asks for the larger of the
// two shared-memory footprints.
template <int BN, int PASSES, int CG, typename TOut>
struct TcEpi {
    using Cfg = TcCfg<BN, PASSES, CG>;
    static constexpr int STG_BYTES = kBlockM * kEpiCols * (int)sizeof(TOut);
    static constexpr int STAGES_RAW = (kSmemLimit - Cfg::BAR_BYTES - 1024 - 2 * STG_BYTES) / Cfg::STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW < Cfg::STAGES ? STAGES_RAW : Cfg::STAGES;
    static constexpr int SMEM_BYTES_EPI = STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + 2 * STG_BYTES + 1024;
    static constexpr int SMEM_BYTES = SMEM_BYTES_EPI > Cfg::SMEM_BYTES ? SMEM_BYTES_EPI : Cfg::SMEM_BYTES;
    static_assert(STAGES >= 2, "pipeline too shallow beside the epilogue staging blocks");
    static_assert(SMEM_BYTES <= kSmemLimit, "shared-memory budget exceeded");
};

// Tile index -> (batch, m-tile, n-tile); groups of kSuperM m-tiles sweep N together.
__device__ __forceinline__ void tile_coords(uint32_t t, const TcArgs &a, uint32_t &bt, uint32_t &mt, uint32_t &nt) {
    const uint32_t per_batch = a.tiles_m * a.tiles_n;
    bt = t / per_batch;
    const uint32_t r = t - bt * per_batch;
    const uint32_t group_span = kSuperM * a.tiles_n;
    const uint32_t g = r / group_span;
    const uint32_t first_m = g * kSuperM;
    const uint32_t gsz = min((uint32_t)kSuperM, a.tiles_m - first_m);
    const uint32_t rr = r - g * group_span;
    mt = first_m + rr % gsz;
    nt = rr / gsz;
}

template <typename T>
__device__ __forceinline__ void store_out(T *p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// KIND: 0 = bf16 (kind::f16), 1 = tf32.  A_MN: operand A is MN-major (the non-transposed product).
// B_MN: operand B is MN-major, i.e. N is its contiguous axis (a row-major m2; bf16 only, f32 operands are re-materialised).
// FS ("fused split", 3xTF32 only): TMA loads the caller's RAW f32 tiles; warps 2 and 3 write lo = x - trunc_tf32(x) beside them
// (same swizzled layout, element for element) and the MMAs use the raw tile as `hi` — the tensor core reads an f32 operand by
// truncating it to TF32 (measured: tools/tf32_input_probe.py), so hi + lo = x exactly.  No operand pre-pass, no dense hi / lo
// copies in HBM.
template <int KIND, bool A_MN, bool B_MN, int BN, int PASSES, typename TOut, int CG, bool FS = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const __grid_constant__ CUtensorMap tmBt, const __grid_constant__ CUtensorMap tmBlot,
               const __grid_constant__ TcDstMaps dmaps, const TcArgs args) {
    using Cfg = TcCfg<BN, PASSES, CG>;
    constexpr int ES = KIND == 0 ? 2 : 4;                 // operand element size
    constexpr int BLOCK_K = kRowBytes / ES;               // 64 bf16 / 32 tf32
    constexpr int UMMA_K = 32 / ES;                       // 16 bf16 / 8 tf32
    constexpr int K_STEPS = BLOCK_K / UMMA_K;             // 4
    constexpr int A_ATOMS = A_MN ? (kBlockM * ES) / kRowBytes : 1;   // MN-major: 128-byte atoms along M (2 bf16 / 4 tf32)
    constexpr int A_ATOM_ELEMS = kRowBytes / ES;
    constexpr int A_ATOM_BYTES = BLOCK_K * kRowBytes;     // one atom column: BLOCK_K rows of 128 B
    constexpr uint32_t IDESC = make_idesc(KIND, A_MN, B_MN, kBlockM * CG, BN);
    static_assert(!B_MN || KIND == 0, "MN-major B is implemented for 16-bit operands only");
    static_assert(!FS || (PASSES == 3 && KIND == 1), "the in-kernel operand split belongs to 3xTF32");
    constexpr bool CHUNKED = PASSES == 3;
    constexpr int CHUNK_KB = FS ? kChunkKbFs : kChunkKb;
    using Epi = TcEpi<BN, PASSES, CG, TOut>;
    constexpr int STG_BYTES = Epi::STG_BYTES;
    const bool tma_epi = args.epi_tma != 0;
    const int STAGES = tma_epi ? Epi::STAGES : Cfg::STAGES;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t *full_bar = bars;                    // [STAGES]
    uint64_t *empty_bar = bars + STAGES;          // [STAGES]
    uint64_t *tmem_full_bar = bars + 2 * STAGES;  // [2]
    uint64_t *tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
    uint64_t *conv_bar = bars + 2 * STAGES + 4;   // [STAGES]  (FS: lo tiles written by the converter warps of both CTAs)
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 4);
    volatile uint32_t *split_flag = tmem_ptr_smem + 1;   // epilogue-warps-only broadcast slot
    uint8_t *staging = reinterpret_cast<uint8_t *>(bars) + Cfg::BAR_BYTES;   // [2][kEpiCols][128] TOut (TMA-store epilogue only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long tr_entry = 0;
    if (args.trace) tr_entry = globaltimer_ns();
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const uint32_t cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmBt);
        if (PASSES == 3 && !FS) {
            prefetch_tmap(&tmAlo);
            prefetch_tmap(&tmBlo);
            prefetch_tmap(&tmBlot);
        }
        if (args.epi_tma == 1)
            for (uint32_t d = 0; d < args.npeers; ++d) prefetch_tmap(&dmaps.m[d]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);    // one arrive.expect_tx (leader producer) + TMA bytes
            mbar_init(smem_u32(empty_bar + s), 1);   // one tcgen05.commit
            if (FS) mbar_init(smem_u32(conv_bar + s), CG);   // one arrive per CTA of the pair: its converter warp for this k-block
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(tmem_full_bar + s), 1);         // one tcgen05.commit
            mbar_init(smem_u32(tmem_empty_bar + s), 4 * CG);   // one arrive per epilogue warp (of both CTAs)
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<CG>(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the
    // tail of the previous kernel on the queue; nothing below may touch global memory before that kernel has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        // The whole warp runs the loop converged (barrier waits by all lanes); one elected lane issues the copies.
        uint32_t stage = 0, phase = 0;
        const uint32_t skip = args.debug_skip;
        for (uint32_t u = cluster_id; u < args.total_units; u += num_clusters) {
            const WorkUnit wu = decode_unit(u, args, BN);
            uint32_t bt, mt, nt;
            tile_coords(wu.tile, args, bt, mt, nt);
            const int m0 = (int)(mt * (kBlockM * CG) + cta_rank * kBlockM);
            const uint32_t b_rows = wu.bn / CG;      // rows of the B tile this CTA loads
            const int n0 = (int)(nt * BN + wu.n_off + cta_rank * b_rows);
            // FS: only the raw tiles are loaded, and each CTA's copies complete on its OWN barrier (its converter warps wait there)
            const uint32_t stage_tx = FS ? (uint32_t)kATileBytes + b_rows * kRowBytes
                                         : (uint32_t)Cfg::SETS * (((skip & 1u) ? 0u : (uint32_t)kATileBytes) +
                                                                  ((skip & 2u) ? 0u : b_rows * kRowBytes)) * CG;
            constexpr bool TMA2 = CG == 2 && !FS;    // cta_group::2 form: completion signalled on the leader's barrier
            const CUtensorMap *tb_hi = wu.narrow ? &tmBt : &tmB, *tb_lo = wu.narrow ? &tmBlot : &tmBlo;
            int k0 = (int)(wu.kb0 * BLOCK_K);
            for (uint32_t kb = wu.kb0; kb < wu.kb1; ++kb, k0 += BLOCK_K) {
                mbar_wait(smem_u32(empty_bar + stage), phase ^ 1, args.mbar_timeout);
                if (elect_one()) {
                    const uint32_t fb = smem_u32(full_bar + stage);
                    if (leader || FS) mbar_arrive_expect_tx(fb, stage_tx);
                    const uint32_t sbase = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
                    for (int set = 0; set < (FS ? 1 : Cfg::SETS); ++set) {
                        const CUtensorMap *ta = set == 0 ? &tmA : &tmAlo;
                        const CUtensorMap *tb = set == 0 ? tb_hi : tb_lo;
                        const uint32_t sa = sbase + set * kATileBytes;
                        const uint32_t sb = sbase + Cfg::SETS * kATileBytes + set * Cfg::B_TILE_BYTES;
                        if (skip & 1u) {
                        } else if (A_MN) {
#pragma unroll
                            for (int at = 0; at < A_ATOMS; ++at) {
                                if (TMA2) tma_load_3d_2sm(sa + at * A_ATOM_BYTES, ta, fb, m0 + at * A_ATOM_ELEMS, k0, (int)bt);
                                else tma_load_3d(sa + at * A_ATOM_BYTES, ta, fb, m0 + at * A_ATOM_ELEMS, k0, (int)bt);
                            }
                        } else {
                            if (TMA2) tma_load_3d_2sm(sa, ta, fb, k0, m0, (int)bt);
                            else tma_load_3d(sa, ta, fb, k0, m0, (int)bt);
                        }
                        if (skip & 2u) {
                        } else if (B_MN) {   // 128-byte atoms along N, each BLOCK_K rows deep
                            for (uint32_t at = 0; at * A_ATOM_ELEMS < b_rows; ++at) {
                                if (TMA2) tma_load_3d_2sm(sb + at * A_ATOM_BYTES, tb, fb, n0 + (int)(at * A_ATOM_ELEMS), k0, (int)bt);
                                else tma_load_3d(sb + at * A_ATOM_BYTES, tb, fb, n0 + (int)(at * A_ATOM_ELEMS), k0, (int)bt);
                            }
                        } else if (TMA2) tma_load_3d_2sm(sb, tb, fb, k0, n0, (int)bt);
                        else tma_load_3d(sb, tb, fb, k0, n0, (int)bt);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================
        // Converged warp, one elected lane issues.  The loop body is kept to a few instructions per MMA (descriptor low words
        // advance by constants, high words are compile-time): measured on B200, the previous single-lane form (vote-and-retry
        // loop around every UTCHMMA, 64-bit descriptor rebuilds, a clock read per wait) spent 666 SM cycles per k-block of four
        // 256x256x16 MMAs whose tensor-pipe floor is 512 — the issue loop, not the pipe or the operand traffic, set the pace.
        if (leader) {
            unsigned long long tr_wait = 0, tr_first = 0, tr_c0 = 0, tr_kb = 0, tr_units = 0;
            bool tr_pending = args.trace != nullptr;
            if (tr_pending) tr_wait = globaltimer_ns();
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            // MN-major 32-bit operands use 32-byte swizzle atoms: 4 K-rows (512 B) per atom
            constexpr bool A32 = A_MN && KIND == 1;
            constexpr uint32_t A_HI = smem_desc_hi(A32 ? 512 : 1024, A32 ? 1 : 2), B_HI = smem_desc_hi(1024, 2);
            constexpr uint32_t A_STEP = (A_MN ? UMMA_K * kRowBytes : 32) >> 4;   // K advance per MMA, 16-byte units
            constexpr uint32_t B_STEP = (B_MN ? UMMA_K * kRowBytes : 32) >> 4;
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), A_MN ? A_ATOM_BYTES : 16);
            const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem) + Cfg::SETS * kATileBytes, B_MN ? A_ATOM_BYTES : 16);
            for (uint32_t u = cluster_id; u < args.total_units; u += num_clusters) {
                const WorkUnit wu = decode_unit(u, args, BN);
                const uint32_t idesc = wu.narrow ? args.idesc_tail : IDESC;
                uint32_t kb = wu.kb0;
                while (kb < wu.kb1) {
                    // one TMEM accumulation chain: the whole K range, or kChunkKb k-blocks for 3xTF32
                    const uint32_t chain_end = CHUNKED ? min(wu.kb1, kb + (uint32_t)CHUNK_KB) : wu.kb1;
                    mbar_wait(smem_u32(tmem_empty_bar + acc), acc_phase ^ 1, args.mbar_timeout);   // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    uint32_t accum = 0;   // the first MMA of a chain overwrites the accumulator
                    for (; kb < chain_end; ++kb) {
                        if (FS) mbar_wait_cluster(smem_u32(conv_bar + stage), phase, args.mbar_timeout);   // raw tiles landed AND lo tiles written
                        else mbar_wait(smem_u32(full_bar + stage), phase, args.mbar_timeout);
                        tc_fence_after();
                        if (tr_pending) {
                            tr_pending = false;
                            tr_first = globaltimer_ns();
                            tr_c0 = clock64();
                        }
                        if (elect_one()) {
                            const uint32_t a_lo = a_lo0 + stage * (Cfg::STAGE_BYTES >> 4);
                            const uint32_t b_lo = b_lo0 + stage * (Cfg::STAGE_BYTES >> 4);
#pragma unroll
                            for (int j = 0; j < K_STEPS; ++j) {
                                const uint64_t da_hi = pack_desc(a_lo + j * A_STEP, A_HI);
                                const uint64_t db_hi = pack_desc(b_lo + j * B_STEP, B_HI);
                                if (PASSES == 3) {
                                    const uint64_t da_lo = pack_desc(a_lo + (kATileBytes >> 4) + j * A_STEP, A_HI);
                                    const uint64_t db_lo = pack_desc(b_lo + (Cfg::B_TILE_BYTES >> 4) + j * B_STEP, B_HI);
                                    umma<KIND, CG>(d_tmem, da_lo, db_hi, idesc, j == 0 ? accum : 1u);   // small terms first
                                    umma<KIND, CG>(d_tmem, da_hi, db_lo, idesc, 1u);
                                    umma<KIND, CG>(d_tmem, da_hi, db_hi, idesc, 1u);
                                } else {
                                    umma<KIND, CG>(d_tmem, da_hi, db_hi, idesc, j == 0 ? accum : 1u);
                                }
                            }
                            umma_commit<CG>(smem_u32(empty_bar + stage));            // smem slot free once these MMAs retire
                            if (kb + 1 == chain_end) umma_commit<CG>(smem_u32(tmem_full_bar + acc));   // accumulator ready
                        }
                        __syncwarp();
                        accum = 1u;
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
                tr_kb += wu.kb1 - wu.kb0;
                ++tr_units;
            }
            if (args.trace && cluster_id < 256 && lane == 0) {
                unsigned long long *t = args.trace + (size_t)cluster_id * 8;
                t[0] = tr_entry; t[1] = tr_wait; t[2] = tr_first; t[3] = globaltimer_ns();
                t[4] = (unsigned long long)clock64() - tr_c0; t[5] = tr_kb; t[7] = tr_units;
            }
        }
    } else if (FS && (warp == 2 || warp == 3)) {
        // ===================================== operand split (FS): lo = x - trunc_tf32(x) ===========
        // Warp 2 takes the even k-blocks, warp 3 the odd ones, so two stages are being split at any time and the shared-memory
        // latencies of one overlap the arithmetic of the other.  A warp walks the raw A and B tiles of its stage 16 bytes per lane,
        // eight chunks in flight, and writes the lo tiles at the same offsets (the swizzle is a permutation of 16-byte chunks
        // inside a tile, identical for raw and lo).  lo is left unrounded: the tensor core truncates it like every f32 operand,
        // an error of 2^-11 of lo = 2^-21 of x at most.
        uint32_t stage = 0, phase = 0, count = 0;
        const uint32_t mine = warp - 2;
        auto split8 = [](uint32_t raw_addr, uint32_t lo_addr, uint32_t n, uint32_t stride) {   // n <= 8 chunks, `stride` bytes apart
            uint32_t x[8][4];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((uint32_t)q < n)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[q][0]), "=r"(x[q][1]), "=r"(x[q][2]), "=r"(x[q][3])
                                 : "r"(raw_addr + q * stride) : "memory");
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((uint32_t)q < n) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[q][e] = __float_as_uint(__uint_as_float(x[q][e]) - __uint_as_float(x[q][e] & 0xFFFFE000u));
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr + q * stride), "r"(x[q][0]), "r"(x[q][1]), "r"(x[q][2]),
                                 "r"(x[q][3]) : "memory");
                }
        };
        for (uint32_t u = cluster_id; u < args.total_units; u += num_clusters) {
            const WorkUnit wu = decode_unit(u, args, BN);
            const uint32_t b_chunks = (wu.bn / CG) * (kRowBytes / 16);      // 16-byte chunks of this CTA's B tile
            for (uint32_t kb = wu.kb0; kb < wu.kb1; ++kb, ++count) {
                if ((count & 1u) == mine) {
                    mbar_wait(smem_u32(full_bar + stage), phase, args.mbar_timeout);
                    const uint32_t sbase = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t a_raw = sbase + lane * 16, b_raw = sbase + 2 * kATileBytes + lane * 16;
                    // A: 1024 chunks = 4 rounds of 8 x 32 lanes; B: b_chunks (a multiple of 32) in rounds of up to 8 x 32
#pragma unroll
                    for (uint32_t r = 0; r < (uint32_t)kATileBytes / (16 * 32 * 8); ++r)
                        split8(a_raw + r * 4096, a_raw + kATileBytes + r * 4096, 8, 512);
                    for (uint32_t c0 = 0; c0 < b_chunks; c0 += 256) {
                        const uint32_t left = (b_chunks - c0) / 32;
                        split8(b_raw + c0 * 16, b_raw + Cfg::B_TILE_BYTES + c0 * 16, left < 8 ? left : 8, 512);
                    }
                    fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's (async proxy) reads
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(smem_u32(conv_bar + stage), 0);
                        else mbar_arrive(smem_u32(conv_bar + stage));
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================================== epilogue: TMEM -> registers -> global ==============
        const uint32_t q = warp & 3;   // TMEM lane quarter this warp may access
        uint32_t acc = 0, acc_phase = 0;
        if (args.handshake) {
            // fused all-gather: do not write into a peer before it has reached this step on its own queue
            if (threadIdx.x == 128)
                for (uint32_t r = 0; r < args.npeers; ++r)
                    if (r != args.my_rank) {
                        const long long t0 = clock64();
                        while ((int)(ld_acquire_sys(args.ready_local + r) - args.epoch) < 0)
                            if (args.peer_timeout > 0 && clock64() - t0 > args.peer_timeout) __trap();
                    }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const uint32_t lane_base = tmem_base + ((q * 32u) << 16);
        uint32_t staged = 0;   // blocks staged so far by this CTA (selects the staging buffer; uniform over the 128 threads)
        for (uint32_t u = cluster_id; u < args.total_units; u += num_clusters) {
            const WorkUnit wu = decode_unit(u, args, BN);
            uint32_t bt, mt, nt;
            tile_coords(wu.tile, args, bt, mt, nt);
            const uint32_t row_in_cta = q * 32 + lane;
            const uint32_t row = mt * (kBlockM * CG) + cta_rank * kBlockM + row_in_cta;
            const uint32_t n0 = nt * BN + wu.n_off;
            const int nblk = (int)(wu.bn / 32);      // 32-column blocks in this unit's accumulator
            const bool split_unit = wu.split_idx != 0xFFFFFFFFu;
            const uint64_t crow_off = (uint64_t)bt * args.sc + row;   // element offset of (row, col 0) inside a panel
            const bool row_ok = row < args.M;
            // split units park their f32 partial in the workspace: [slot][split][rank][col][row]
            const uint32_t slot = wu.tile - args.full_tiles;
            float *wsp = split_unit ? args.ws + (((uint64_t)slot * args.split + wu.split_idx) * CG + cta_rank) * (uint64_t)(BN * kBlockM) + row_in_cta
                                    : nullptr;
            const uint32_t nkb = wu.kb1 - wu.kb0;
            const uint32_t nchains = CHUNKED ? (nkb + CHUNK_KB - 1) / CHUNK_KB : 1u;
            for (uint32_t ch = 0; ch < nchains; ++ch) {
                const bool final_chain = ch + 1 == nchains;
                mbar_wait(smem_u32(tmem_full_bar + acc), acc_phase, args.mbar_timeout);
                tc_fence_after();
                const uint32_t taddr = lane_base + acc * BN;
                const uint32_t tsum = lane_base + 2 * BN;   // running sum of the chains (3xTF32 only)
                // The 32-column blocks are processed in pairs by a loop that is NOT unrolled: the two register blocks keep static
                // names (cur / nxt swap roles inside the pair) while the code stays a quarter of the fully unrolled form — the
                // epilogue runs once per tile, so its instruction footprint is fetched cold every time.
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr, v0);
                auto block = [&](const int c, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
                    uint32_t sprev[32];
                    if (CHUNKED && ch > 0) tmem_ld32(tsum + c * 32, sprev);
                    tmem_ld_wait();
                    if (c + 1 < nblk) tmem_ld32(taddr + (c + 1) * 32, nxt);
                    else {
                        // all TMEM reads of this accumulator are done: hand it back to the MMA issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_cluster(smem_u32(tmem_empty_bar + acc), 0);
                            else mbar_arrive(smem_u32(tmem_empty_bar + acc));
                        }
                    }
                    if (CHUNKED && ch > 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) cur[i] = __float_as_uint(__uint_as_float(cur[i]) + __uint_as_float(sprev[i]));
                    }
                    if (CHUNKED && !final_chain) {
                        tmem_st32(tsum + c * 32, cur);
                    } else if (split_unit) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) __stcg(wsp + (uint64_t)(c * 32 + i) * kBlockM, __uint_as_float(cur[i]));
                    } else if (args.red_axis != 0) {
                        // fused reduction: neutral element outside the view, SqNorm squares first
                        const int rop = args.red_op;
                        const float init = rop == WGB_RED_MIN ? 3.4e38f : rop == WGB_RED_MAX ? -3.4e38f : rop == WGB_RED_PROD ? 1.0f : 0.0f;
                        auto comb = [rop](float a, float b) {
                            return rop == WGB_RED_MIN ? fminf(a, b) : rop == WGB_RED_MAX ? fmaxf(a, b) : rop == WGB_RED_PROD ? a * b : a + b;
                        };
                        float x[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float val = __uint_as_float(cur[i]);
                            x[i] = (row_ok && n0 + c * 32 + i < args.N) ? (rop == WGB_RED_SQNORM ? val * val : val) : init;
                        }
                        if (args.red_axis == 2) {
                            // one value per row: fold the 32 columns of the block in column order
                            float a = x[0];
#pragma unroll
                            for (int i = 1; i < 32; ++i) a = comb(a, x[i]);
                            if (row_ok && n0 + c * 32 < args.N) args.red_partials[(uint64_t)((n0 + c * 32) / 32) * args.M + row] = a;
                        } else {
                            // one value per column: transpose-reduce across the warp's 32 rows (31 shuffles): after the step with
                            // offset s, the lanes with bit s set keep the upper half of the remaining columns
#pragma unroll
                            for (int s = 16; s > 0; s >>= 1) {
                                const bool upper = (lane & s) != 0;
#pragma unroll
                                for (int j = 0; j < s; ++j) {
                                    const float mine = upper ? x[j + s] : x[j], send = upper ? x[j] : x[j + s];
                                    x[j] = comb(mine, __shfl_xor_sync(0xFFFFFFFFu, send, s));
                                }
                            }
                            const uint32_t col = n0 + c * 32 + lane;      // lane l now holds column l of the block
                            if (col < args.N && row - lane < args.M) args.red_partials[(uint64_t)((row - lane) / 32) * args.N + col] = x[0];
                        }
                    } else if (tma_epi) {
                        // TMEM -> registers -> shared memory block [32 columns][128 rows] -> one bulk tensor store per
                        // destination.  Whole 256 B (bf16) / 512 B (f32) column segments leave the SM instead of per-lane
                        // 2 / 4-byte stores, and the epilogue warps issue 32 st.shared per block whatever the number of peers.
                        if (args.ep_op >= 0 && row_ok) {
                            const TOut *erow = reinterpret_cast<const TOut *>(args.ep) + (uint64_t)bt * args.ep_sm + row;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const uint32_t col = n0 + c * 32 + i;
                                if (col < args.N)
                                    cur[i] = __float_as_uint(epilogue_apply<TOut>(args.ep_op, __uint_as_float(cur[i]), erow + (uint64_t)col * args.ep_ld));
                            }
                        }
                        uint8_t *blk = staging + (staged & 1u) * STG_BYTES;
                        ++staged;
                        // the stores issued from this block two blocks ago have read it (the newest group may still be reading
                        // the other block)
                        if (threadIdx.x == 128) bulk_wait_read<1>();
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        TOut *sp = reinterpret_cast<TOut *>(blk) + row_in_cta;
#pragma unroll
                        for (int i = 0; i < 32; ++i) store_out<TOut>(sp + i * kBlockM, __uint_as_float(cur[i]));
                        fence_proxy_async_smem();
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (args.epi_tma == 2) {
                            // multicast: every thread re-reads 16-byte pieces of the staged block (a column of the block is
                            // 128 rows = 16 (bf16) / 32 (f32) pieces) and stores each ONCE; the switch replicates it
                            constexpr int PPC = kBlockM * (int)sizeof(TOut) / 16;          // pieces per column
                            constexpr int EPP = 16 / (int)sizeof(TOut);                    // elements per piece
                            const uint32_t r0 = mt * (kBlockM * CG) + cta_rank * kBlockM, c0 = n0 + c * 32;
                            char *mbase = args.mc_dst + ((uint64_t)bt * args.sc + r0) * sizeof(TOut);
#pragma unroll
                            for (int j = 0; j < kEpiCols * PPC / 128; ++j) {
                                const uint32_t p = (threadIdx.x - 128) + 128 * j;
                                const uint32_t col = p / PPC, piece = p % PPC;
                                if (c0 + col < args.N && r0 + piece * EPP < args.M)
                                    multimem_st16(mbase + ((uint64_t)(c0 + col) * args.ldc) * sizeof(TOut) + piece * 16,
                                                  *reinterpret_cast<const uint4 *>(blk + (col * PPC + piece) * 16));
                            }
                        } else if (threadIdx.x == 128) {
                            const int r0 = (int)(mt * (kBlockM * CG) + cta_rank * kBlockM), c0 = (int)(n0 + c * 32);
                            uint32_t d = args.first_dst;
                            for (uint32_t i = 0; i < args.npeers; ++i) {
                                if ((args.store_mask >> d) & 1u) tma_store_3d(&dmaps.m[d], smem_u32(blk), r0, c0, (int)bt);
                                if (++d == args.npeers) d = 0;
                            }
                            bulk_commit();
                        }
                        __syncwarp();
                    } else if (row_ok) {
                        if (args.ep_op >= 0) {   // fused OpAssign: out = acc (op) e, e read coalesced like the store
                            const TOut *erow = reinterpret_cast<const TOut *>(args.ep) + (uint64_t)bt * args.ep_sm + row;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const uint32_t col = n0 + c * 32 + i;
                                if (col < args.N)
                                    cur[i] = __float_as_uint(epilogue_apply<TOut>(args.ep_op, __uint_as_float(cur[i]), erow + (uint64_t)col * args.ep_ld));
                            }
                        }
                        uint32_t d = args.first_dst;
                        for (uint32_t k = 0; k < args.npeers; ++k) {   // npeers == 1 unless the all-gather is fused in
                            if ((args.store_mask >> d) & 1u) {
                                TOut *crow = reinterpret_cast<TOut *>(args.dst[d]) + crow_off;
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const uint32_t col = n0 + c * 32 + i;
                                    if (col < args.N) store_out<TOut>(crow + (uint64_t)col * args.ldc, __uint_as_float(cur[i]));
                                }
                            }
                            if (++d == args.npeers) d = 0;
                        }
                    }
                };
#pragma unroll 1
                for (int c = 0; c < nblk; c += 2) {
                    block(c, v0, v1);
                    if (c + 1 < nblk) block(c + 1, v1, v0);
                }
                if (CHUNKED && !final_chain) tmem_st_wait();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            if (split_unit) {
                // last CTA to park its partial folds all of them in split order (deterministic) and writes the tile
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 128) *split_flag = atomicAdd(args.counters + slot * CG + cta_rank, 1u);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const bool last = *split_flag == args.split - 1;
                asm volatile("bar.sync 1, 128;" ::: "memory");   // everyone has read the flag before it can be rewritten
                if (last) {
                    __threadfence();
                    const float *wbase = args.ws + ((uint64_t)slot * args.split * CG + cta_rank) * (uint64_t)(BN * kBlockM) + row_in_cta;
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        float sum[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[i] = 0.f;
                        for (uint32_t sp = 0; sp < args.split; ++sp) {
                            const float *src = wbase + (uint64_t)sp * CG * (BN * kBlockM) + (uint64_t)(c * 32) * kBlockM;
#pragma unroll
                            for (int i = 0; i < 32; ++i) sum[i] += __ldcg(src + (uint64_t)i * kBlockM);   // 32 loads in flight
                        }
                        if (row_ok) {
                            if (args.ep_op >= 0) {
                                const TOut *erow = reinterpret_cast<const TOut *>(args.ep) + (uint64_t)bt * args.ep_sm + row;
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const uint32_t col = n0 + c * 32 + i;
                                    if (col < args.N) sum[i] = epilogue_apply<TOut>(args.ep_op, sum[i], erow + (uint64_t)col * args.ep_ld);
                                }
                            }
                            for (uint32_t d = 0; d < args.npeers; ++d) {
                                TOut *crow = reinterpret_cast<TOut *>(args.dst[d]) + crow_off;
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const uint32_t col = n0 + c * 32 + i;
                                    if (col < args.N) store_out<TOut>(crow + (uint64_t)col * args.ldc, sum[i]);
                                }
                            }
                        }
                    }
                    if (threadIdx.x == 128) args.counters[slot * CG + cta_rank] = 0u;
                }
            }
        }
        if (args.trace && leader && threadIdx.x == 128 && cluster_id < 256) args.trace[(size_t)cluster_id * 8 + 6] = globaltimer_ns();
        if (tma_epi && threadIdx.x == 128) {
            // every bulk store of this CTA has been performed (not merely read out of shared memory) before the CTA may exit
            // or count itself as done; order the async-proxy writes before the generic-proxy flag traffic below
            bulk_wait_all();
            fence_proxy_async_all();
        }
        if (args.handshake) {
            // all of this CTA's peer stores are out: make them visible system-wide, count the CTA, and let the last CTA of
            // the grid publish "rank my_rank's panel is complete" to every peer
            __threadfence_system();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 128) {
                __threadfence();
                const unsigned int prev = atomicAdd(args.cta_counter, 1u);
                if (prev == gridDim.x - 1) {
                    __threadfence_system();
                    for (uint32_t r = 0; r < args.npeers; ++r)
                        if (r != args.my_rank) st_release_sys(args.done_remote[r] + args.my_rank, args.epoch);
                    *args.cta_counter = 0u;
                }
            }
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
    }
}

inline int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

struct TcMaps {
    CUtensorMap a, alo, b, blo, bt, blot;   // operands (hi / lo), and B with the narrow box of the tail strips
    TcDstMaps dst;                          // outputs (TMA-store epilogue only)
};

template <int KIND, bool A_MN, bool B_MN, int BN, int PASSES, typename TOut, int CG, bool FS = false>
wgb_status launch_cfg(wgb_pass *p, const TcMaps &m, TcArgs args) {
    constexpr int SMEM_BYTES = TcEpi<BN, PASSES, CG, TOut>::SMEM_BYTES;
    auto kern = gemm_tc_kernel<KIND, A_MN, B_MN, BN, PASSES, TOut, CG, FS>;
    static bool attr_set[64] = {};   // per instantiation, per device
    const int dev = p->ctx->device & 63;
    if (!attr_set[dev]) {
        WGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[dev] = true;
    }
    uint32_t sms = (uint32_t)p->ctx->prop.multiProcessorCount;
    const uint32_t margin = (uint32_t)comm_sm_margin(p->ctx);
    if (margin < sms / 2) sms -= margin;
    uint32_t clusters = sms / CG;
    if (clusters > args.total_units) clusters = args.total_units;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * CG);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = p->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env_int("WGB_TC_PDL", 1) != 0 ? 2 : 1;
    WGB_CUDA(cudaLaunchKernelEx(&cfg, kern, m.a, m.alo, m.b, m.blo, m.bt, m.blot, m.dst, args));
    count_launch(p->ctx);
    return WGB_OK;
}

template <int KIND, bool A_MN, bool B_MN, int PASSES, typename TOut>
wgb_status launch_sel(wgb_pass *p, int bn, int cg, const TcMaps &m, const TcArgs &args) {
    if constexpr (PASSES == 3) {   // 3xTF32: BLOCK_N = 128 only (two accumulator stages + running sum in TMEM)
        if (args.fused_split) {
            if (cg == 2) return launch_cfg<KIND, A_MN, B_MN, 128, PASSES, TOut, 2, true>(p, m, args);
            return launch_cfg<KIND, A_MN, B_MN, 128, PASSES, TOut, 1, true>(p, m, args);
        }
        if (cg == 2) return launch_cfg<KIND, A_MN, B_MN, 128, PASSES, TOut, 2>(p, m, args);
        return launch_cfg<KIND, A_MN, B_MN, 128, PASSES, TOut, 1>(p, m, args);
    } else {
        if (cg == 2) {
            if (bn == 256) return launch_cfg<KIND, A_MN, B_MN, 256, PASSES, TOut, 2>(p, m, args);
            return launch_cfg<KIND, A_MN, B_MN, 128, PASSES, TOut, 2>(p, m, args);
        }
        if (bn == 256) return launch_cfg<KIND, A_MN, B_MN, 256, PASSES, TOut, 1>(p, m, args);
        return launch_cfg<KIND, A_MN, B_MN, 128, PASSES, TOut, 1>(p, m, args);
    }
}

}  // namespace tc
}  // namespace wgb
