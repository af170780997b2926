// gemm_tc.cu — the tensor-core GEMM: tcgen05.mma with TMEM accumulators, TMA-staged operand tiles,
// warp-specialised producer / MMA-issuer / epilogue, persistent tile loop.  sm_100a only.
//
// Replaces the four WGSL GEMM kernels of /root/reference/crates/wgebra/src/linalg/gemm.wgsl (:29-199),
// which are global-memory bound with zero operand reuse.  Same contract as
// gemm.rs:65-127: out = m1 * m2 or tr(m1) * m2, column-major views, overwrite, batched over size[2].
//
// Data layout
//   column-major m1 [M x K] (non-tr)  -> UMMA operand A, "MN-major": M is the contiguous axis.
//   column-major m1 [K x M] (tr)      -> UMMA operand A, K-major.  No data transpose: only the major bit.
//   column-major m2 [K x N]           -> UMMA operand B, K-major (K contiguous).
//   out [M x N] column-major          <- accumulator rows = TMEM lanes, so a warp's 32 lanes store 32
//                                        consecutive elements of one output column (coalesced).
//   Shared-memory tiles use the 128-byte swizzle written by TMA and read by the UMMA descriptors:
//     K-major tile  : [rows][128 B of K]            SBO = 1024 B, K advance = 32 B per MMA
//     MN-major tile : [M atom][BLOCK_K][128 B of M] LBO = BLOCK_K*128 B, SBO = 1024 B, K advance = UMMA_K*128 B
//   BLOCK_K = 128 B of K per stage (64 bf16 / 32 tf32), BLOCK_M = 128 per CTA, BLOCK_N = 128 or 256.
//
// Kernel shape (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
// allocator, warps 4-7 = epilogue (tcgen05.ld -> registers -> global).  Three pipelines: smem full/empty
// (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two accumulator stages so the epilogue of tile i
// overlaps the main loop of tile i+1), and the persistent tile loop (grid = #SMs, supertile rasterisation
// so concurrently running tiles share operand panels in L2).
// CG = 2 pairs two CTAs (cta_group::2): UMMA M = 256 across the pair, each CTA loads half of the B tile.
//
// f32 operands: kind::tf32.  PASSES = 3 is the error-compensated 3xTF32 scheme (hi/lo operand split made
// by split_tf32_kernel; three MMAs per k-step: lo*hi + hi*lo + hi*hi, f32 accumulation in TMEM), which is
// what meets the 1e-5 parity bound; PASSES = 1 is plain TF32 (informational fast mode).
//
// Roofline: tensor pipe.  Algorithmic work = 2*M*N*K flop per launch (x3 MMA issue for 3xTF32).
#include "gemm_tc_kernel.cuh"

namespace wgb {

namespace tc {
// instantiated in gemm_tc_inst_*.cu
extern template wgb_status launch_sel<0, false, true, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, true, true, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, false, false, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, true, false, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, false, true, 1, __nv_bfloat16>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, true, true, 1, __nv_bfloat16>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, false, false, 1, __nv_bfloat16>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<0, true, false, 1, __nv_bfloat16>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<1, false, false, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<1, true, false, 1, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<1, false, false, 3, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
extern template wgb_status launch_sel<1, true, false, 3, float>(wgb_pass *, int, int, const TcMaps &, const TcArgs &);
}  // namespace tc

using namespace tc;

namespace {

// ---------------------------------------------------------------------------------------------
// 3xTF32 operand split: x = hi + lo, hi = tf32(x) (round-to-nearest), lo = x - hi (exact in f32; the tensor
// core truncates lo to its leading 11 bits, leaving a relative error of ~2^-22 per operand).
// Reads any strided view, writes two dense column-major copies with leading dimension ld_out.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ src, uint32_t rows, uint32_t cols,
                                                         uint32_t mats, uint64_t ld, uint64_t smat, float *__restrict__ hi,
                                                         float *__restrict__ lo, uint64_t ld_out, uint64_t smat_out) {
    const uint64_t total = (uint64_t)ld_out * cols * mats;   // padded rows are written as zeros
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = e % ld_out;
        const uint64_t jt = e / ld_out;
        const uint64_t j = jt % cols, t = jt / cols;
        float x = 0.f;
        if (i < rows) x = __ldg(src + t * smat + j * ld + i);
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
        const float hf = __uint_as_float(h);
        const uint64_t o = t * smat_out + j * ld_out + i;
        hi[o] = hf;
        lo[o] = x - hf;
    }
}

// Same split, but the output is transposed: src is [rows x cols] column-major, hi/lo are [cols x rows] column-major
// (leading dimension ld_out >= cols).  Used to hand a non-transposed f32 m1 to the K-major kernel variant.
// lo may be null (plain tf32-rounded transposed copy).  32x32 tiles through shared memory, coalesced both ways.
__global__ void __launch_bounds__(256) split_tf32_transpose_kernel(const float *__restrict__ src, uint32_t rows, uint32_t cols,
                                                                   uint64_t ld, uint64_t smat, float *__restrict__ hi,
                                                                   float *__restrict__ lo, uint64_t ld_out, uint64_t smat_out) {
    __shared__ float tile[32][33];
    const uint32_t t = blockIdx.z;
    const uint32_t r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float *sp = src + (uint64_t)t * smat;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t r = r0 + tx, c = c0 + ty + 8 * j;
        tile[ty + 8 * j][tx] = (r < rows && c < cols) ? __ldg(sp + (uint64_t)c * ld + r) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t r = r0 + ty + 8 * j, c = c0 + tx;   // output element (c, r): c is the contiguous axis
        if (r < rows && c < cols) {
            const float x = tile[tx][ty + 8 * j];
            uint32_t h;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
            const float hf = __uint_as_float(h);
            const uint64_t o = (uint64_t)t * smat_out + (uint64_t)r * ld_out + c;
            hi[o] = hf;
            if (lo) lo[o] = x - hf;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct TmapKey {
    const void *ptr;
    uint64_t d0, d1, d2, s1, s2;
    uint32_t b0, b1, es;
    bool operator==(const TmapKey &o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b0 == o.b0 && b1 == o.b1 &&
               es == o.es;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey &k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
        for (uint64_t v : {k.d0, k.d1, k.d2, k.s1, k.s2, (uint64_t)k.b0 << 32 | k.b1, (uint64_t)k.es})
            h = (h ^ v) * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
        return (size_t)h;
    }
};
using TmapCache = std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>;

// 3-D map: dim0 contiguous (d0 elements), dim1 stride s1 elements, dim2 (batch) stride s2 elements.
wgb_status get_tmap(wgb_ctx *ctx, const void *ptr, uint32_t es, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                    uint32_t b0, uint32_t b1, CUtensorMap *out, bool atom32 = false, bool no_swizzle = false) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) WGB_FAIL(WGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if (d2 <= 1) {   // single matrix: the batch stride is never used, but must still be a legal (16-byte multiple) stride
        const uint64_t q = 16 / es;
        d2 = 1;
        s2 = d0 > s1 * d1 ? d0 : s1 * d1;
        s2 = (s2 + q - 1) / q * q;
    }
    TmapKey key{ptr, d0, d1, d2, s1, s2, b0, b1, es | (atom32 ? 0x100u : 0u) | (no_swizzle ? 0x200u : 0u)};
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->tmap_cache) ctx->tmap_cache = new TmapCache();
    TmapCache &cache = *static_cast<TmapCache *>(ctx->tmap_cache);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return WGB_OK;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1 * es, s2 * es};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(ptr),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE : atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        WGB_FAIL(WGB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu, strides %llu / %llu B, box %u x %u", (int)r,
                 (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)(s1 * es),
                 (unsigned long long)(s2 * es), b0, b1);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return WGB_OK;
}

}  // namespace

void tmap_cache_destroy(wgb_ctx *ctx) {
    if (ctx->tmap_cache) delete static_cast<TmapCache *>(ctx->tmap_cache);
    ctx->tmap_cache = nullptr;
}

static bool aligned_operand(const void *base, uint64_t off, uint64_t ld, uint64_t smat, uint32_t nmats, size_t es) {
    const uint64_t q = 16 / es;
    return (((uintptr_t)base + off * es) & 15u) == 0 && ld % q == 0 && (nmats <= 1 || smat % q == 0);
}

// single-pass TF32 reads the caller's f32 views in place: both need TMA alignment (3xTF32 re-materialises them dense)
bool gemm_tc_direct_f32_ok(const GemmProblem &g) {
    return aligned_operand(g.a, g.a_off, g.lda, g.sa, g.nmats, 4) && aligned_operand(g.b, g.b_off, g.ldb, g.sb, g.nmats, 4);
}

bool gemm_tc_eligible(const GemmProblem &g) {
    if (g.M == 0 || g.N == 0 || g.K == 0) return false;
    if (g.in_dtype == WGB_F32) return g.out_dtype == WGB_F32;   // operands are re-materialised dense by the split (any view is fine)
    const size_t es = 2;
    return aligned_operand(g.a, g.a_off, g.lda, g.sa, g.nmats, es) && aligned_operand(g.b, g.b_off, g.ldb, g.sb, g.nmats, es);
}

// Picks BLOCK_N.  BLOCK_N = 128 moves 1.5x the operand bytes per flop and measured ~0.64x the throughput of 256
// (probe: 840 vs 1310 TFLOP/s at 4096^3), so it only wins when 256 would leave most of the machine idle; the tail
// wave of 256-wide tiles is handled by split-K instead (plan_split).
static int pick_bn(uint32_t M, uint32_t N, uint32_t nmats, int cg, uint32_t sms) {
    const int forced = env_int("WGB_TC_BN", 0);
    if (forced == 128 || forced == 256) return forced;
    if (N <= 128) return 128;
    return 256;
}

// Tail plan.  With T tiles on C clusters the last T mod C tiles would run alone while the other clusters idle.
// Default: cut each tail tile into 2 or 4 column strips (narrower complete tiles, no fix-up) when that still fits one wave:
// a 128- / 64-column strip costs ~0.75 / ~0.63 of a 256-column tile (the A tile is loaded either way).
// WGB_TC_SPLITK=1 selects the K split with a workspace fix-up instead (measured slower, kept for reference).
static void plan_tail(TcArgs &a, uint32_t clusters, int bn, int cg, int kind, bool a_mn, bool b_mn) {
    a.full_tiles = a.total_tiles;
    a.split = 1;
    a.kb_per_split = a.num_kb;
    a.nsplit = 1;
    a.tail_bn = (uint32_t)bn;
    a.idesc_tail = 0;
    a.total_units = a.total_tiles;
    const uint32_t tail = a.total_tiles % clusters;
    if (tail == 0) return;
    if (env_int("WGB_TC_SPLITK", 0) != 0 && a.red_axis == 0) {   // (split partials are not final values: no fused reduction there)
        uint32_t split = clusters / tail;
        if (split > 8) split = 8;
        if (split > a.num_kb / 4) split = a.num_kb / 4;
        if (split < 2) return;
        const uint32_t per = (a.num_kb + split - 1) / split;
        split = (a.num_kb + per - 1) / per;
        if (split < 2) return;
        a.full_tiles = a.total_tiles - tail;
        a.split = split;
        a.kb_per_split = per;
        a.total_units = a.full_tiles + tail * split;
        return;
    }
    if (env_int("WGB_TC_NSPLIT", 1) == 0) return;
    uint32_t nsplit = 1;
    for (uint32_t cand : {4u, 2u}) {
        const uint32_t w = (uint32_t)bn / cand;
        if ((uint64_t)tail * cand <= clusters && w >= 32 && (w / cg) % (b_mn ? 64 : 8) == 0) {   // MN-major B: whole 128 B atoms
            nsplit = cand;
            break;
        }
    }
    if (nsplit == 1) return;
    a.full_tiles = a.total_tiles - tail;
    a.nsplit = nsplit;
    a.tail_bn = (uint32_t)bn / nsplit;
    a.idesc_tail = make_idesc(kind, a_mn, b_mn, kBlockM * cg, (int)a.tail_bn);
    a.total_units = a.full_tiles + tail * nsplit;
}

wgb_status launch_gemm_tc(wgb_pass *p, const GemmProblem &g, wgb_f32_mode mode, int *path_out) {
    wgb_ctx *ctx = p->ctx;
    const bool f32 = g.in_dtype == WGB_F32;
    int passes = 1;
    if (f32) passes = (mode == WGB_F32_TF32) ? 1 : 3;
    // A non-transposed f32 m1 is an MN-major 32-bit operand.  "direct": TMA 32-byte-atom swizzle + BASE32B descriptor.
    // "transpose": the prep kernel writes tr(m1) dense and the K-major kernel variant runs (always valid).
    const bool mn_direct = env_int("WGB_TF32_MN_DIRECT", 1) != 0;   // validated on B200: direct is the default
    const bool transpose_a = f32 && !g.tr && !mn_direct;
    // An N-contiguous (row-major) f32 m2 is always re-materialised K-major by the transposing prep; a bf16 one is read in
    // place as an MN-major B operand.
    const bool transpose_b = f32 && g.b_nmajor;
    // 3xTF32 with the operand split inside the GEMM (FS kernels): the raw f32 views are TMA-loaded as they are, so both must meet
    // TMA's alignment rules and need no transposing prep; anything else is re-materialised as dense hi / lo copies by the
    // split kernels first (WGB_TF32_FUSED_SPLIT=0 forces that form)
    // Which one: both forms are shared-memory-bandwidth bound at BLOCK_N = 128 (per k-block the three MMA passes read 96 KB, TMA
    // writes 48 KB of hi / lo or 24 KB of raw tiles, and the in-kernel split adds 24 KB read + 24 KB written through the LSU,
    // which measured 1.5x slower per k-block), so the in-kernel split wins only while the two split launches (~7 us) cost more than
    // that: measured 13.3 vs 17.3 us at 256^3, 19.3 vs 21.1 at 512^3, 35.4 vs 35.9 at 1024^3, 132 vs 115 at 2048^3
    // (profiles/README.md).  WGB_TF32_FUSED_SPLIT = 0 / 1 forces one form, unset = by size.
    const int fs_env = env_int("WGB_TF32_FUSED_SPLIT", -1);
    const bool fs_wanted = fs_env >= 0 ? fs_env != 0 : (uint64_t)g.M * g.N * g.K <= ((uint64_t)1 << 30);
    const bool fused_split = f32 && passes == 3 && !transpose_a && !transpose_b && fs_wanted &&
                             aligned_operand(g.a, g.a_off, g.lda, g.sa, g.nmats, 4) && aligned_operand(g.b, g.b_off, g.ldb, g.sb, g.nmats, 4);
    const bool prep_a = f32 && !fused_split && (passes == 3 || transpose_a), prep_b = f32 && !fused_split && (passes == 3 || transpose_b);
    if (f32 && ((!prep_a && !aligned_operand(g.a, g.a_off, g.lda, g.sa, g.nmats, 4)) ||
                (!prep_b && !aligned_operand(g.b, g.b_off, g.ldb, g.sb, g.nmats, 4)))) {
        if (g.fused && g.fused->nranks > 1)   // the FFMA kernel has no peer stores: never drop the all-gather silently
            WGB_FAIL(WGB_ERR_UNSUPPORTED, "fused all-gather: single-pass TF32 needs 16-byte aligned operand views");
        *path_out = 1;   // single-pass TF32 reads the caller's views directly: needs TMA alignment
        return launch_gemm_simt(p, g);
    }
    if (g.nmats > 65535 && f32) {
        if (g.fused && g.fused->nranks > 1) WGB_FAIL(WGB_ERR_UNSUPPORTED, "fused all-gather: too many matrices in the batch");
        *path_out = 1;
        return launch_gemm_simt(p, g);
    }
    const size_t es = f32 ? 4 : 2;
    // CTA pairs own 256 output rows; matrices of at most 128 rows (batches of small matrices) run one CTA per tile instead
    const int cg = env_int("WGB_TC_CG", g.M <= 128 ? 1 : 2) == 1 ? 1 : 2;
    const uint32_t sms_total = (uint32_t)ctx->prop.multiProcessorCount;
    const int bn = passes == 3 ? 128 : pick_bn(g.M, g.N, g.nmats, cg, sms_total);
    const uint32_t block_k = 128 / (uint32_t)es;

    bool tr = g.tr;
    uint32_t a_rows = g.tr ? g.K : g.M, a_cols = g.tr ? g.M : g.K;
    const char *a_ptr = (const char *)g.a + g.a_off * es, *b_ptr = (const char *)g.b + g.b_off * es;
    const char *alo_ptr = a_ptr, *blo_ptr = b_ptr;
    uint64_t lda = g.lda, ldb = g.ldb, sa = g.sa, sb = g.sb;
    const int sgrid = ctx->prop.multiProcessorCount * 8;
    if (prep_a) {
        // dense re-materialised operands in the context workspace: slot 0 = A (hi, lo), slot 1 = B (hi, lo)
        const uint32_t oa_rows = transpose_a ? g.K : a_rows, oa_cols = transpose_a ? g.M : a_cols;
        const uint64_t lda_d = ((uint64_t)oa_rows + 3) & ~3ull;
        const uint64_t sa_d = lda_d * oa_cols;
        const size_t a_bytes = sa_d * g.nmats * 4;
        void *wa = nullptr;
        WGB_TRY(workspace_reserve(ctx, 0, (passes == 3 ? 2 : 1) * a_bytes, &wa));
        float *ahi = (float *)wa, *alo = passes == 3 ? (float *)((char *)wa + a_bytes) : nullptr;
        if (transpose_a) {
            dim3 tg((g.M + 31) / 32, (g.K + 31) / 32, g.nmats);
            if (tg.y > 65535) WGB_FAIL(WGB_ERR_UNSUPPORTED, "gemm: K too large for the transposing prep");
            split_tf32_transpose_kernel<<<tg, 256, 0, p->stream>>>((const float *)a_ptr, g.M, g.K, g.lda, g.sa, ahi, alo, lda_d, sa_d);
            tr = true;
            a_rows = g.K; a_cols = g.M;
        } else {
            split_tf32_kernel<<<sgrid, 256, 0, p->stream>>>((const float *)a_ptr, a_rows, a_cols, g.nmats, g.lda, g.sa, ahi, alo, lda_d, sa_d);
        }
        count_launch(ctx);
        a_ptr = (const char *)ahi; alo_ptr = (const char *)alo;
        lda = lda_d; sa = sa_d;
        WGB_CUDA(cudaGetLastError());
    }
    if (prep_b) {
        const uint64_t ldb_d = ((uint64_t)g.K + 3) & ~3ull, sb_d = ldb_d * g.N;
        const size_t b_bytes = sb_d * g.nmats * 4;
        void *wb = nullptr;
        WGB_TRY(workspace_reserve(ctx, 1, (passes == 3 ? 2 : 1) * b_bytes, &wb));
        float *bhi = (float *)wb, *blo = passes == 3 ? (float *)((char *)wb + b_bytes) : nullptr;
        if (transpose_b) {   // memory holds [N x K] with N contiguous: write [K x N] with K contiguous
            dim3 tg((g.N + 31) / 32, (g.K + 31) / 32, g.nmats);
            if (tg.y > 65535) WGB_FAIL(WGB_ERR_UNSUPPORTED, "gemm: K too large for the transposing prep");
            split_tf32_transpose_kernel<<<tg, 256, 0, p->stream>>>((const float *)b_ptr, g.N, g.K, g.ldb, g.sb, bhi, blo, ldb_d, sb_d);
        } else {
            split_tf32_kernel<<<sgrid, 256, 0, p->stream>>>((const float *)b_ptr, g.K, g.N, g.nmats, g.ldb, g.sb, bhi, blo, ldb_d, sb_d);
        }
        count_launch(ctx);
        b_ptr = (const char *)bhi; blo_ptr = (const char *)blo;
        ldb = ldb_d; sb = sb_d;
        WGB_CUDA(cudaGetLastError());
    }
    const bool b_mn = !f32 && g.b_nmajor;   // bf16: MN-major B read in place

    TcMaps maps{};
    CUtensorMap &ta = maps.a, &talo = maps.alo, &tb = maps.b, &tblo = maps.blo;
    const bool atom32 = f32 && !tr;                                     // MN-major 32-bit operand A
    const uint32_t a_box0 = tr ? block_k : (uint32_t)(128 / es);       // K-major: 128 B of K; MN-major: one 128 B atom of M
    const uint32_t a_box1 = tr ? 128u : block_k;
    WGB_TRY(get_tmap(ctx, a_ptr, (uint32_t)es, a_rows, a_cols, g.nmats, lda, sa, a_box0, a_box1, &ta, atom32));
    if (b_mn) WGB_TRY(get_tmap(ctx, b_ptr, (uint32_t)es, g.N, g.K, g.nmats, ldb, sb, (uint32_t)(128 / es), block_k, &tb));   // one 128 B atom of N
    else WGB_TRY(get_tmap(ctx, b_ptr, (uint32_t)es, g.K, g.N, g.nmats, ldb, sb, block_k, (uint32_t)(bn / cg), &tb));
    if (passes == 3 && !fused_split) {
        WGB_TRY(get_tmap(ctx, alo_ptr, (uint32_t)es, a_rows, a_cols, g.nmats, lda, sa, a_box0, a_box1, &talo, atom32));
        WGB_TRY(get_tmap(ctx, blo_ptr, (uint32_t)es, g.K, g.N, g.nmats, ldb, sb, block_k, (uint32_t)(bn / cg), &tblo));
    } else {
        talo = ta;
        tblo = tb;
    }

    TcArgs args{};
    args.c = (char *)g.c + g.c_off * dtype_size(g.out_dtype);
    args.npeers = 1;
    args.store_mask = 0xFFFFFFFFu;
    args.dst[0] = (char *)args.c;
    args.mbar_timeout = 8000000000ll;
    if (g.fused && g.fused->nranks > 1) {
        const FusedGather &f = *g.fused;
        args.npeers = (uint32_t)f.nranks;
        args.handshake = 1;
        args.my_rank = (uint32_t)f.rank;
        args.epoch = f.epoch;
        for (int r = 0; r < f.nranks; ++r) {
            args.dst[r] = (char *)f.peer_c[r] + g.c_off * dtype_size(g.out_dtype);
            args.done_remote[r] = f.done_remote[r];
        }
        args.ready_local = f.ready_local;
        args.cta_counter = f.cta_counter;
        args.peer_timeout = f.timeout;
        if (env_int("WGB_FUSED_ROTATE", 1) != 0) args.first_dst = (uint32_t)((f.rank + 1) % f.nranks);
        // diagnostics only (results on the skipped ranks are garbage): which destinations the epilogue really stores to
        const int mask = env_int("WGB_FUSED_DEBUG_STORE_MASK", -1);
        if (mask == -2) args.store_mask = 1u << f.rank;              // local copy only: the GEMM and the handshake without NVLink traffic
        else if (mask >= 0) args.store_mask = (uint32_t)mask;
        args.mbar_timeout = f.timeout > 0 ? f.timeout + 8000000000ll : 0;   // the pipeline stalls behind an epilogue that waits for peers
    }
    args.ldc = g.ldc; args.sc = g.sc;
    args.M = g.M; args.N = g.N; args.K = g.K; args.nmats = g.nmats;
    args.tiles_m = (g.M + 128 * cg - 1) / (128 * cg);
    args.tiles_n = (g.N + bn - 1) / bn;
    args.num_kb = (g.K + block_k - 1) / block_k;
    const uint64_t total = (uint64_t)args.tiles_m * args.tiles_n * g.nmats;
    if (total > 0x7FFFFFFFull) WGB_FAIL(WGB_ERR_UNSUPPORTED, "gemm: too many output tiles");
    args.total_tiles = (uint32_t)total;
    uint32_t sms = sms_total;
    const uint32_t margin = (uint32_t)comm_sm_margin(ctx);
    if (margin < sms / 2) sms -= margin;
    if (args.npeers == 1) {
        // diagnostics only (WGB_TC_DEBUG_FAKE_PEERS=n): issue the epilogue stores n times (all to the local buffer) to measure the
        // SM-side cost of the fused all-gather's extra stores without NVLink
        const int fake = env_int("WGB_TC_DEBUG_FAKE_PEERS", 0);
        if (fake > 1 && fake <= kMaxPeers) {
            // replicas 1..n-1 go to distinct scratch panels (like distinct peers), replica 0 is the real output
            const size_t c_bytes = ((size_t)(g.nmats - 1) * g.sc + (size_t)(g.N - 1) * g.ldc + g.M) * dtype_size(g.out_dtype);
            void *scratch = nullptr;
            if (workspace_reserve(ctx, 6, (size_t)(fake - 1) * c_bytes, &scratch) == WGB_OK) {
                args.npeers = (uint32_t)fake;
                for (int r = 1; r < fake; ++r) args.dst[r] = (char *)scratch + (size_t)(r - 1) * c_bytes;
            }
        }
    }
    args.fused_split = fused_split ? 1u : 0u;
    args.red_axis = (uint32_t)g.red_axis;
    args.red_op = g.red_op;
    args.red_partials = g.red_partials;
    args.ep_op = g.ep_op;
    args.ep = g.e ? (const char *)g.e + g.e_off * dtype_size(g.out_dtype) : nullptr;
    args.ep_ld = g.lde;
    args.ep_sm = g.se;
    args.trace = ctx->tc_trace;
    args.debug_skip = (uint32_t)env_int("WGB_TC_DEBUG_SKIP", 0) & 3u;   // timing diagnostics only: results are garbage
    plan_tail(args, sms / cg, bn, cg, f32 ? 1 : 0, !tr, b_mn);
    maps.bt = tb;
    maps.blot = tblo;
    if (args.nsplit > 1 && !b_mn) {   // (MN-major B: the box is one atom whatever the strip width)
        WGB_TRY(get_tmap(ctx, b_ptr, (uint32_t)es, g.K, g.N, g.nmats, ldb, sb, block_k, args.tail_bn / (uint32_t)cg, &maps.bt));
        if (passes == 3 && !fused_split) WGB_TRY(get_tmap(ctx, blo_ptr, (uint32_t)es, g.K, g.N, g.nmats, ldb, sb, block_k, args.tail_bn / (uint32_t)cg, &maps.blot));
        else maps.blot = maps.bt;
    }
    if (args.split > 1) {
        const uint32_t tail = args.total_tiles - args.full_tiles;
        if ((uint64_t)tail * cg > ctx->scratch.n_counters) {
            args.full_tiles = args.total_tiles; args.split = 1; args.kb_per_split = args.num_kb; args.total_units = args.total_tiles;
        } else {
            void *w = nullptr;
            WGB_TRY(workspace_reserve(ctx, 2, (size_t)tail * args.split * cg * bn * 128 * 4, &w));
            args.ws = (float *)w;
            args.counters = ctx->scratch.counters;
        }
    }

    // Epilogue form.  WGB_TC_EPI: 0 = per-lane global stores, 1 = shared-memory staged TMA bulk stores, unset = TMA stores when
    // the all-gather is fused in (every output block then leaves the SM once per rank of the box).
    {
        const size_t os = dtype_size(g.out_dtype);
        const bool fused = g.fused && g.fused->nranks > 1;
        const bool mc_ok = fused && g.fused->mc_c != nullptr;
        // fused all-gather: multicast stores when the group has a multicast mapping, else TMA bulk stores to every rank
        int want = env_int("WGB_TC_EPI", fused ? (mc_ok ? 2 : 1) : 0);
        if (want == 2 && !mc_ok) want = fused ? 1 : 0;
        bool ok = want != 0 && (g.ldc * os) % 16 == 0 && (g.nmats <= 1 || (g.sc * os) % 16 == 0);
        for (uint32_t d = 0; ok && d < args.npeers; ++d) ok = ((uintptr_t)args.dst[d] & 15u) == 0;
        if (ok && want == 2) {
            // 16-byte pieces must not straddle the last row: M a multiple of the piece (8 bf16 / 4 f32 rows)
            char *mc = (char *)g.fused->mc_c + g.c_off * os;
            if (((uintptr_t)mc & 15u) == 0 && (g.M * os) % 16 == 0) {
                args.epi_tma = 2;
                args.mc_dst = mc;
            } else {
                want = 1;
            }
        }
        if (ok && want == 1) {
            for (uint32_t d = 0; d < args.npeers; ++d)
                WGB_TRY(get_tmap(ctx, args.dst[d], (uint32_t)os, g.M, g.N, g.nmats, g.ldc, g.sc, 128u, (uint32_t)kEpiCols, &maps.dst.m[d],
                                 false, true));
            args.epi_tma = 1;
        }
    }

    {   // what ran, for wgb_pass_last_gemm_config (tests name every kernel instantiation through it)
        int *c = p->last_tc;
        c[0] = f32 ? 1 : 0; c[1] = tr ? 0 : 1; c[2] = b_mn ? 1 : 0; c[3] = bn; c[4] = passes; c[5] = (int)g.out_dtype; c[6] = cg;
        c[7] = (int)args.epi_tma; c[8] = (int)args.nsplit; c[9] = (int)args.split; c[10] = (int)args.npeers; c[11] = (int)args.total_units;
        c[12] = (int)args.fused_split;
    }
    wgb_status st;
    if (!f32) {
        if (g.out_dtype == WGB_F32) {
            if (b_mn) st = tr ? launch_sel<0, false, true, 1, float>(p, bn, cg, maps, args) : launch_sel<0, true, true, 1, float>(p, bn, cg, maps, args);
            else st = tr ? launch_sel<0, false, false, 1, float>(p, bn, cg, maps, args) : launch_sel<0, true, false, 1, float>(p, bn, cg, maps, args);
        } else {
            if (b_mn) st = tr ? launch_sel<0, false, true, 1, __nv_bfloat16>(p, bn, cg, maps, args)
                              : launch_sel<0, true, true, 1, __nv_bfloat16>(p, bn, cg, maps, args);
            else st = tr ? launch_sel<0, false, false, 1, __nv_bfloat16>(p, bn, cg, maps, args)
                         : launch_sel<0, true, false, 1, __nv_bfloat16>(p, bn, cg, maps, args);
        }
        *path_out = 2;
    } else if (passes == 1) {
        st = tr ? launch_sel<1, false, false, 1, float>(p, bn, cg, maps, args) : launch_sel<1, true, false, 1, float>(p, bn, cg, maps, args);
        *path_out = 3;
    } else {
        st = tr ? launch_sel<1, false, false, 3, float>(p, bn, cg, maps, args) : launch_sel<1, true, false, 3, float>(p, bn, cg, maps, args);
        *path_out = 4;
    }
    return st;
}

}  // namespace wgb
