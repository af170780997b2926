// geometry.cu — batched small-matrix factorizations: out[i] = f(in[i]) over a device array, one matrix per thread.
//
// This is the shape of every test kernel in the reference's geometry module (e.g. crates/wgebra/src/geometry/cholesky.rs:53-63,
// lu.rs:101-111, eig3.rs:36-46: `out[i] = cholesky(in[i])` with @workgroup_size(1,1,1), i.e. ONE invocation per workgroup)
// done the way the hardware wants it:
//   * a warp owns 32 consecutive elements; their storage (32 x 16..64 B in, 32 x 16..128 B out) is one contiguous span,
//     moved with fully coalesced 128-byte warp transactions through padded shared-memory tiles, so HBM sees each byte once
//     although every thread works on its own array-of-structs element; input tiles arrive through a cp.async ring so the
//     next tiles' loads are in flight while the current one is factorized and stored;
//   * the element lives in registers while it is factorized (geometry.cuh: compile-time loops, predicated static indexing);
//   * grid = a multiple of the SM count, warps stride over tiles.
// HBM bound for the closed-form / direct ops (algorithmic bytes per element = in + out struct size); the iterative ones
// (eig3 / eig4 QR sweeps, svd3's 36 Jacobi conjugations) are FP32-pipe bound — bench.py reports both against the HBM roofline.
// Compiled with -fmad=false: see geometry.cuh.
#include "common.cuh"
#include "geometry.cuh"

namespace wgb {

namespace {

constexpr int kGeomThreads = 128;
constexpr int kGeomWarps = kGeomThreads / 32;

__host__ __device__ constexpr int col_stride(int d) { return d == 2 ? 2 : 4; }   // vec3 columns are padded to vec4
__host__ __device__ constexpr int mat_words(int d) { return d * col_stride(d); }
__host__ __device__ constexpr int out_words(int op, int d) {
    return (op == WGB_GEOM_CHOLESKY || op == WGB_GEOM_INV) ? mat_words(d)
           : op == WGB_GEOM_LU                             ? (d == 2 ? 10 : d == 3 ? 20 : 28)
           : op == WGB_GEOM_QR                             ? 2 * mat_words(d)
           : op == WGB_GEOM_SYMMETRIC_EIGEN                ? mat_words(d) + col_stride(d)
           : (op == WGB_GEOM_SVD && d < 4)                 ? 2 * mat_words(d) + col_stride(d)
                                                           : 0;
}

template <int D>
__device__ __forceinline__ geom::Mat<D> tile_load_mat(const float *e) {
    geom::Mat<D> x;
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) x.m[c][r] = e[c * col_stride(D) + r];
    return x;
}
template <int D>
__device__ __forceinline__ void tile_store_mat(float *e, const geom::Mat<D> &x) {
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < col_stride(D); ++r) e[c * col_stride(D) + r] = r < D ? x.m[c][r < D ? r : 0] : 0.0f;
}
template <int D>
__device__ __forceinline__ void tile_store_vec(float *e, const float (&v)[D]) {
#pragma unroll
    for (int r = 0; r < col_stride(D); ++r) e[r] = r < D ? v[r < D ? r : 0] : 0.0f;
}

template <int OP, int D>
__device__ __forceinline__ void apply(const float *src, float *e) {   // this lane's element: input tile -> output tile
    const geom::Mat<D> x = tile_load_mat<D>(src);
    constexpr int MW = mat_words(D), CS = col_stride(D);
    if constexpr (OP == WGB_GEOM_CHOLESKY) {
        tile_store_mat<D>(e, geom::cholesky<D>(x));
    } else if constexpr (OP == WGB_GEOM_INV) {
        tile_store_mat<D>(e, geom::inverse(x));
    } else if constexpr (OP == WGB_GEOM_LU) {
        const geom::LU<D> f = geom::lu<D>(x);
        tile_store_mat<D>(e, f.lu);
        uint32_t *p = reinterpret_cast<uint32_t *>(e + MW);
#pragma unroll
        for (int k = 0; k < out_words(OP, D) - MW; ++k) p[k] = 0u;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            p[k] = f.ia[k];
            p[CS + k] = f.ib[k];
        }
        p[D == 3 ? 7 : 2 * CS] = f.len;   // a vec3<u32> is 12 bytes: `len` packs right behind `ib`
    } else if constexpr (OP == WGB_GEOM_QR) {
        const geom::QR<D> f = geom::qr<D>(x);
        tile_store_mat<D>(e, f.q);
        tile_store_mat<D>(e + MW, f.r);
    } else if constexpr (OP == WGB_GEOM_SYMMETRIC_EIGEN) {
        const geom::SymmetricEigen<D> f = geom::symmetric_eigen(x);
        tile_store_mat<D>(e, f.eigenvectors);
        tile_store_vec<D>(e + MW, f.eigenvalues);
    } else {
        const geom::Svd<D> f = geom::svd(x);
        tile_store_mat<D>(e, f.U);
        tile_store_vec<D>(e + MW, f.S);
        tile_store_mat<D>(e + MW + CS, f.Vt);
    }
}

__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Input tiles are prefetched with cp.async into a ring of kStages buffers per warp, so a warp always has kStages - 1 tiles of
// loads in flight while it factorizes / stores the current one (one tile of 2x2 matrices is only 512 bytes: four stages
// there, two for the 48- / 64-byte matrices).
template <int D>
constexpr int geom_stages() { return D == 2 ? 4 : 2; }

template <int OP, int D>
__global__ void __launch_bounds__(kGeomThreads) geom_batch_kernel(const float *__restrict__ in, float *__restrict__ out, uint64_t n) {
    constexpr int IW = mat_words(D), OW = out_words(OP, D);
    constexpr int TI = IW + 1, TO = OW + 1;   // odd row pitch: lane-strided element access is bank-conflict free
    constexpr int kStages = geom_stages<D>();
    __shared__ float in_tiles[kGeomWarps][kStages][32 * TI];
    __shared__ float out_tiles[kGeomWarps][32 * TO];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *otile = out_tiles[warp];
    const uint64_t n_tiles = (n + 31) / 32;
    const uint64_t stride = (uint64_t)gridDim.x * kGeomWarps;
    auto prefetch = [&](uint64_t t, int stage) {   // word w of the tile's span belongs to element w / IW
        if (t < n_tiles) {
            const uint64_t first = t * 32;
            const int valid = n - first < 32 ? (int)(n - first) : 32;
            const float *src = in + first * IW;
            float *tile = in_tiles[warp][stage];
#pragma unroll
            for (int k = 0; k < IW; ++k) {
                const int w = k * 32 + lane;
                if (w < valid * IW) cp_async_4(tile + (w / IW) * TI + (w % IW), src + w);
            }
        }
        cp_async_commit();   // one group per slot, empty past the end, so wait_group counts stay uniform
    };
    uint64_t t = (uint64_t)blockIdx.x * kGeomWarps + warp;
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) prefetch(t + s * stride, s);
    int stage = 0;
    for (; t < n_tiles; t += stride) {
        prefetch(t + (kStages - 1) * stride, (stage + kStages - 1) % kStages);
        cp_async_wait<kStages - 1>();   // this lane's copies of tile t have landed ...
        __syncwarp();                   // ... and so have every other lane's
        const uint64_t first = t * 32;
        const int valid = n - first < 32 ? (int)(n - first) : 32;
        if (lane < valid) apply<OP, D>(in_tiles[warp][stage] + lane * TI, otile + lane * TO);
        __syncwarp();
        float *dst = out + first * OW;
#pragma unroll
        for (int k = 0; k < OW; ++k) {
            const int w = k * 32 + lane;
            if (w < valid * OW) dst[w] = otile[(w / OW) * TO + (w % OW)];
        }
        __syncwarp();   // the out tile and this input stage are free again
        stage = (stage + 1) % kStages;
    }
    cp_async_wait<0>();
}

template <int OP, int D>
wgb_status launch_one(wgb_pass *p, const float *in, float *out, uint64_t n) {
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        WGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, geom_batch_kernel<OP, D>, kGeomThreads, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const uint64_t want = (n + 32 * kGeomWarps - 1) / (32 * kGeomWarps);
    const uint64_t cap = (uint64_t)p->ctx->prop.multiProcessorCount * blocks_per_sm;   // one resident wave, warps stride
    geom_batch_kernel<OP, D><<<(unsigned)(want < cap ? want : cap), kGeomThreads, 0, p->stream>>>(in, out, n);
    WGB_CUDA(cudaGetLastError());
    count_launch(p->ctx);
    return WGB_OK;
}

template <int OP>
wgb_status launch_dim(wgb_pass *p, int dim, const float *in, float *out, uint64_t n) {
    switch (dim) {
    case 2: return launch_one<OP, 2>(p, in, out, n);
    case 3: return launch_one<OP, 3>(p, in, out, n);
    case 4:
        if constexpr (OP != WGB_GEOM_SVD) return launch_one<OP, 4>(p, in, out, n);
    }
    WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: op %d has no %dx%d variant", OP, dim, dim);
}

}  // namespace

}  // namespace wgb

using namespace wgb;

extern "C" {

uint32_t wgb_geometry_in_bytes(int dim) { return dim < 2 || dim > 4 ? 0u : 4u * (uint32_t)mat_words(dim); }
uint32_t wgb_geometry_out_bytes(wgb_geom_op op, int dim) {
    return dim < 2 || dim > 4 ? 0u : 4u * (uint32_t)out_words((int)op, dim);
}

wgb_status wgb_geometry_batch(wgb_pass *pass, wgb_geom_op op, int dim, const wgb_buffer *in, uint64_t in_first, wgb_buffer *out,
                              uint64_t out_first, uint64_t n) {
    if (!pass || !in || !out) WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: null argument");
    const uint32_t ib = wgb_geometry_in_bytes(dim), ob = wgb_geometry_out_bytes(op, dim);
    if (!ib || !ob || (int)op < 0 || (int)op > WGB_GEOM_INV)
        WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: op %d has no %dx%d variant", (int)op, dim, dim);
    if (n == 0 || in->bytes == 0 || out->bytes == 0) return WGB_OK;   // kernel.rs:111-113,144: nothing to queue
    if (in->host_pinned || out->host_pinned) WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: MAP_READ staging buffers cannot be bound");
    if (in_first > in->bytes / ib || n > in->bytes / ib - in_first)
        WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_geometry_batch: input elements [%llu, %llu) exceed the buffer (%zu bytes, %u per element)",
                 (unsigned long long)in_first, (unsigned long long)(in_first + n), in->bytes, ib);
    if (out_first > out->bytes / ob || n > out->bytes / ob - out_first)
        WGB_FAIL(WGB_ERR_OUT_OF_BOUNDS, "wgb_geometry_batch: output elements [%llu, %llu) exceed the buffer (%zu bytes, %u per element)",
                 (unsigned long long)out_first, (unsigned long long)(out_first + n), out->bytes, ob);
    const char *src = (const char *)in->ptr + in_first * ib;
    char *dst = (char *)out->ptr + out_first * ob;
    // in place is fine when element i of the output covers exactly element i of the input (a warp reads its tile before
    // it writes it); any other overlap races between warps
    if (src < dst + n * ob && dst < src + n * ib && !(src == dst && ib == ob))
        WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: input and output ranges overlap");
    DeviceGuard g(pass->ctx->device);
    const float *fi = (const float *)src;
    float *fo = (float *)dst;
    switch (op) {
    case WGB_GEOM_CHOLESKY: return launch_dim<WGB_GEOM_CHOLESKY>(pass, dim, fi, fo, n);
    case WGB_GEOM_LU: return launch_dim<WGB_GEOM_LU>(pass, dim, fi, fo, n);
    case WGB_GEOM_QR: return launch_dim<WGB_GEOM_QR>(pass, dim, fi, fo, n);
    case WGB_GEOM_SYMMETRIC_EIGEN: return launch_dim<WGB_GEOM_SYMMETRIC_EIGEN>(pass, dim, fi, fo, n);
    case WGB_GEOM_SVD: return launch_dim<WGB_GEOM_SVD>(pass, dim, fi, fo, n);
    case WGB_GEOM_INV: return launch_dim<WGB_GEOM_INV>(pass, dim, fi, fo, n);
    }
    WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: unknown op %d", (int)op);
}

}  // extern "C"
