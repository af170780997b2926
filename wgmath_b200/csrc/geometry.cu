// geometry.cu — batched small-matrix factorizations: out[i] = f(in[i]) over a device array, one matrix per thread.
//
// This is the shape of every test kernel in the reference's geometry module (e.g. crates/wgebra/src/geometry/cholesky.rs:53-63,
// lu.rs:101-111, eig3.rs:36-46: `out[i] = cholesky(in[i])` with @workgroup_size(1,1,1), i.e. ONE invocation per workgroup)
// done the way the hardware wants it:
//   * a warp owns 32 consecutive elements; their storage (32 x 16..64 B in, 32 x 16..128 B out) is one contiguous span,
//     moved with fully coalesced 128-byte warp transactions through padded shared-memory tiles, so HBM sees each byte once
//     although every thread works on its own array-of-structs element; 3x3 / 4x4 input tiles arrive through a 16-byte
//     cp.async ring so the next tile's loads are in flight while the current one is factorized and stored, 2x2 matrices are
//     one float4 per lane and are loaded directly, four tiles ahead;
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

// One element, registers to registers: iw = the input matrix in storage order, ow = the output struct in storage order
// (padding words zero).  Everything between is geometry.cuh.
template <int OP, int D>
__device__ __forceinline__ void apply(const float (&iw)[mat_words(D)], float (&ow)[out_words(OP, D)]) {
    constexpr int MW = mat_words(D), CS = col_stride(D), OW = out_words(OP, D);
    geom::Mat<D> x;
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) x.m[c][r] = iw[c * CS + r];
#pragma unroll
    for (int k = 0; k < OW; ++k) ow[k] = 0.0f;
    auto put_mat = [&](int at, const geom::Mat<D> &m) {
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
            for (int r = 0; r < D; ++r) ow[at + c * CS + r] = m.m[c][r];
    };
    auto put_vec = [&](int at, const float(&v)[D]) {
#pragma unroll
        for (int r = 0; r < D; ++r) ow[at + r] = v[r];
    };
    if constexpr (OP == WGB_GEOM_CHOLESKY) {
        put_mat(0, geom::cholesky<D>(x));
    } else if constexpr (OP == WGB_GEOM_INV) {
        put_mat(0, geom::inverse(x));
    } else if constexpr (OP == WGB_GEOM_LU) {
        const geom::LU<D> f = geom::lu<D>(x);
        put_mat(0, f.lu);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            ow[MW + k] = __uint_as_float(f.ia[k]);
            ow[MW + CS + k] = __uint_as_float(f.ib[k]);
        }
        ow[MW + (D == 3 ? 7 : 2 * CS)] = __uint_as_float(f.len);   // a vec3<u32> is 12 bytes: `len` packs right behind `ib`
    } else if constexpr (OP == WGB_GEOM_QR) {
        const geom::QR<D> f = geom::qr<D>(x);
        put_mat(0, f.q);
        put_mat(MW, f.r);
    } else if constexpr (OP == WGB_GEOM_SYMMETRIC_EIGEN) {
        const geom::SymmetricEigen<D> f = geom::symmetric_eigen(x);
        put_mat(0, f.eigenvectors);
        put_vec(MW, f.eigenvalues);
    } else {
        const geom::Svd<D> f = geom::svd(x);
        put_mat(0, f.U);
        put_vec(MW, f.S);
        put_mat(MW + CS, f.Vt);
    }
}

__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int V> struct VecT;
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

// Row pitch (in words) of a shared-memory tile holding 32 structs of W words that lanes access V words at a time with a stride
// of one struct per lane: conflict-free when pitch / V is odd.
template <int W, int V>
constexpr int tile_pitch() { return (W / V) % 2 == 1 ? W : W + V; }

// Stage this lane's output struct into the warp's tile (V-word vector stores), then write the tile's span to global memory
// with coalesced V-word vector stores: chunk q of the span belongs to element q / (OW / V).
template <int OW, int V>
__device__ __forceinline__ void store_tile(float *otile, const float (&ow)[OW], float *dst, int lane, int valid) {
    using Vec = typename VecT<V>::type;
    constexpr int P = tile_pitch<OW, V>(), C = OW / V;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        Vec v;
        if constexpr (V == 4) v = make_float4(ow[4 * c], ow[4 * c + 1], ow[4 * c + 2], ow[4 * c + 3]);
        else v = make_float2(ow[2 * c], ow[2 * c + 1]);
        *reinterpret_cast<Vec *>(otile + lane * P + V * c) = v;
    }
    __syncwarp();
    Vec *d = reinterpret_cast<Vec *>(dst);
    if (valid == 32) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int q = k * 32 + lane;
            d[q] = *reinterpret_cast<const Vec *>(otile + (q / C) * P + V * (q % C));
        }
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int q = k * 32 + lane;
            if (q < valid * C) d[q] = *reinterpret_cast<const Vec *>(otile + (q / C) * P + V * (q % C));
        }
    }
    __syncwarp();   // the tile is free again
}

// ---- 3x3 / 4x4: 48- / 64-byte inputs arrive through a ring of cp.async (16-byte) staged tiles per warp, so the next tile's
// loads are in flight while the current one is factorized and stored.
template <int OP, int D>
__device__ __forceinline__ void batch_body_staged(const float *__restrict__ in, float *__restrict__ out, uint64_t n) {
    constexpr int IW = mat_words(D), OW = out_words(OP, D), CI = IW / 4;
    constexpr int PI = tile_pitch<IW, 4>(), PO = tile_pitch<OW, 4>();
    constexpr int kStages = 2;
    __shared__ __align__(16) float in_tiles[kGeomWarps][kStages][32 * PI];
    __shared__ __align__(16) float out_tiles[kGeomWarps][32 * PO];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_tiles = (n + 31) / 32;
    const uint64_t stride = (uint64_t)gridDim.x * kGeomWarps;
    auto prefetch = [&](uint64_t t, int stage) {   // 16-byte chunk q of the tile's span belongs to element q / CI
        if (t < n_tiles) {
            const uint64_t first = t * 32;
            const int valid = n - first < 32 ? (int)(n - first) : 32;
            const float *src = in + first * IW;
            float *tile = in_tiles[warp][stage];
#pragma unroll
            for (int k = 0; k < CI; ++k) {
                const int q = k * 32 + lane;
                if (q < valid * CI) cp_async_16(tile + (q / CI) * PI + 4 * (q % CI), src + 4 * q);
            }
        }
        cp_async_commit();   // one group per slot, empty past the end, so wait_group counts stay uniform
    };
    uint64_t t = (uint64_t)blockIdx.x * kGeomWarps + warp;
    prefetch(t, 0);
    int stage = 0;
    for (; t < n_tiles; t += stride) {
        prefetch(t + stride, stage ^ 1);
        cp_async_wait<kStages - 1>();   // this lane's copies of tile t have landed ...
        __syncwarp();                   // ... and so have every other lane's
        const uint64_t first = t * 32;
        const int valid = n - first < 32 ? (int)(n - first) : 32;
        float iw[IW], ow[OW];
        const float *e = in_tiles[warp][stage] + lane * PI;
#pragma unroll
        for (int c = 0; c < CI; ++c) {
            const float4 v = *reinterpret_cast<const float4 *>(e + 4 * c);
            iw[4 * c] = v.x, iw[4 * c + 1] = v.y, iw[4 * c + 2] = v.z, iw[4 * c + 3] = v.w;
        }
        if (lane >= valid) {   // lanes past the end factorize the identity (their smem slot was never written)
#pragma unroll
            for (int k = 0; k < IW; ++k) iw[k] = (k % (col_stride(D) + 1)) == 0 ? 1.0f : 0.0f;
        }
        apply<OP, D>(iw, ow);
        store_tile<OW, 4>(out_tiles[warp], ow, out + first * OW, lane, valid);
        stage ^= 1;
    }
    cp_async_wait<0>();
}

// ---- 2x2: a matrix is one float4, so a lane loads its own element straight from global memory (the warp's 32 loads are one
// contiguous 512-byte span); four tiles per iteration keep four independent 128-bit loads in flight per lane.
template <int OP>
__device__ __forceinline__ void batch_body_2x2(const float *__restrict__ in, float *__restrict__ out, uint64_t n) {
    constexpr int OW = out_words(OP, 2), U = 4;
    constexpr int PO = tile_pitch<OW, 2>();
    __shared__ __align__(16) float out_tiles[kGeomWarps][32 * PO];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_super = (n + 32 * U - 1) / (32 * U);
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    for (uint64_t t = (uint64_t)blockIdx.x * kGeomWarps + warp; t < n_super; t += (uint64_t)gridDim.x * kGeomWarps) {
        const uint64_t base = t * (32 * U);
        float4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t i = base + j * 32 + lane;
            v[j] = i < n ? __ldg(in4 + i) : make_float4(1.0f, 0.0f, 0.0f, 1.0f);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t first = base + j * 32;
            if (first >= n) break;
            const int valid = n - first < 32 ? (int)(n - first) : 32;
            const float iw[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            float ow[OW];
            apply<OP, 2>(iw, ow);
            if constexpr (OW == 4) {   // matrix out: again one float4 per lane, no staging
                if (lane < valid) reinterpret_cast<float4 *>(out)[first + lane] = make_float4(ow[0], ow[1], ow[2], ow[3]);
            } else {
                store_tile<OW, 2>(out_tiles[warp], ow, out + first * OW, lane, valid);
            }
        }
    }
}

template <int OP, int D>
__global__ void __launch_bounds__(kGeomThreads) geom_batch_kernel(const float *__restrict__ in, float *__restrict__ out, uint64_t n) {
    if constexpr (D == 2) batch_body_2x2<OP>(in, out, n);
    else batch_body_staged<OP, D>(in, out, n);
}

template <int OP, int D>
wgb_status launch_one(wgb_pass *p, const float *in, float *out, uint64_t n) {
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        WGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, geom_batch_kernel<OP, D>, kGeomThreads, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const uint64_t per_block = 32ull * kGeomWarps * (D == 2 ? 4 : 1);   // elements one block covers per loop iteration
    const uint64_t want = (n + per_block - 1) / per_block;
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
    // 16-byte cp.async / float4 accesses; the 24- and 40-byte 2x2 structs are stored as float2 (only a wrapped, misaligned
    // pointer can fail this)
    const uintptr_t out_align = (dim == 2 && ob != 16) ? 7u : 15u;
    if (((uintptr_t)src & 15u) || ((uintptr_t)dst & out_align))
        WGB_FAIL(WGB_ERR_INVALID, "wgb_geometry_batch: element ranges must start 16-byte aligned");
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
