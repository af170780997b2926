// Replays the reference's four unit tests (gemm.rs:144-202, gemv.rs:153-197, op_assign.rs:109-157,
// reduce.rs:139-179) through the C++ host mirror (wgmath_b200/host/wgebra_b200.hpp -> C ABI -> CUDA) and
// checks every result against the CPU oracle (oracle/wgsl_oracle.c) on the same seeded inputs.
// Exit code 0 = all within tolerance (1e-5 relative; op_assign bit-exact).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../wgmath_b200/host/wgebra_b200.hpp"

extern "C" {
typedef struct { uint32_t nrows, ncols, nmats, stride, stride_mat, offset; } orc_shape;
int orc_gemm(int variant, float *out, const orc_shape *so, const float *m1, const orc_shape *s1, const float *m2, const orc_shape *s2);
int orc_gemv(int variant, float *out, const orc_shape *so, const float *m, const orc_shape *sm, const float *v, const orc_shape *sv, int *ran);
int orc_op_assign(int op, float *a, const orc_shape *sa, const float *b, const orc_shape *sb);
int orc_reduce(int op, const float *x, const orc_shape *s, float *result);
int orc_prefix_sum(uint32_t *data, uint32_t n);
int orc_radix_sort(const uint32_t *input_keys, const uint32_t *input_values, uint32_t len, uint32_t n_sort, uint32_t sorting_bits,
                   uint32_t *output_keys, uint32_t *output_values);
int orc_geom_batch(int op, int dim, const float *in, void *out, uint64_t n);
int orc_gemm_ord(int variant, float *out, const orc_shape *so, int out_rm, const float *m1, const orc_shape *s1, int m1_rm, const float *m2,
                 const orc_shape *s2, int m2_rm);
}

static uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static std::vector<float> uniform(uint64_t seed, uint32_t rows, uint32_t cols = 1) {   // == oracle.uniform / wgb_fill_uniform
    std::vector<float> v((size_t)rows * cols);
    for (uint32_t j = 0; j < cols; ++j)
        for (uint32_t i = 0; i < rows; ++i)
            v[(size_t)j * rows + i] = (float)(splitmix64((seed * 0xD1342543DE82EF95ull) ^ (((uint64_t)j << 32) | i)) >> 40) * 5.9604644775390625e-8f;
    return v;
}
static orc_shape oshape(uint32_t r, uint32_t c = 1) { return orc_shape{r, c, 1, r, r * c, 0}; }
static double rel_err(const std::vector<float> &got, const std::vector<float> &ref) {
    double w = 0;
    for (size_t i = 0; i < ref.size(); ++i) w = std::fmax(w, std::fabs((double)got[i] - ref[i]) / std::fmax(std::fabs((double)ref[i]), 1e-30));
    return w;
}

int main() {
    using namespace wgb;
    const uint64_t SEED = 0x5EED0000;
    int failures = 0;
    try {
        GpuInstance gpu(0);
        ViewShapeBuffers shapes;
        const uint32_t usage = STORAGE | COPY_SRC | COPY_DST;
        {   // gpu_gemm
            const uint32_t n = 256;
            auto m1c = uniform(SEED + 1, n, n), m2c = uniform(SEED + 2, n, n);
            auto gemm = Gemm::from_device(gpu.device());
            auto m1 = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), m1c);
            auto m2 = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), m2c);
            auto result = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), std::vector<float>(n * n, 0.f));
            auto staging = Tensors::matrix(n, n, MAP_READ | COPY_DST).build<float>(gpu.device());
            for (int variant = 0; variant < 4; ++variant) {
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    gemm.dispatch_generic<float>(gpu.device(), shapes, pass, result.as_embedded_view<3>(), m1.as_embedded_view<3>(),
                                                 m2.as_embedded_view<3>(), (GemmVariant)variant);
                }
                staging.copy_from(enc, result);
                gpu.queue().submit(enc.finish());
                auto got = staging.read(gpu.device());
                std::vector<float> ref(n * n, 0.f);
                orc_shape s = oshape(n, n);
                orc_gemm(variant, ref.data(), &s, m1c.data(), &s, m2c.data(), &s);
                const double e = rel_err(got, ref);
                std::printf("gpu_gemm variant %d: rel err %.3e\n", variant, e);
                failures += !(e < 1e-5);
            }
        }
        {   // gpu_gemv
            const uint32_t n = 1024;
            auto mc = uniform(SEED + 1, n, n), vc = uniform(SEED + 3, n), oc = uniform(SEED + 4, n);
            auto gemv = Gemv::from_device(gpu.device());
            auto m = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), mc);
            auto v = Tensors::vector(n, usage).build_init<float>(gpu.device(), vc);
            auto result = Tensors::vector(n, usage).build_init<float>(gpu.device(), oc);
            for (int variant = 0; variant < 4; ++variant) {
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    gemv.dispatch_generic<float>(gpu.device(), shapes, pass, result, m, v, (GemvVariant)variant);
                }
                gpu.queue().submit(enc.finish());
                auto got = result.read(gpu.device());
                std::vector<float> ref = oc;
                orc_shape sm = oshape(n, n), sv = oshape(n);
                int ran = -1;
                orc_gemv(variant, ref.data(), &sv, mc.data(), &sm, vc.data(), &sv, &ran);
                const double e = rel_err(got, ref);
                std::printf("gpu_gemv variant %d: rel err %.3e\n", variant, e);
                failures += !(e < 1e-5);
            }
        }
        {   // gpu_op_assign
            const uint32_t n = 1757;
            std::vector<float> v0(n), v1(n);
            for (uint32_t i = 0; i < n; ++i) { v0[i] = (float)i + 0.1f; v1[i] = (float)i * 10.0f + 0.1f; }
            for (int op = 0; op < 5; ++op) {
                auto a = Tensors::vector(n, usage).build_init<float>(gpu.device(), v0);
                auto b = Tensors::vector(n, usage).build_init<float>(gpu.device(), v1);
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    OpAssign::make(gpu.device(), (OpAssignVariant)op).dispatch<float>(gpu.device(), shapes, pass, a, b);
                }
                gpu.queue().submit(enc.finish());
                auto got = a.read(gpu.device());
                std::vector<float> ref = v0;
                orc_shape s = oshape(n);
                orc_op_assign(op, ref.data(), &s, v1.data(), &s);
                const bool same = std::memcmp(got.data(), ref.data(), n * sizeof(float)) == 0;
                std::printf("gpu_op_assign op %d: %s\n", op, same ? "bit-exact" : "MISMATCH");
                failures += !same;
            }
        }
        {   // gpu_reduce
            const uint32_t n = 345;
            auto x = uniform(SEED + 3, n);
            auto vec = Tensors::vector(n, usage).build_init<float>(gpu.device(), x);
            auto res = Tensors::scalar(usage).build<float>(gpu.device());
            for (int op = 0; op < 5; ++op) {
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    Reduce::make(gpu.device(), (ReduceOp)op).dispatch<float>(gpu.device(), shapes, pass, vec, res);
                }
                gpu.queue().submit(enc.finish());
                const float got = res.read(gpu.device())[0];
                float ref = 0.f;
                orc_shape s = oshape(n);
                orc_reduce(op, x.data(), &s, &ref);
                const double e = std::fabs((double)got - ref) / std::fmax(std::fabs((double)ref), 1e-30);
                std::printf("gpu_reduce op %d: got %.8g ref %.8g rel %.3e\n", op, got, ref, e);
                failures += !(e < 1e-5);
            }
        }
        {   // gpu_prefix_sum (wgrapier prefix_sum.rs:243-288): LEN = 15071, ones / iota / pseudo-random % 10000, exact
            const uint32_t n = 15071;
            auto ps = WgPrefixSum::from_device(gpu.device());
            for (int which = 0; which < 3; ++which) {
                std::vector<uint32_t> v(n);
                uint32_t lcg = 12345u;
                for (uint32_t i = 0; i < n; ++i) {
                    lcg = lcg * 1664525u + 1013904223u;
                    v[i] = which == 0 ? 1u : which == 1 ? i : (lcg >> 8) % 10000u;
                }
                auto t = Tensors::vector(n, usage).build_init<uint32_t>(gpu.device(), v);
                auto ws = PrefixSumWorkspace::with_capacity(gpu.device(), n);
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    ps.dispatch(gpu.device(), pass, ws, t);
                }
                gpu.queue().submit(enc.finish());
                auto got = t.read(gpu.device());
                std::vector<uint32_t> ref = v, cpu = v;
                orc_prefix_sum(ref.data(), n);
                WgPrefixSum::eval_cpu(cpu);
                const bool same = got == ref && got == cpu;
                std::printf("gpu_prefix_sum input %d: %s\n", which, same ? "exact" : "MISMATCH");
                failures += !same;
            }
        }
        {   // test_sorting (wgparry radix_sort/mod.rs:238-330): 15 keys, values = 2 * key + 5, 32 bits
            auto sort = RadixSort::from_device(gpu.device());
            auto ws = RadixSortWorkspace::make(gpu.device());
            int bad = 0;
            for (uint32_t i = 0; i < 128; i += 5) {
                std::vector<uint32_t> keys = {5 + i * 4, i, 6, 123, 74657, 123, 999, (1u << 24) + 123, 6, 7, 8, 0, i * 2, 16 + i, 128 * i};
                std::vector<uint32_t> vals(keys.size());
                for (size_t k = 0; k < keys.size(); ++k) vals[k] = keys[k] * 2 + 5;
                const uint32_t n = (uint32_t)keys.size();
                auto tk = Tensors::vector(n, usage).build_init<uint32_t>(gpu.device(), keys);
                auto tv = Tensors::vector(n, usage).build_init<uint32_t>(gpu.device(), vals);
                auto ok = Tensors::vector(n, usage).build_init<uint32_t>(gpu.device(), keys);
                auto ov = Tensors::vector(n, usage).build_init<uint32_t>(gpu.device(), vals);
                auto ns = Tensors::scalar(usage).build_init<uint32_t>(gpu.device(), std::vector<uint32_t>{n});
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    sort.dispatch(gpu.device(), pass, ws, tk, tv, ns, 32, ok, ov);
                }
                gpu.queue().submit(enc.finish());
                auto gk = ok.read(gpu.device()), gv = ov.read(gpu.device());
                std::vector<uint32_t> rk = keys, rv = vals;
                orc_radix_sort(keys.data(), vals.data(), n, n, 32, rk.data(), rv.data());
                bad += !(gk == rk && gv == rv);
            }
            std::printf("gpu_radix_sort: %s\n", bad == 0 ? "exact" : "MISMATCH");
            failures += bad;
        }
        {   // wgebra::geometry test kernels (cholesky.rs:86-149, lu.rs:129-181, qr3.rs, eig4.rs:72-131, svd3.rs:70-111): LEN = 345
            // random matrices, out[i] = f(in[i]); every output word equal to the oracle's (both evaluate the WGSL's sequence
            // without contraction); NaNs (non-SDP Cholesky inputs) only need to be NaN on both sides
            const uint32_t n = 345;
            auto check_op = [&](auto shader, auto mat_tag, auto out_tag, int op, bool symmetric, const char *name) {
                using Mat = decltype(mat_tag);
                using Out = decltype(out_tag);
                constexpr int dim = sizeof(Mat) == 16 ? 2 : sizeof(Mat) == 48 ? 3 : 4;
                auto a = uniform(SEED + 20 + dim, n * dim * dim);
                std::vector<Mat> in(n);
                for (uint32_t i = 0; i < n; ++i) {
                    std::memset(&in[i], 0, sizeof(Mat));
                    for (int c = 0; c < dim; ++c)
                        for (int r = 0; r < dim; ++r) {
                            const float *m = &a[(size_t)i * dim * dim];
                            if (!symmetric) { in[i].m[c][r] = m[c * dim + r]; continue; }
                            float acc = 0.0f;                         // m^T m (cholesky.rs:98-102)
                            for (int k = 0; k < dim; ++k) acc += m[r * dim + k] * m[c * dim + k];
                            in[i].m[c][r] = acc;
                        }
                }
                auto tin = Tensors::vector(n, usage).build_init<Mat>(gpu.device(), in);
                auto tout = Tensors::vector(n, usage).build<Out>(gpu.device());
                auto enc = gpu.device().create_command_encoder();
                {
                    auto pass = enc.compute_pass("test");
                    shader.dispatch(gpu.device(), pass, tin, tout);
                }
                gpu.queue().submit(enc.finish());
                auto got = tout.read(gpu.device());
                std::vector<Out> ref(n);
                orc_geom_batch(op, dim, reinterpret_cast<const float *>(in.data()), ref.data(), n);
                const uint32_t *g = reinterpret_cast<const uint32_t *>(got.data()), *r = reinterpret_cast<const uint32_t *>(ref.data());
                size_t bad = 0;
                for (size_t k = 0; k < n * sizeof(Out) / 4; ++k) {
                    float fg, fr;
                    std::memcpy(&fg, &g[k], 4);
                    std::memcpy(&fr, &r[k], 4);
                    bad += !(g[k] == r[k] || (fg != fg && fr != fr));
                }
                std::printf("gpu_%s%d: %s\n", name, dim, bad == 0 ? "exact" : "MISMATCH");
                failures += bad != 0;
            };
            check_op(WgCholesky<2>::from_device(gpu.device()), GpuMat<2>{}, GpuMat<2>{}, 0, true, "cholesky");
            check_op(WgCholesky<3>::from_device(gpu.device()), GpuMat<3>{}, GpuMat<3>{}, 0, true, "cholesky");
            check_op(WgCholesky<4>::from_device(gpu.device()), GpuMat<4>{}, GpuMat<4>{}, 0, true, "cholesky");
            check_op(WgLU<2>::from_device(gpu.device()), GpuMat<2>{}, GpuLU<2>{}, 1, true, "lu");
            check_op(WgLU<3>::from_device(gpu.device()), GpuMat<3>{}, GpuLU<3>{}, 1, true, "lu");
            check_op(WgLU<4>::from_device(gpu.device()), GpuMat<4>{}, GpuLU<4>{}, 1, true, "lu");
            check_op(WgQR<2>::from_device(gpu.device()), GpuMat<2>{}, GpuQR<2>{}, 2, false, "qr");
            check_op(WgQR<3>::from_device(gpu.device()), GpuMat<3>{}, GpuQR<3>{}, 2, false, "qr");
            check_op(WgQR<4>::from_device(gpu.device()), GpuMat<4>{}, GpuQR<4>{}, 2, false, "qr");
            check_op(WgSymmetricEigen<2>::from_device(gpu.device()), GpuMat<2>{}, GpuSymmetricEigen<2>{}, 3, true, "eig");
            check_op(WgSymmetricEigen<3>::from_device(gpu.device()), GpuMat<3>{}, GpuSymmetricEigen<3>{}, 3, true, "eig");
            check_op(WgSymmetricEigen<4>::from_device(gpu.device()), GpuMat<4>{}, GpuSymmetricEigen<4>{}, 3, true, "eig");
            check_op(WgSvd<3>::from_device(gpu.device()), GpuMat<3>{}, GpuSvd<3>{}, 4, false, "svd");
            check_op(WgInv<4>::from_device(gpu.device()), GpuMat<4>{}, GpuMat<4>{}, 5, false, "inv");
        }
        {   // extensions through the C++ mirror: fused Gemm + OpAssign(Add) epilogue, recorded into a graph and replayed
            const uint32_t n = 256;
            auto a = uniform(SEED + 1, n, n), b = uniform(SEED + 2, n, n), e = uniform(SEED + 3, n * n);
            auto gemm = Gemm::from_device(gpu.device());
            auto ta = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), a);
            auto tb = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), b);
            auto te = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), e);
            auto out = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), std::vector<float>(n * n, 0.f));
            auto run = [&]() {
                auto enc = gpu.device().create_command_encoder();
                auto pass = enc.compute_pass("fused");
                gemm.dispatch_op<float>(gpu.device(), shapes, pass, out.as_embedded_view<3>(), ta.as_embedded_view<3>(), tb.as_embedded_view<3>(),
                                        (int)OpAssignVariant::Add, te.as_embedded_view<3>());
            };
            run();   // eager (warms the workspaces)
            Graph::capture_begin(gpu.device());
            run();
            Graph g = Graph::capture_end(gpu.device());
            g.launch();
            auto got = out.read(gpu.device());
            std::vector<float> ref(n * n, 0.f);
            orc_shape s = oshape(n, n), sv = oshape(n * n);
            orc_gemm(0, ref.data(), &s, a.data(), &s, b.data(), &s);
            orc_op_assign(0, ref.data(), &sv, e.data(), &sv);
            const double err = rel_err(got, ref);
            std::printf("fused gemm+add via graph replay: rel err %.3e\n", err);
            failures += !(err < 1e-5);
        }
        {   // fused Gemv + OpAssign(Add) with operand == out: the residual update out = m * v + out
            const uint32_t n = 1024;
            auto mc = uniform(SEED + 1, n, n), vc = uniform(SEED + 3, n), oc = uniform(SEED + 4, n);
            auto gemv = Gemv::from_device(gpu.device());
            auto tm = Tensors::matrix(n, n, usage).build_init<float>(gpu.device(), mc);
            auto tv = Tensors::vector(n, usage).build_init<float>(gpu.device(), vc);
            auto out = Tensors::vector(n, usage).build_init<float>(gpu.device(), oc);
            auto enc = gpu.device().create_command_encoder();
            {
                auto pass = enc.compute_pass("fused gemv");
                gemv.dispatch_op<float>(gpu.device(), shapes, pass, out.as_embedded_view<3>(), tm.as_embedded_view<3>(), tv.as_embedded_view<3>(),
                                        OpAssignVariant::Add, out.as_embedded_view<3>());
            }
            gpu.queue().submit(enc.finish());
            auto got = out.read(gpu.device());
            std::vector<float> ref(n, 0.f);
            orc_shape sm = oshape(n, n), sv = oshape(n);
            int ran = 0;
            orc_gemv(0, ref.data(), &sv, mc.data(), &sm, vc.data(), &sv, &ran);
            orc_op_assign(0, ref.data(), &sv, oc.data(), &sv);
            const double err = rel_err(got, ref);
            std::printf("fused gemv+add (residual update): rel err %.3e\n", err);
            failures += !(err < 1e-5);
        }
        {   // RowMajor views (tensor.rs:19-39, shape.wgsl:49-57): row-major out and m2, column-major m1, 3xTF32 path
            const uint32_t M = 192, N = 320, K = 256;
            auto a = uniform(SEED + 1, M, K), b = uniform(SEED + 2, N, K);   // b holds the K x N matrix row by row (N contiguous)
            auto gemm = Gemm::from_device(gpu.device());
            auto ta = Tensors::matrix(M, K, usage).build_init<float>(gpu.device(), a);
            auto tb = Tensors::matrix(K, N, usage).build_init<float>(gpu.device(), b);
            auto to = Tensors::matrix(M, N, usage).build_init<float>(gpu.device(), std::vector<float>((size_t)M * N, -1.f));
            auto enc = gpu.device().create_command_encoder();
            {
                auto pass = enc.compute_pass("row-major");
                gemm.dispatch<float>(gpu.device(), shapes, pass, to.as_embedded_view<3, RowMajor>(), ta.as_embedded_view<3>(),
                                     tb.as_embedded_view<3, RowMajor>());
            }
            gpu.queue().submit(enc.finish());
            auto got = to.read(gpu.device());
            std::vector<float> ref((size_t)M * N, -1.f);
            orc_shape so{M, N, 1, N, M * N, 0}, s1{M, K, 1, M, M * K, 0}, s2{K, N, 1, N, K * N, 0};
            orc_gemm_ord(0, ref.data(), &so, 1, a.data(), &s1, 0, b.data(), &s2, 1);
            const double err = rel_err(got, ref);
            std::printf("row-major out / m2 gemm: rel err %.3e\n", err);
            failures += !(err < 1e-5);
        }
        {   // the reference's panic on mismatched dimensions (gemm.rs:91)
            auto gemm = Gemm::from_device(gpu.device());
            auto a = Tensors::matrix(8, 4, usage).build<float>(gpu.device());
            auto b = Tensors::matrix(8, 8, usage).build<float>(gpu.device());
            auto o = Tensors::matrix(8, 8, usage).build<float>(gpu.device());
            bool threw = false;
            auto enc = gpu.device().create_command_encoder();
            try {
                auto pass = enc.compute_pass("test");
                gemm.dispatch<float>(gpu.device(), shapes, pass, o, a, b);
            } catch (const DimensionMismatch &e) {
                threw = std::strstr(e.what(), "Gemm: dimension mismatch") != nullptr;
            }
            std::printf("dimension mismatch -> %s\n", threw ? "DimensionMismatch" : "NOT RAISED");
            failures += !threw;
        }
    } catch (const wgb::Error &e) {
        std::printf("wgb error %d: %s\n", e.status, e.what());
        return 2;
    }
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "ALL OK", failures);
    return failures ? 1 : 0;
}
