// geometry.cuh — the small-matrix function library of wgebra::geometry as CUDA device templates.
//
// In the reference these are WGSL modules (`#define_import_path wgebra::cholesky2`, ...) that other shaders import and call
// per invocation (crates/wgebra/src/geometry/*.wgsl).  The counterpart here is a header of __device__ templates any kernel of
// this library can include and call per thread; geometry.cu wraps each one in a batched kernel (out[i] = f(in[i])) behind
// wgb_geometry_batch.  Matrices are register-resident (`m[c][r]` = column c, row r, like WGSL's m[c][r]); every loop has
// compile-time bounds and every data-dependent index of the WGSL (pivot rows, sweep windows) is turned into predicated
// static indexing, so nothing spills to local memory.
//
// The arithmetic sequence per element is the WGSL's, statement by statement (file:line cited per function); geometry.cu is
// compiled with -fmad=false so the only fused operations are the fma() calls the WGSL itself makes (svd3).
#pragma once

#include <cstdint>

namespace wgb {
namespace geom {

#define WGB_GD __device__ __forceinline__

template <int D>
struct Mat {
    float m[D][D];
};

template <int D>
struct LU {  // lu.wgsl:12-34
    Mat<D> lu;
    uint32_t ia[D], ib[D], len;
};
template <int D>
struct QR {  // qr2.wgsl:7-12
    Mat<D> q, r;
};
template <int D>
struct SymmetricEigen {  // eig2.wgsl:7-12
    Mat<D> eigenvectors;
    float eigenvalues[D];
};
template <int D>
struct Svd {  // svd2.wgsl:5-9, svd3.wgsl:12-16
    Mat<D> U;
    float S[D];
    Mat<D> Vt;
};

WGB_GD float wsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }  // WGSL sign(): 0 for 0

template <int N>
WGB_GD float dget(const float (&a)[N], unsigned i) {
    float v = a[0];
#pragma unroll
    for (int k = 1; k < N; ++k) v = (i == (unsigned)k) ? a[k] : v;
    return v;
}
template <int N, typename T>
WGB_GD void dset(T (&a)[N], unsigned i, T v) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = (i == (unsigned)k) ? v : a[k];
}

template <int D>
WGB_GD Mat<D> identity() {
    Mat<D> q;
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) q.m[c][r] = c == r ? 1.0f : 0.0f;
    return q;
}

// cholesky.wgsl:16-35 — lower-triangular factor in the lower triangle, the rest of x passes through
template <int D>
WGB_GD Mat<D> cholesky(Mat<D> x) {
#pragma unroll
    for (int j = 0; j < D; j++) {
#pragma unroll
        for (int k = 0; k < j; k++) {
            const float factor = -x.m[k][j];
#pragma unroll
            for (int l = j; l < D; l++) x.m[j][l] += factor * x.m[k][l];
        }
        const float denom = sqrtf(x.m[j][j]);
        x.m[j][j] = denom;
#pragma unroll
        for (int l = j + 1; l < D; l++) x.m[j][l] /= denom;
    }
    return x;
}

// lu.wgsl:37-132.  gauss_step_swap (:103-132) interleaves the row swap with the elimination column by column; swapping rows
// i and piv in every column first (columns < i: :64-68) and then running gauss_step (:85-99) performs the same operations
// on the same values.
template <int D>
WGB_GD LU<D> lu(Mat<D> x) {
    LU<D> out;
#pragma unroll
    for (int k = 0; k < D; ++k) out.ia[k] = out.ib[k] = 0u;
    out.len = 0u;
#pragma unroll
    for (int i = 0; i < D; i++) {
        int piv = i;
        float piv_val = fabsf(x.m[i][i]);
#pragma unroll
        for (int r = i + 1; r < D; r++) {
            const float abs_val = fabsf(x.m[i][r]);
            if (abs_val > piv_val) {
                piv = r;
                piv_val = abs_val;
            }
        }
        if (piv_val != 0.0f) {
            if (piv != i) {
                dset(out.ia, out.len, (uint32_t)i);
                dset(out.ib, out.len, (uint32_t)piv);
                out.len++;
#pragma unroll
                for (int c = 0; c < D; ++c)
#pragma unroll
                    for (int r = i + 1; r < D; ++r)
                        if (piv == r) {
                            const float t = x.m[c][i];
                            x.m[c][i] = x.m[c][r];
                            x.m[c][r] = t;
                        }
            }
            const float inv_diag = 1.0f / x.m[i][i];
#pragma unroll
            for (int r = i + 1; r < D; r++) x.m[i][r] *= inv_diag;
#pragma unroll
            for (int c = i + 1; c < D; c++) {
                const float pivot = x.m[c][i];
#pragma unroll
                for (int r = i + 1; r < D; r++) x.m[c][r] -= pivot * x.m[i][r];
            }
        }
    }
    out.lu = x;
    return out;
}

// Householder axis of column i from row R0 on (householder::reflection_axis_mut as ported in qr2.wgsl:22-52 and
// eig3.wgsl:217-248): normalises m[i][R0..] in place, returns the signed norm for the (off-)diagonal; *applied = factor != 0.
template <int D>
WGB_GD float reflection_axis(Mat<D> &x, int i, int r0, bool *applied) {
    float axis_sq_norm = 0.0f;
#pragma unroll
    for (int r = 0; r < D; r++)
        if (r >= r0) axis_sq_norm += x.m[i][r] * x.m[i][r];
    const float axis_norm = sqrtf(axis_sq_norm);
    const float head = dget(x.m[i], (unsigned)r0);
    const float modulus = fabsf(head);
    const float sgn = wsign(head);
    const float signed_norm = sgn * axis_norm;
    const float factor = (axis_sq_norm + modulus * axis_norm) * 2.0f;
    dset(x.m[i], (unsigned)r0, head + signed_norm);
    *applied = factor != 0.0f;
    if (factor != 0.0f) {
        const float factor_sqrt = sqrtf(factor);
        float norm = 0.0f;
#pragma unroll
        for (int r = 0; r < D; r++)
            if (r >= r0) {
                x.m[i][r] /= factor_sqrt;
                norm += x.m[i][r] * x.m[i][r];
            }
        norm = sqrtf(norm);
#pragma unroll
        for (int r = 0; r < D; r++)
            if (r >= r0) x.m[i][r] /= norm;
        return -signed_norm;
    }
    return signed_norm;
}

// refl.reflect_with_sign on columns c >= i, rows >= r0, with the axis held in column i of `axis` (qr2.wgsl:58-69, :83-92)
template <int D>
WGB_GD void reflect_with_sign(Mat<D> &target, const Mat<D> &axis, int i, int r0, float sgn) {
#pragma unroll
    for (int c = 0; c < D; c++)
        if (c >= i) {
            const float m_two = -2.0f * sgn;
            float factor = 0.0f;
#pragma unroll
            for (int r = 0; r < D; r++)
                if (r >= r0) factor += axis.m[i][r] * target.m[c][r];
#pragma unroll
            for (int r = 0; r < D; r++)
                if (r >= r0) target.m[c][r] = m_two * factor * axis.m[i][r] + target.m[c][r] * sgn;
        }
}

// qr2.wgsl:15-107 = qr3.wgsl:15-109 = qr4.wgsl:15-111
template <int D>
WGB_GD QR<D> qr(Mat<D> x) {
    float diag[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
        bool applied;
        diag[i] = reflection_axis(x, i, i, &applied);
        if (applied) {
            // the WGSL reflects column i itself with its own (already normalised) axis as well: the in-place update of
            // m[i][r] while it is read as the axis is reproduced by reflecting columns in ascending order on one matrix
            const float sgn = wsign(diag[i]);
#pragma unroll
            for (int c = i; c < D; c++) {
                const float m_two = -2.0f * sgn;
                float factor = 0.0f;
#pragma unroll
                for (int r = i; r < D; r++) factor += x.m[i][r] * x.m[c][r];
#pragma unroll
                for (int r = i; r < D; r++) x.m[c][r] = m_two * factor * x.m[i][r] + x.m[c][r] * sgn;
            }
        }
    }
    QR<D> out;
    out.q = identity<D>();
#pragma unroll
    for (int i = D - 1; i >= 0; i--) reflect_with_sign(out.q, x, i, i, wsign(diag[i]));
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) out.r.m[c][r] = r < c ? x.m[c][r] : (r == c ? fabsf(diag[c]) : 0.0f);
    return out;
}

// eig2.wgsl:42-56
WGB_GD void eigenvalues2(float a, float c, float b, float &e0, float &e1) {
    if (c == 0.0f) {
        e0 = a;
        e1 = b;
        return;
    }
    const float ab = a - b;
    const float sigma = sqrtf(4.0f * c * c + ab * ab);
    e0 = (a + b + sigma) / 2.0f;
    e1 = (a + b - sigma) / 2.0f;
}

// eig2.wgsl:15-40
WGB_GD SymmetricEigen<2> symmetric_eigen(Mat<2> x) {
    SymmetricEigen<2> out;
    const float a = x.m[0][0], c = x.m[0][1], b = x.m[1][1];
    if (c == 0.0f) {
        out.eigenvectors = identity<2>();
        out.eigenvalues[0] = a;
        out.eigenvalues[1] = b;
        return out;
    }
    const float ab = a - b;
    const float sigma = sqrtf(4.0f * c * c + ab * ab);
    out.eigenvalues[0] = (a + b + sigma) / 2.0f;
    out.eigenvalues[1] = (a + b - sigma) / 2.0f;
    const float e1x = (a - b + sigma) / (2.0f * c), e2x = (a - b - sigma) / (2.0f * c);
    const float l1 = sqrtf(e1x * e1x + 1.0f * 1.0f), l2 = sqrtf(e2x * e2x + 1.0f * 1.0f);
    out.eigenvectors.m[0][0] = e1x / l1;
    out.eigenvectors.m[0][1] = 1.0f / l1;
    out.eigenvectors.m[1][0] = e2x / l2;
    out.eigenvectors.m[1][1] = 1.0f / l2;
    return out;
}

// eig3.wgsl:162-197 (off_diag has D - 1 entries)
template <int D>
WGB_GD void delimit_subproblem(const float (&diag)[D], float (&off_diag)[D - 1], unsigned end, float eps, unsigned &start_out,
                               unsigned &end_out) {
    unsigned n = end;
    bool stop = false;
#pragma unroll
    for (int k = D - 1; k >= 1; --k)
        if (!stop && n == (unsigned)k) {
            if (fabsf(off_diag[k - 1]) > eps * (fabsf(diag[k]) + fabsf(diag[k - 1]))) stop = true;
            else n = (unsigned)(k - 1);
        }
    if (n == 0u) {
        start_out = 0u;
        end_out = 0u;
        return;
    }
    unsigned new_start = n - 1u;
    stop = false;
#pragma unroll
    for (int k = D - 2; k >= 1; --k)
        if (!stop && new_start == (unsigned)k) {
            if (off_diag[k - 1] == 0.0f || fabsf(off_diag[k - 1]) <= eps * (fabsf(diag[k]) + fabsf(diag[k - 1]))) {
                off_diag[k - 1] = 0.0f;
                stop = true;
            } else {
                new_start = (unsigned)(k - 1);
            }
        }
    start_out = new_start;
    end_out = n;
}

// eig3.wgsl:199-209
WGB_GD float wilkinson_shift(float tmm, float tnn, float tmn) {
    const float sq_tmn = tmn * tmn;
    if (sq_tmn != 0.0f) {
        const float d = (tmm - tnn) * 0.5f;
        return tnn - sq_tmn / (d + wsign(d) * sqrtf(d * d + sq_tmn));
    }
    return tnn;
}

// rot2.wgsl:75-94 with invMulVec :70-72: rotate "rows" (WGSL columns) i and i + 1
template <int D>
WGB_GD void rotate_rows(float rc, float rs, Mat<D> &q, int i) {
#pragma unroll
    for (int r = 0; r < D; r++) {
        const float vx = q.m[i][r], vy = q.m[i + 1][r];
        q.m[i][r] = rc * vx + rs * vy;
        q.m[i + 1][r] = -rs * vx + rc * vy;
    }
}

constexpr int kEigMaxSweeps = 256;  // the reference loop is unbounded (eig3.wgsl:77); same bound as the oracle

// eig3.wgsl:24-160 = eig4.wgsl:24-162, tridiagonalize eig3.wgsl:211-282
template <int D>
WGB_GD SymmetricEigen<D> symmetric_eigen(Mat<D> x) {
    static_assert(D == 3 || D == 4, "eig2 is the closed form above");
    const float EPS = 1.1920929e-7f;
    float m_amax;
    {  // min_max.wgsl:27-30 / :44-47
        float vm[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            vm[r] = fabsf(x.m[0][r]);
#pragma unroll
            for (int c = 1; c < D; ++c) vm[r] = fmaxf(vm[r], fabsf(x.m[c][r]));
        }
        m_amax = vm[0];
#pragma unroll
        for (int r = 1; r < D; ++r) m_amax = fmaxf(m_amax, vm[r]);
    }
    if (m_amax != 0.0f) {
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
            for (int r = 0; r < D; ++r) x.m[c][r] /= m_amax;
    }
    // tridiagonalize
    float tri_off[D - 1];
#pragma unroll
    for (int i = 0; i < D - 1; i++) {
        bool applied;
        tri_off[i] = reflection_axis(x, i, i + 1, &applied);
        if (applied) {
            float p[D];
#pragma unroll
            for (int r = 0; r < D; ++r) p[r] = 0.0f;
#pragma unroll
            for (int c = i + 1; c < D; c++)
#pragma unroll
                for (int r = i + 1; r < D; r++) p[r] += 2.0f * x.m[c][r] * x.m[i][c];
            float dot = 0.0f;
#pragma unroll
            for (int r = i + 1; r < D; r++) dot += x.m[i][r] * p[r];
#pragma unroll
            for (int c = i + 1; c < D; c++)
#pragma unroll
                for (int r = i + 1; r < D; r++)
                    x.m[c][r] += 2.0f * dot * x.m[i][r] * x.m[i][c] - p[r] * x.m[i][c] - x.m[i][r] * p[c];
        }
    }
    float diag[D], off_diag[D - 1];
#pragma unroll
    for (int i = 0; i < D; ++i) diag[i] = x.m[i][i];
#pragma unroll
    for (int i = 0; i < D - 1; ++i) off_diag[i] = fabsf(tri_off[i]);
    Mat<D> q = identity<D>();
#pragma unroll
    for (int i = D - 2; i >= 0; i--) reflect_with_sign(q, x, i, i + 1, wsign(tri_off[i]));

    unsigned start, end;
    delimit_subproblem<D>(diag, off_diag, D - 1, EPS, start, end);
    int niter = 0;
    while (end != start && niter < kEigMaxSweeps) {
        const unsigned subdim = end - start + 1u;
        if (subdim > 2u) {
            const unsigned mm = end - 1u, n = end;
            const float shift = wilkinson_shift(dget(diag, mm), dget(diag, n), dget(off_diag, mm));
            float vx = dget(diag, start) - shift, vy = dget(off_diag, start);
            bool live = true;
#pragma unroll
            for (int i = 0; i < D - 1; i++) {
                if (live && (unsigned)i >= start && (unsigned)i < n) {
                    float rc = 0.0f, rs = 0.0f;  // rot2.wgsl:29-37 cancel_y
                    if (vy != 0.0f) {
                        const float r = wsign(vx) / sqrtf(vx * vx + vy * vy);
                        rc = vx * r;
                        rs = -vy * r;
                    }
                    if (rc != 0.0f || rs != 0.0f) {  // rot2.wgsl:15-17 is_valid
                        if (i > 0 && (unsigned)i > start) off_diag[i > 0 ? i - 1 : 0] = wsign(vx) * sqrtf(vx * vx + vy * vy);
                        const float mii = diag[i], mjj = diag[i + 1], mij = off_diag[i];
                        const float cc = rc * rc, ss = rs * rs, cs = rc * rs;
                        const float b = cs * 2.0f * mij;
                        diag[i] = (cc * mii + ss * mjj) - b;
                        diag[i + 1] = (ss * mii + cc * mjj) + b;
                        off_diag[i] = cs * (mii - mjj) + mij * (cc - ss);
                        if (i + 1 < D - 1 && (unsigned)i != n - 1u) {
                            constexpr int kLast = D - 2;
                            const int i1 = i + 1 < D - 1 ? i + 1 : kLast;
                            vx = off_diag[i];
                            vy = -rs * off_diag[i1];
                            off_diag[i1] *= rc;
                        }
                        rotate_rows(rc, -rs, q, i);  // Rot::inv(rot), rot2.wgsl:53-55
                    } else {
                        live = false;
                    }
                }
            }
            if (fabsf(dget(off_diag, mm)) <= EPS * (fabsf(dget(diag, mm)) + fabsf(dget(diag, n)))) end -= 1u;
        } else if (subdim == 2u) {
#pragma unroll
            for (int s = 0; s < D - 1; ++s)
                if (start == (unsigned)s) {
                    float e0, e1;
                    eigenvalues2(diag[s], off_diag[s], diag[s + 1], e0, e1);
                    const float bx = e0 - diag[s + 1], by = off_diag[s];
                    diag[s] = e0;
                    diag[s + 1] = e1;
                    const float basis_len = sqrtf(bx * bx + by * by);
                    if (basis_len > EPS) {
                        const float sc = wsign(bx) / basis_len;
                        rotate_rows(bx * sc, by * sc, q, s);
                    }
                }
            end -= 1u;
        }
        delimit_subproblem<D>(diag, off_diag, end, EPS, start, end);
        niter++;
    }
    SymmetricEigen<D> out;
    out.eigenvectors = q;
#pragma unroll
    for (int i = 0; i < D; ++i) out.eigenvalues[i] = diag[i] * m_amax;
    return out;
}

// trig.wgsl:26-41
WGB_GD float stable_atan2(float y, float x) {
    const float PI = 3.14159265358979323846264338327950288f;
    const float ang = atanf(y / x);
    if (x > 0.0f) return ang;
    if (x < 0.0f && y > 0.0f) return ang + PI;
    if (x < 0.0f && y < 0.0f) return ang - PI;
    return 0.0f;
}

// svd2.wgsl:12-39
WGB_GD Svd<2> svd(Mat<2> x) {
    const float e = (x.m[0][0] + x.m[1][1]) * 0.5f, f = (x.m[0][0] - x.m[1][1]) * 0.5f;
    const float g = (x.m[0][1] + x.m[1][0]) * 0.5f, h = (x.m[0][1] - x.m[1][0]) * 0.5f;
    const float q = sqrtf(e * e + h * h), r = sqrtf(f * f + g * g);
    const float sx = q + r, sy = q - r;
    const float sy_sign = sy < 0.0f ? -1.0f : 1.0f;
    const float a1 = stable_atan2(g, f), a2 = stable_atan2(h, e);
    const float theta = (a2 - a1) * 0.5f, phi = (a2 + a1) * 0.5f;
    const float st = sinf(theta), ct = cosf(theta), sp = sinf(phi), cp = cosf(phi);
    Svd<2> out;
    out.S[0] = sx;
    out.S[1] = sy * sy_sign;
    out.U.m[0][0] = cp;
    out.U.m[0][1] = sp;
    out.U.m[1][0] = -sp;
    out.U.m[1][1] = cp;
    out.Vt.m[0][0] = ct;
    out.Vt.m[0][1] = st * sy_sign;
    out.Vt.m[1][0] = -st;
    out.Vt.m[1][1] = ct * sy_sign;
    return out;
}

// svd3.wgsl:55-80: bit-trick seed + Newton steps, written with fma so CPU and GPU agree bit for bit
template <int STEPS>
WGB_GD float rsqrt_newton(float val) {
    float x = val;
    const float xhalf = -0.5f * x;
    int i = __float_as_int(x);
    i = 0x5f375a82 - (i >> 1);
    x = __int_as_float(i);
#pragma unroll
    for (int k = 0; k < STEPS; k++) x = x * fmaf(x * x, xhalf, 1.5f);
    return x;
}

struct Sym3 {  // svd3.wgsl:37-46
    float mxx, myx, myy, mzx, mzy, mzz;
};

// svd3.wgsl:116-129
WGB_GD void approximate_givens_quaternion(const Sym3 &A, float &ch, float &sh) {
    const float gch = 2.0f * (A.mxx - A.myy), gsh = A.myx;
    bool b = 5.828427124f * gsh * gsh < gch * gch;
    const float w = rsqrt_newton<4>(fmaf(gch, gch, gsh * gsh));
    if (w != w) b = false;
    ch = b ? w * gch : 0.923879532f;
    sh = b ? w * gsh : 0.3826834323f;
}

// svd3.wgsl:132-167; (X, Y, Z) are compile-time here
template <int X, int Y, int Z>
WGB_GD void jacobi_conjugation(Sym3 &S, float (&q)[4]) {
    float gch, gsh;
    approximate_givens_quaternion(S, gch, gsh);
    const float scale = 1.0f / fmaf(gch, gch, gsh * gsh);
    const float a = fmaf(gch, gch, -gsh * gsh) * scale;
    const float b = 2.0f * gsh * gch * scale;
    const Sym3 T = S;
    Sym3 N;
    N.mxx = fmaf(a, fmaf(a, T.mxx, b * T.myx), b * (fmaf(a, T.myx, b * T.myy)));
    N.myx = fmaf(a, fmaf(-b, T.mxx, a * T.myx), b * (fmaf(-b, T.myx, a * T.myy)));
    N.myy = fmaf(-b, fmaf(-b, T.mxx, a * T.myx), a * (fmaf(-b, T.myx, a * T.myy)));
    N.mzx = fmaf(a, T.mzx, b * T.mzy);
    N.mzy = fmaf(-b, T.mzx, a * T.mzy);
    N.mzz = T.mzz;
    const float tmp[3] = {q[0] * gsh, q[1] * gsh, q[2] * gsh};
    gsh *= q[3];
    q[Z] = fmaf(q[Z], gch, gsh);
    q[3] = fmaf(q[3], gch, -tmp[Z]);
    q[X] = fmaf(q[X], gch, tmp[Y]);
    q[Y] = fmaf(q[Y], gch, -tmp[X]);
    S.mxx = N.myy;
    S.myx = N.mzy;
    S.myy = N.mzz;
    S.mzx = N.myx;
    S.mzy = N.mzx;
    S.mzz = N.mxx;
}

// svd3.wgsl:215-229
WGB_GD void qr_givens_quaternion(float a1, float a2, float &ch_out, float &sh_out) {
    const float epsilon = 1e-6f;
    const float rho = 1.0f / rsqrt_newton<6>(fmaf(a1, a1, a2 * a2));
    float ch = fabsf(a1) + fmaxf(rho, epsilon);
    float sh = rho > epsilon ? a2 : 0.0f;
    if (a1 < 0.0f) {
        const float t = sh;
        sh = ch;
        ch = t;
    }
    const float w = rsqrt_newton<4>(fmaf(ch, ch, sh * sh));
    ch_out = ch * w;
    sh_out = sh * w;
}

// WGSL `A * B` for mat3x3: out[c][r] = sum_k A[k][r] * B[c][k], k ascending
WGB_GD Mat<3> mul(const Mat<3> &A, const Mat<3> &B) {
    Mat<3> o;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) o.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
    return o;
}
WGB_GD Mat<3> transpose(const Mat<3> &A) {
    Mat<3> o;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) o.m[c][r] = A.m[r][c];
    return o;
}

// quat.wgsl:31-53
WGB_GD Mat<3> quat_to_matrix(const float (&q)[4]) {
    const float i = q[0], j = q[1], k = q[2], w = q[3];
    const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    const float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f, ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
    Mat<3> V;
    V.m[0][0] = ww + ii - jj - kk;
    V.m[0][1] = wk + ij;
    V.m[0][2] = ik - wj;
    V.m[1][0] = ij - wk;
    V.m[1][1] = ww - ii + jj - kk;
    V.m[1][2] = wi + jk;
    V.m[2][0] = wj + ik;
    V.m[2][1] = jk - wi;
    V.m[2][2] = ww - ii - jj + kk;
    return V;
}

WGB_GD void cond_neg_swap3(bool c, float (&x)[3], float (&y)[3]) {  // svd3.wgsl:108-112
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x0 = -x[k];
        x[k] = c ? y[k] : x[k];
        y[k] = c ? x0 : y[k];
    }
}

// svd3.wgsl:291-305 (jacobiEigenanalysis :171-180, sortSingularValues :190-212, QRDecomposition :232-288)
WGB_GD Svd<3> svd(const Mat<3> &A) {
    const Mat<3> ata = mul(transpose(A), A);
    Sym3 s = {ata.m[0][0], ata.m[0][1], ata.m[1][1], ata.m[0][2], ata.m[1][2], ata.m[2][2]};
    float qv[4] = {0.0f, 0.0f, 0.0f, 1.0f};
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        jacobi_conjugation<0, 1, 2>(s, qv);
        jacobi_conjugation<1, 2, 0>(s, qv);
        jacobi_conjugation<2, 0, 1>(s, qv);
    }
    Mat<3> V = quat_to_matrix(qv);
    Mat<3> B = mul(A, V);
    {
        float rho1 = B.m[0][0] * B.m[0][0] + B.m[0][1] * B.m[0][1] + B.m[0][2] * B.m[0][2];
        float rho2 = B.m[1][0] * B.m[1][0] + B.m[1][1] * B.m[1][1] + B.m[1][2] * B.m[1][2];
        float rho3 = B.m[2][0] * B.m[2][0] + B.m[2][1] * B.m[2][1] + B.m[2][2] * B.m[2][2];
        bool c = rho1 < rho2;
        cond_neg_swap3(c, B.m[0], B.m[1]);
        cond_neg_swap3(c, V.m[0], V.m[1]);
        {
            const float t = rho1;
            rho1 = c ? rho2 : rho1;
            rho2 = c ? t : rho2;
        }
        c = rho1 < rho3;
        cond_neg_swap3(c, B.m[0], B.m[2]);
        cond_neg_swap3(c, V.m[0], V.m[2]);
        {
            const float t = rho1;
            rho1 = c ? rho3 : rho1;
            rho3 = c ? t : rho3;
        }
        c = rho2 < rho3;
        cond_neg_swap3(c, B.m[1], B.m[2]);
        cond_neg_swap3(c, V.m[1], V.m[2]);
    }
    float g1c, g1s, g2c, g2s, g3c, g3s;
    qr_givens_quaternion(B.m[0][0], B.m[0][1], g1c, g1s);
    float a = fmaf(-2.0f, g1s * g1s, 1.0f), b = 2.0f * g1c * g1s;
    float r00 = fmaf(a, B.m[0][0], b * B.m[0][1]), r01 = fmaf(a, B.m[1][0], b * B.m[1][1]), r02 = fmaf(a, B.m[2][0], b * B.m[2][1]);
    float r10 = fmaf(-b, B.m[0][0], a * B.m[0][1]), r11 = fmaf(-b, B.m[1][0], a * B.m[1][1]), r12 = fmaf(-b, B.m[2][0], a * B.m[2][1]);
    const float r20 = B.m[0][2], r21 = B.m[1][2], r22 = B.m[2][2];
    qr_givens_quaternion(r00, r20, g2c, g2s);
    a = fmaf(-2.0f, g2s * g2s, 1.0f);
    b = 2.0f * g2c * g2s;
    const float b00 = fmaf(a, r00, b * r20);
    const float b10 = r10, b11 = r11, b12 = r12;
    const float b20 = fmaf(-b, r00, a * r20), b21 = fmaf(-b, r01, a * r21), b22 = fmaf(-b, r02, a * r22);
    qr_givens_quaternion(b11, b21, g3c, g3s);
    a = fmaf(-2.0f, g3s * g3s, 1.0f);
    b = 2.0f * g3c * g3s;
    (void)b10;
    (void)b20;
    const float s0 = b00;
    const float s1 = fmaf(a, b11, b * b21);
    const float s2 = fmaf(-b, b12, a * b22);
    const float sh12 = 2.0f * fmaf(g1s, g1s, -0.5f), sh22 = 2.0f * fmaf(g2s, g2s, -0.5f), sh32 = 2.0f * fmaf(g3s, g3s, -0.5f);
    Svd<3> out;
    out.U.m[0][0] = sh12 * sh22;
    out.U.m[1][0] = fmaf(4.0f * g2c * g3c, sh12 * g2s * g3s, 2.0f * g1c * g1s * sh32);
    out.U.m[2][0] = fmaf(4.0f * g1c * g3c, g1s * g3s, -2.0f * g2c * sh12 * g2s * sh32);
    out.U.m[0][1] = -2.0f * g1c * g1s * sh22;
    out.U.m[1][1] = fmaf(-8.0f * g1c * g2c * g3c, g1s * g2s * g3s, sh12 * sh32);
    out.U.m[2][1] = fmaf(-2.0f * g3c, g3s, 4.0f * g1s * fmaf(g3c * g1s, g3s, g1c * g2c * g2s * sh32));
    out.U.m[0][2] = 2.0f * g2c * g2s;
    out.U.m[1][2] = -2.0f * g3c * sh22 * g3s;
    out.U.m[2][2] = sh22 * sh32;
    out.S[0] = s0;
    out.S[1] = s1;
    out.S[2] = s2;
    out.Vt = transpose(V);
    return out;
}

// inv.wgsl:8-20
WGB_GD Mat<2> inverse(const Mat<2> &m) {
    const float det = m.m[0][0] * m.m[1][1] - m.m[1][0] * m.m[0][1];
    const float s = 1.0f / det;
    Mat<2> o;
    o.m[0][0] = m.m[1][1] * s;
    o.m[0][1] = -m.m[0][1] * s;
    o.m[1][0] = -m.m[1][0] * s;
    o.m[1][1] = m.m[0][0] * s;
    return o;
}
// inv.wgsl:26-45
WGB_GD Mat<3> inverse(const Mat<3> &x) {
    const float(&m)[3][3] = x.m;
    Mat<3> adj;
    adj.m[0][0] = (m[1][1] * m[2][2] - m[2][1] * m[1][2]);
    adj.m[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]);
    adj.m[2][0] = (m[1][0] * m[2][1] - m[2][0] * m[1][1]);
    adj.m[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]);
    adj.m[1][1] = (m[0][0] * m[2][2] - m[2][0] * m[0][2]);
    adj.m[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]);
    adj.m[0][2] = (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    adj.m[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]);
    adj.m[2][2] = (m[0][0] * m[1][1] - m[1][0] * m[0][1]);
    const float det = (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]));
    const float s = 1.0f / det;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) adj.m[c][r] *= s;
    return adj;
}
// inv.wgsl:52-88
WGB_GD Mat<4> inverse(const Mat<4> &x) {
    const float(&m)[4][4] = x.m;
    const float sf00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    const float sf01 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const float sf02 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    const float sf03 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const float sf04 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    const float sf05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const float sf06 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    const float sf07 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float sf08 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    const float sf09 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    const float sf10 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    const float sf11 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float sf12 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    const float sf13 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float sf14 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float sf15 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float sf16 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float sf17 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float sf18 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    Mat<4> adj;
    adj.m[0][0] = (m[1][1] * sf00 - m[1][2] * sf01 + m[1][3] * sf02);
    adj.m[1][0] = -(m[1][0] * sf00 - m[1][2] * sf03 + m[1][3] * sf04);
    adj.m[2][0] = (m[1][0] * sf01 - m[1][1] * sf03 + m[1][3] * sf05);
    adj.m[3][0] = -(m[1][0] * sf02 - m[1][1] * sf04 + m[1][2] * sf05);
    adj.m[0][1] = -(m[0][1] * sf00 - m[0][2] * sf01 + m[0][3] * sf02);
    adj.m[1][1] = (m[0][0] * sf00 - m[0][2] * sf03 + m[0][3] * sf04);
    adj.m[2][1] = -(m[0][0] * sf01 - m[0][1] * sf03 + m[0][3] * sf05);
    adj.m[3][1] = (m[0][0] * sf02 - m[0][1] * sf04 + m[0][2] * sf05);
    adj.m[0][2] = (m[0][1] * sf06 - m[0][2] * sf07 + m[0][3] * sf08);
    adj.m[1][2] = -(m[0][0] * sf06 - m[0][2] * sf09 + m[0][3] * sf10);
    adj.m[2][2] = (m[0][0] * sf11 - m[0][1] * sf09 + m[0][3] * sf12);
    adj.m[3][2] = -(m[0][0] * sf08 - m[0][1] * sf10 + m[0][2] * sf12);
    adj.m[0][3] = -(m[0][1] * sf13 - m[0][2] * sf14 + m[0][3] * sf15);
    adj.m[1][3] = (m[0][0] * sf13 - m[0][2] * sf16 + m[0][3] * sf17);
    adj.m[2][3] = -(m[0][0] * sf14 - m[0][1] * sf16 + m[0][3] * sf18);
    adj.m[3][3] = (m[0][0] * sf15 - m[0][1] * sf17 + m[0][2] * sf18);
    const float det = (m[0][0] * adj.m[0][0] + m[0][1] * adj.m[1][0] + m[0][2] * adj.m[2][0] + m[0][3] * adj.m[3][0]);
    const float s = 1.0f / det;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) adj.m[c][r] *= s;
    return adj;
}

#undef WGB_GD

}  // namespace geom
}  // namespace wgb
