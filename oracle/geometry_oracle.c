/*
 * geometry_oracle.c — CPU restatement of the small-matrix factorization libraries of wgebra::geometry:
 * Cholesky, LU with partial pivoting, Householder QR, symmetric eigendecomposition (2 / 3 / 4), SVD (2 / 3) and the closed-form
 * inverses, each applied to a batch `out[i] = f(in[i])` exactly like the reference's test kernels do.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as wgsl_oracle.c: only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load it).
 *
 * What it restates (paths relative to /root/reference/crates/wgebra/src/):
 *   geometry/cholesky.wgsl:16-35       cholesky (DIM = 2, 3, 4 by textual substitution, cholesky.rs:3-19)
 *   geometry/lu.wgsl:37-132            lu, gauss_step, gauss_step_swap (+ Permutations / LU structs :12-34)
 *   geometry/qr2.wgsl:15-107, qr3.wgsl:15-109, qr4.wgsl:15-111    qr (identical text up to DIM)
 *   geometry/eig2.wgsl:15-56           symmetric_eigen / eigenvalues of a 2x2
 *   geometry/eig3.wgsl:24-282, eig4.wgsl:24-284    tridiagonalize, implicit-shift QR sweeps, delimit_subproblem, wilkinson_shift
 *   geometry/rot2.wgsl:29-94           cancel_y, is_valid, inv, invMulVec, rotate_rows3 / rotate_rows4
 *   geometry/svd2.wgsl:12-46           svd of a 2x2 (+ utils/trig.wgsl:26-41 stable_atan2)
 *   geometry/svd3.wgsl:55-312          McAdams et al. 3x3 SVD (rsqrt Newton steps, Jacobi eigenanalysis, sort, Givens QR)
 *   geometry/quat.wgsl:10-53           identity, toMatrix
 *   geometry/inv.wgsl:8-88             inv2, inv3, inv4
 *   utils/min_max.wgsl:27-50           amax3x3, amax4x4
 *
 * Storage layout = WGSL's for array<MAT> / array<struct> (what the reference's Rust side uploads and reads back, e.g. Matrix4x3
 * for a mat3x3, cholesky.rs:151-153, svd3.rs:12-23): a matrix is DIM columns of CS floats, CS = 2 for DIM 2 and 4 for DIM 3 / 4
 * (the 4th float of a vec3 column is padding).  Struct outputs in 4-byte words:
 *   LU   {lu: MAT, ia: PERM, ib: PERM, len: u32}  DIM 2: 4+2+2+1(+1 pad) = 10 | DIM 3: 12+4+3+1 = 20 | DIM 4: 16+4+4+1(+3) = 28
 *   QR   {q: MAT, r: MAT}                         8 | 24 | 32
 *   Eig  {eigenvectors: MAT, eigenvalues: VEC}    6 | 16 | 20
 *   Svd  {U: MAT, S: VEC, Vt: MAT}                10 | 28
 * Padding words are written as zero (WGSL leaves them undefined).
 *
 * Evaluation order: left to right as written in the WGSL, no FMA contraction (-ffp-contract=off) except where the WGSL itself
 * calls fma() (svd3).  WGSL leaves contraction and matrix-product summation order to the backend, so — as for the linalg path —
 * bit parity with "the" reference is undefined; the contract is the reference tests' own tolerance (reconstruction to 1e-4,
 * factors to 1e-3 / 1e-4 against nalgebra).  The CUDA kernels are compiled without contraction too and agree with this file
 * bit for bit except through sin / cos / atan (svd2).
 *
 * One deliberate superset: the QR-sweep loop of eig3 / eig4 (`while end != start`, eig3.wgsl:77) has no iteration bound in the
 * reference; here and in the CUDA kernels it stops after ORC_EIG_MAX_SWEEPS sweeps.
 *
 * PARITY PINNING: the reference holds no golden vectors for these functions; its tests (cholesky.rs:86-149, lu.rs:129-181,
 * qr2.rs:70-121, eig2.rs:69-114, eig3.rs:71-131, eig4.rs:72-131, svd2.rs:68-107, svd3.rs:70-111) compare with nalgebra on
 * unseeded random batches of 345.  tests/test_geometry_oracle.py replays them against this file with numpy float64 in
 * nalgebra's place.  Beyond that it is pinned to outputs of the reference's own shader text: the unmodified .wgsl libraries,
 * composed with the test kernels embedded in those Rust tests, are executed by tests/golden/wgsl_interp.py
 * (tests/golden/make_reference_vectors.py -> ref_wgsl_geometry.npz) and tests/test_reference_vectors.py requires this file to
 * reproduce all 17 families word for word (factors, permutations, padding).  Not pinnable here: what WGSL leaves to the
 * backend (contraction, accuracy of sin / cos / atan) — there is no wgpu runtime in this image.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { ORC_OK = 0, ORC_UNSUPPORTED = 4 };
enum { GEOM_CHOLESKY = 0, GEOM_LU = 1, GEOM_QR = 2, GEOM_EIG = 3, GEOM_SVD = 4, GEOM_INV = 5 };
#define ORC_EIG_MAX_SWEEPS 256

typedef float mat[4][4]; /* m[c][r]: column c, row r — WGSL's m[c][r] */

static inline float wsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); } /* WGSL sign() */
static inline int col_stride(int dim) { return dim == 2 ? 2 : 4; }
static inline int mat_words(int dim) { return dim * col_stride(dim); }

static void load_mat(mat m, const float *src, int dim) {
    const int cs = col_stride(dim);
    memset(m, 0, sizeof(mat));
    for (int c = 0; c < dim; ++c)
        for (int r = 0; r < dim; ++r) m[c][r] = src[c * cs + r];
}
static void store_mat(float *dst, mat m, int dim) {
    const int cs = col_stride(dim);
    for (int c = 0; c < dim; ++c)
        for (int r = 0; r < cs; ++r) dst[c * cs + r] = r < dim ? m[c][r] : 0.0f;
}
static void store_vec(float *dst, const float *v, int dim) {
    const int cs = col_stride(dim);
    for (int r = 0; r < cs; ++r) dst[r] = r < dim ? v[r] : 0.0f;
}

/* cholesky.wgsl:16-35 */
static void cholesky(mat m, int DIM) {
    for (int j = 0; j < DIM; j++) {
        for (int k = 0; k < j; k++) {
            const float factor = -m[k][j];
            for (int l = j; l < DIM; l++) m[j][l] += factor * m[k][l];
        }
        const float denom = sqrtf(m[j][j]);
        m[j][j] = denom;
        for (int l = j + 1; l < DIM; l++) m[j][l] /= denom;
    }
}

/* lu.wgsl:85-99 */
static void gauss_step(mat m, float diag, int i, int N) {
    const float inv_diag = 1.0f / diag;
    for (int r = i + 1; r < N; r++) m[i][r] *= inv_diag;
    for (int c = i + 1; c < N; c++) {
        const float pivot = m[c][i];
        for (int r = i + 1; r < N; r++) m[c][r] -= pivot * m[i][r];
    }
}
/* lu.wgsl:103-132 */
static void gauss_step_swap(mat m, float diag, int i, int piv, int N) {
    const float inv_diag = 1.0f / diag;
    const float mii = m[i][i];
    m[i][i] = m[i][piv];
    m[i][piv] = mii;
    for (int r = i + 1; r < N; r++) m[i][r] *= inv_diag;
    for (int c = i + 1; c < N; c++) {
        const float mci = m[c][i];
        m[c][i] = m[c][piv];
        m[c][piv] = mci;
        const float pivot = m[c][i];
        for (int r = i + 1; r < N; r++) m[c][r] -= pivot * m[i][r];
    }
}
/* lu.wgsl:37-81 */
static void lu(mat m, uint32_t ia[4], uint32_t ib[4], uint32_t *len, int N) {
    *len = 0;
    for (int k = 0; k < 4; ++k) ia[k] = ib[k] = 0;
    for (int i = 0; i < N; i++) {
        int piv = i;
        float piv_val = fabsf(m[i][i]);
        for (int r = i + 1; r < N; r++) {
            const float abs_val = fabsf(m[i][r]);
            if (abs_val > piv_val) {
                piv = r;
                piv_val = abs_val;
            }
        }
        if (piv_val == 0.0f) continue;
        const float diag = m[i][piv];
        if (piv != i) {
            ia[*len] = (uint32_t)i;
            ib[*len] = (uint32_t)piv;
            (*len)++;
            for (int k = 0; k < i; k++) {
                const float mki = m[k][i];
                m[k][i] = m[k][piv];
                m[k][piv] = mki;
            }
            gauss_step_swap(m, diag, i, piv, N);
        } else {
            gauss_step(m, diag, i, N);
        }
    }
}

/* qr2.wgsl:15-107 (= qr3.wgsl, qr4.wgsl up to DIM) */
static void qr(mat m, mat q, mat rr, int DIM) {
    float diag[4] = {0, 0, 0, 0};
    for (int i = 0; i < DIM; i++) {
        float axis_sq_norm = 0.0f;
        for (int r = i; r < DIM; r++) axis_sq_norm += m[i][r] * m[i][r];
        const float axis_norm = sqrtf(axis_sq_norm);
        const float modulus = fabsf(m[i][i]);
        const float sgn = wsign(m[i][i]);
        const float signed_norm = sgn * axis_norm;
        const float factor = (axis_sq_norm + modulus * axis_norm) * 2.0f;
        m[i][i] += signed_norm;
        if (factor != 0.0f) {
            const float factor_sqrt = sqrtf(factor);
            float norm = 0.0f;
            for (int r = i; r < DIM; r++) {
                m[i][r] /= factor_sqrt;
                norm += m[i][r] * m[i][r];
            }
            norm = sqrtf(norm);
            for (int r = i; r < DIM; r++) m[i][r] /= norm;
            diag[i] = -signed_norm;
        } else {
            diag[i] = signed_norm;
        }
        if (factor != 0.0f) {
            const float sgn2 = wsign(diag[i]);
            for (int c = i; c < DIM; c++) {
                const float m_two = -2.0f * sgn2;
                float f = 0.0f;
                for (int r = i; r < DIM; r++) f += m[i][r] * m[c][r];
                for (int r = i; r < DIM; r++) m[c][r] = m_two * f * m[i][r] + m[c][r] * sgn2;
            }
        }
    }
    memset(q, 0, sizeof(mat));
    for (int i = 0; i < DIM; ++i) q[i][i] = 1.0f;
    for (int i = DIM - 1; i >= 0; i--) {
        const float sgn = wsign(diag[i]);
        for (int c = i; c < DIM; c++) {
            const float m_two = -2.0f * sgn;
            float f = 0.0f;
            for (int r = i; r < DIM; r++) f += m[i][r] * q[c][r];
            for (int r = i; r < DIM; r++) q[c][r] = m_two * f * m[i][r] + q[c][r] * sgn;
        }
    }
    memset(rr, 0, sizeof(mat));
    for (int c = 0; c < DIM; ++c) {
        for (int r = 0; r < c; ++r) rr[c][r] = m[c][r];
        rr[c][c] = fabsf(diag[c]);
    }
}

/* eig2.wgsl:42-56 */
static void eig2_values(float a, float c, float b, float ev[2]) {
    if (c == 0.0f) {
        ev[0] = a;
        ev[1] = b;
        return;
    }
    const float ab = a - b;
    const float sigma = sqrtf(4.0f * c * c + ab * ab);
    ev[0] = (a + b + sigma) / 2.0f;
    ev[1] = (a + b - sigma) / 2.0f;
}
/* eig2.wgsl:15-40 */
static void eig2(mat m, mat vecs, float vals[4]) {
    const float a = m[0][0], c = m[0][1], b = m[1][1];
    memset(vecs, 0, sizeof(mat));
    if (c == 0.0f) {
        vecs[0][0] = 1.0f;
        vecs[1][1] = 1.0f;
        vals[0] = a;
        vals[1] = b;
        return;
    }
    const float ab = a - b;
    const float sigma = sqrtf(4.0f * c * c + ab * ab);
    vals[0] = (a + b + sigma) / 2.0f;
    vals[1] = (a + b - sigma) / 2.0f;
    const float e1x = (a - b + sigma) / (2.0f * c), e2x = (a - b - sigma) / (2.0f * c);
    const float l1 = sqrtf(e1x * e1x + 1.0f * 1.0f), l2 = sqrtf(e2x * e2x + 1.0f * 1.0f);
    vecs[0][0] = e1x / l1;
    vecs[0][1] = 1.0f / l1;
    vecs[1][0] = e2x / l2;
    vecs[1][1] = 1.0f / l2;
}

/* eig3.wgsl:211-282 */
static void tridiagonalize(mat m, float off_diagonal[3], int DIM) {
    off_diagonal[0] = off_diagonal[1] = off_diagonal[2] = 0.0f;
    for (int i = 0; i < DIM - 1; i++) {
        float axis_sq_norm = 0.0f;
        for (int r = i + 1; r < DIM; r++) axis_sq_norm += m[i][r] * m[i][r];
        const float axis_norm = sqrtf(axis_sq_norm);
        const float modulus = fabsf(m[i][i + 1]);
        const float sgn = wsign(m[i][i + 1]);
        const float signed_norm = sgn * axis_norm;
        const float factor = (axis_sq_norm + modulus * axis_norm) * 2.0f;
        m[i][i + 1] += signed_norm;
        if (factor != 0.0f) {
            const float factor_sqrt = sqrtf(factor);
            float norm = 0.0f;
            for (int r = i + 1; r < DIM; r++) {
                m[i][r] /= factor_sqrt;
                norm += m[i][r] * m[i][r];
            }
            norm = sqrtf(norm);
            for (int r = i + 1; r < DIM; r++) m[i][r] /= norm;
            off_diagonal[i] = -signed_norm;
        } else {
            off_diagonal[i] = signed_norm;
        }
        if (factor != 0.0f) {
            float p[4] = {0, 0, 0, 0};
            for (int c = i + 1; c < DIM; c++)
                for (int r = i + 1; r < DIM; r++) p[r] += 2.0f * m[c][r] * m[i][c];
            float dot = 0.0f;
            for (int r = i + 1; r < DIM; r++) dot += m[i][r] * p[r];
            for (int c = i + 1; c < DIM; c++)
                for (int r = i + 1; r < DIM; r++) m[c][r] += 2.0f * dot * m[i][r] * m[i][c] - p[r] * m[i][c] - m[i][r] * p[c];
        }
    }
}
/* eig3.wgsl:162-197 */
static void delimit_subproblem(const float diag[4], float off_diag[3], uint32_t end, float eps, uint32_t *start_out, uint32_t *end_out) {
    uint32_t n = end;
    while (n > 0u) {
        const uint32_t m = n - 1u;
        if (fabsf(off_diag[m]) > eps * (fabsf(diag[n]) + fabsf(diag[m]))) break;
        n -= 1u;
    }
    if (n == 0u) {
        *start_out = 0u;
        *end_out = 0u;
        return;
    }
    uint32_t new_start = n - 1u;
    while (new_start > 0u) {
        const uint32_t m = new_start - 1u;
        if (off_diag[m] == 0.0f || fabsf(off_diag[m]) <= eps * (fabsf(diag[new_start]) + fabsf(diag[m]))) {
            off_diag[m] = 0.0f;
            break;
        }
        new_start -= 1u;
    }
    *start_out = new_start;
    *end_out = n;
}
/* eig3.wgsl:199-209 */
static float wilkinson_shift(float tmm, float tnn, float tmn) {
    const float sq_tmn = tmn * tmn;
    if (sq_tmn != 0.0f) {
        const float d = (tmm - tnn) * 0.5f;
        return tnn - sq_tmn / (d + wsign(d) * sqrtf(d * d + sq_tmn));
    }
    return tnn;
}
/* rot2.wgsl:75-94 (invMulVec :70-72) */
static void rotate_rows(float rc, float rs, mat m, uint32_t i, int DIM) {
    for (int r = 0; r < DIM; r++) {
        const float vx = m[i][r], vy = m[i + 1][r];
        m[i][r] = rc * vx + rs * vy;
        m[i + 1][r] = -rs * vx + rc * vy;
    }
}
/* eig3.wgsl:24-160 (= eig4.wgsl) */
static void eig_n(mat m, mat q, float diag[4], int DIMi) {
    const uint32_t DIM = (uint32_t)DIMi;
    const float EPS = 1.1920929e-7f;
    float m_amax = 0.0f; /* min_max.wgsl:27-30, 44-47: max of |.| */
    {
        float vm[4];
        for (int r = 0; r < DIMi; ++r) {
            /* max(abs(m[0]), max(abs(m[1]), ...)) is associative for non-NaN inputs */
            vm[r] = fabsf(m[0][r]);
            for (int c = 1; c < DIMi; ++c) vm[r] = fmaxf(vm[r], fabsf(m[c][r]));
        }
        m_amax = vm[0];
        for (int r = 1; r < DIMi; ++r) m_amax = fmaxf(m_amax, vm[r]);
    }
    if (m_amax != 0.0f)
        for (int c = 0; c < DIMi; ++c)
            for (int r = 0; r < DIMi; ++r) m[c][r] /= m_amax;
    float tri_off[3];
    tridiagonalize(m, tri_off, DIMi);
    float off_diag[3] = {0, 0, 0};
    for (int i = 0; i < DIMi; ++i) diag[i] = m[i][i];
    for (int i = 0; i < DIMi - 1; ++i) off_diag[i] = fabsf(tri_off[i]);
    memset(q, 0, sizeof(mat));
    for (int i = 0; i < DIMi; ++i) q[i][i] = 1.0f;
    for (uint32_t i = DIM - 2;; i--) {
        const float sgn = wsign(tri_off[i]);
        for (uint32_t c = i; c < DIM; c++) {
            const float m_two = -2.0f * sgn;
            float factor = 0.0f;
            for (uint32_t r = i + 1; r < DIM; r++) factor += m[i][r] * q[c][r];
            for (uint32_t r = i + 1; r < DIM; r++) q[c][r] = m_two * factor * m[i][r] + q[c][r] * sgn;
        }
        if (i == 0) break;
    }
    uint32_t start, end;
    delimit_subproblem(diag, off_diag, DIM - 1, EPS, &start, &end);
    int niter = 0;
    while (end != start && niter < ORC_EIG_MAX_SWEEPS) {
        const uint32_t subdim = end - start + 1u;
        if (subdim > 2u) {
            const uint32_t mm = end - 1u, n = end;
            const float shift = wilkinson_shift(diag[mm], diag[n], off_diag[mm]);
            float vx = diag[start] - shift, vy = off_diag[start];
            for (uint32_t i = start; i < n; i++) {
                const uint32_t j = i + 1u;
                /* rot2.wgsl:29-37 cancel_y, :15-17 is_valid */
                float rc = 0.0f, rs = 0.0f;
                if (vy != 0.0f) {
                    const float r = wsign(vx) / sqrtf(vx * vx + vy * vy);
                    rc = vx * r;
                    rs = -vy * r;
                }
                if (rc != 0.0f || rs != 0.0f) {
                    if (i > start) off_diag[i - 1] = wsign(vx) * sqrtf(vx * vx + vy * vy);
                    const float mii = diag[i], mjj = diag[j], mij = off_diag[i];
                    const float cc = rc * rc, ss = rs * rs, cs = rc * rs;
                    const float b = cs * 2.0f * mij;
                    diag[i] = (cc * mii + ss * mjj) - b;
                    diag[j] = (ss * mii + cc * mjj) + b;
                    off_diag[i] = cs * (mii - mjj) + mij * (cc - ss);
                    if (i != n - 1) {
                        vx = off_diag[i];
                        vy = -rs * off_diag[i + 1];
                        off_diag[i + 1] *= rc;
                    }
                    rotate_rows(rc, -rs, q, i, DIMi); /* Rot::inv(rot), rot2.wgsl:53-55 */
                } else {
                    break;
                }
            }
            if (fabsf(off_diag[mm]) <= EPS * (fabsf(diag[mm]) + fabsf(diag[n]))) end -= 1u;
        } else if (subdim == 2u) {
            float ev[2];
            eig2_values(diag[start], off_diag[start], diag[start + 1], ev);
            const float bx = ev[0] - diag[start + 1], by = off_diag[start];
            diag[start] = ev[0];
            diag[start + 1] = ev[1];
            const float basis_len = sqrtf(bx * bx + by * by);
            if (basis_len > EPS) {
                const float s = wsign(bx) / basis_len;
                rotate_rows(bx * s, by * s, q, start, DIMi);
            }
            end -= 1u;
        }
        delimit_subproblem(diag, off_diag, end, EPS, &start, &end);
        niter++;
    }
    for (int i = 0; i < DIMi; ++i) diag[i] *= m_amax;
}

/* trig.wgsl:26-41 */
static float stable_atan2(float y, float x) {
    const float PI = 3.14159265358979323846264338327950288f;
    const float ang = atanf(y / x);
    if (x > 0.0f) return ang;
    if (x < 0.0f && y > 0.0f) return ang + PI;
    if (x < 0.0f && y < 0.0f) return ang - PI;
    return 0.0f;
}
/* svd2.wgsl:12-39 */
static void svd2(mat m, mat u, float s[4], mat vt) {
    const float e = (m[0][0] + m[1][1]) * 0.5f, f = (m[0][0] - m[1][1]) * 0.5f;
    const float g = (m[0][1] + m[1][0]) * 0.5f, h = (m[0][1] - m[1][0]) * 0.5f;
    const float q = sqrtf(e * e + h * h), r = sqrtf(f * f + g * g);
    const float sx = q + r, sy = q - r;
    const float sy_sign = sy < 0.0f ? -1.0f : 1.0f;
    s[0] = sx;
    s[1] = sy * sy_sign;
    const float a1 = stable_atan2(g, f), a2 = stable_atan2(h, e);
    const float theta = (a2 - a1) * 0.5f, phi = (a2 + a1) * 0.5f;
    const float st = sinf(theta), ct = cosf(theta), sp = sinf(phi), cp = cosf(phi);
    memset(u, 0, sizeof(mat));
    memset(vt, 0, sizeof(mat));
    u[0][0] = cp;
    u[0][1] = sp;
    u[1][0] = -sp;
    u[1][1] = cp;
    vt[0][0] = ct;
    vt[0][1] = st * sy_sign;
    vt[1][0] = -st;
    vt[1][1] = ct * sy_sign;
}

/* svd3.wgsl:55-80 */
static float rsqrt_steps(float val, int steps) {
    float x = val;
    const float xhalf = -0.5f * x;
    int32_t i;
    memcpy(&i, &x, 4);
    i = 0x5f375a82 - (i >> 1);
    memcpy(&x, &i, 4);
    for (int k = 0; k < steps; k++) x = x * fmaf(x * x, xhalf, 1.5f);
    return x;
}
#define RSQRT(v) rsqrt_steps((v), 4)
#define RSQRT1(v) rsqrt_steps((v), 6)
typedef struct { float mxx, myx, myy, mzx, mzy, mzz; } sym3;
/* svd3.wgsl:116-129 */
static void approx_givens(const sym3 *A, float *ch, float *sh) {
    const float gch = 2.0f * (A->mxx - A->myy), gsh = A->myx;
    int b = 5.828427124f * gsh * gsh < gch * gch;
    const float w = RSQRT(fmaf(gch, gch, gsh * gsh));
    if (w != w) b = 0;
    if (b) {
        *ch = w * gch;
        *sh = w * gsh;
    } else {
        *ch = 0.923879532f;
        *sh = 0.3826834323f;
    }
}
/* svd3.wgsl:132-167 */
static void jacobi_conjugation(int x, int y, int z, sym3 *S, float q[4]) {
    float gch, gsh;
    approx_givens(S, &gch, &gsh);
    const float scale = 1.0f / fmaf(gch, gch, gsh * gsh);
    const float a = fmaf(gch, gch, -gsh * gsh) * scale;
    const float b = 2.0f * gsh * gch * scale;
    sym3 T = *S;
    S->mxx = fmaf(a, fmaf(a, T.mxx, b * T.myx), b * (fmaf(a, T.myx, b * T.myy)));
    S->myx = fmaf(a, fmaf(-b, T.mxx, a * T.myx), b * (fmaf(-b, T.myx, a * T.myy)));
    S->myy = fmaf(-b, fmaf(-b, T.mxx, a * T.myx), a * (fmaf(-b, T.myx, a * T.myy)));
    S->mzx = fmaf(a, T.mzx, b * T.mzy);
    S->mzy = fmaf(-b, T.mzx, a * T.mzy);
    S->mzz = T.mzz;
    const float tmp[3] = {q[0] * gsh, q[1] * gsh, q[2] * gsh};
    gsh *= q[3];
    q[z] = fmaf(q[z], gch, gsh);
    q[3] = fmaf(q[3], gch, -tmp[z]);
    q[x] = fmaf(q[x], gch, tmp[y]);
    q[y] = fmaf(q[y], gch, -tmp[x]);
    T.mxx = S->myy;
    T.myx = S->mzy;
    T.myy = S->mzz;
    T.mzx = S->myx;
    T.mzy = S->mzx;
    T.mzz = S->mxx;
    *S = T;
}
/* svd3.wgsl:215-229 */
static void qr_givens_quaternion(float a1, float a2, float *ch_out, float *sh_out) {
    const float epsilon = 1e-6f;
    const float rho = 1.0f / RSQRT1(fmaf(a1, a1, a2 * a2));
    float ch = fabsf(a1) + fmaxf(rho, epsilon);
    float sh = rho > epsilon ? a2 : 0.0f;
    if (a1 < 0.0f) {
        const float t = sh;
        sh = ch;
        ch = t;
    }
    const float w = RSQRT(fmaf(ch, ch, sh * sh));
    *ch_out = ch * w;
    *sh_out = sh * w;
}
/* mat3 product as WGSL `A * B`: out[c][r] = sum_k A[k][r] * B[c][k], k ascending, no contraction */
static void mul3(mat out, mat A, mat B) {
    mat t;
    memset(t, 0, sizeof(mat));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) t[c][r] = A[0][r] * B[c][0] + A[1][r] * B[c][1] + A[2][r] * B[c][2];
    memcpy(out, t, sizeof(mat));
}
static void neg_swap3(int c, float *x, float *y) { /* svd3.wgsl:108-112 */
    for (int k = 0; k < 3; ++k) {
        const float x0 = -x[k];
        x[k] = c ? y[k] : x[k];
        y[k] = c ? x0 : y[k];
    }
}
/* svd3.wgsl:291-305 (+ :170-212 sortSingularValues, :232-288 QRDecomposition, quat.wgsl:31-53 toMatrix) */
static void svd3(mat A, mat U, float S[4], mat Vt) {
    mat At, ata, V, B;
    memset(At, 0, sizeof(mat));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) At[c][r] = A[r][c];
    mul3(ata, At, A);
    sym3 s = {ata[0][0], ata[0][1], ata[1][1], ata[0][2], ata[1][2], ata[2][2]};
    float qv[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    for (int it = 0; it < 12; ++it) {
        jacobi_conjugation(0, 1, 2, &s, qv);
        jacobi_conjugation(1, 2, 0, &s, qv);
        jacobi_conjugation(2, 0, 1, &s, qv);
    }
    {
        const float i = qv[0], j = qv[1], k = qv[2], w = qv[3];
        const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
        const float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f, ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
        memset(V, 0, sizeof(mat));
        V[0][0] = ww + ii - jj - kk;
        V[0][1] = wk + ij;
        V[0][2] = ik - wj;
        V[1][0] = ij - wk;
        V[1][1] = ww - ii + jj - kk;
        V[1][2] = wi + jk;
        V[2][0] = wj + ik;
        V[2][1] = jk - wi;
        V[2][2] = ww - ii - jj + kk;
    }
    mul3(B, A, V);
    {
        float rho1 = B[0][0] * B[0][0] + B[0][1] * B[0][1] + B[0][2] * B[0][2];
        float rho2 = B[1][0] * B[1][0] + B[1][1] * B[1][1] + B[1][2] * B[1][2];
        float rho3 = B[2][0] * B[2][0] + B[2][1] * B[2][1] + B[2][2] * B[2][2];
        int c = rho1 < rho2;
        neg_swap3(c, B[0], B[1]);
        neg_swap3(c, V[0], V[1]);
        if (c) {
            const float t = rho1;
            rho1 = rho2;
            rho2 = t;
        }
        c = rho1 < rho3;
        neg_swap3(c, B[0], B[2]);
        neg_swap3(c, V[0], V[2]);
        if (c) {
            const float t = rho1;
            rho1 = rho3;
            rho3 = t;
        }
        c = rho2 < rho3;
        neg_swap3(c, B[1], B[2]);
        neg_swap3(c, V[1], V[2]);
    }
    float g1c, g1s, g2c, g2s, g3c, g3s;
    qr_givens_quaternion(B[0][0], B[0][1], &g1c, &g1s);
    float a = fmaf(-2.0f, g1s * g1s, 1.0f), b = 2.0f * g1c * g1s;
    float r00 = fmaf(a, B[0][0], b * B[0][1]), r01 = fmaf(a, B[1][0], b * B[1][1]), r02 = fmaf(a, B[2][0], b * B[2][1]);
    float r10 = fmaf(-b, B[0][0], a * B[0][1]), r11 = fmaf(-b, B[1][0], a * B[1][1]), r12 = fmaf(-b, B[2][0], a * B[2][1]);
    float r20 = B[0][2], r21 = B[1][2], r22 = B[2][2];
    qr_givens_quaternion(r00, r20, &g2c, &g2s);
    a = fmaf(-2.0f, g2s * g2s, 1.0f);
    b = 2.0f * g2c * g2s;
    const float b00 = fmaf(a, r00, b * r20), b01 = fmaf(a, r01, b * r21), b02 = fmaf(a, r02, b * r22);
    const float b10 = r10, b11 = r11, b12 = r12;
    const float b20 = fmaf(-b, r00, a * r20), b21 = fmaf(-b, r01, a * r21), b22 = fmaf(-b, r02, a * r22);
    qr_givens_quaternion(b11, b21, &g3c, &g3s);
    a = fmaf(-2.0f, g3s * g3s, 1.0f);
    b = 2.0f * g3c * g3s;
    r00 = b00;
    r01 = b01;
    r02 = b02;
    r10 = fmaf(a, b10, b * b20);
    r11 = fmaf(a, b11, b * b21);
    r12 = fmaf(a, b12, b * b22);
    r20 = fmaf(-b, b10, a * b20);
    r21 = fmaf(-b, b11, a * b21);
    r22 = fmaf(-b, b12, a * b22);
    (void)r01; (void)r02; (void)r10; (void)r12; (void)r20; (void)r21;
    const float sh12 = 2.0f * fmaf(g1s, g1s, -0.5f), sh22 = 2.0f * fmaf(g2s, g2s, -0.5f), sh32 = 2.0f * fmaf(g3s, g3s, -0.5f);
    const float q00 = sh12 * sh22;
    const float q01 = fmaf(4.0f * g2c * g3c, sh12 * g2s * g3s, 2.0f * g1c * g1s * sh32);
    const float q02 = fmaf(4.0f * g1c * g3c, g1s * g3s, -2.0f * g2c * sh12 * g2s * sh32);
    const float q10 = -2.0f * g1c * g1s * sh22;
    const float q11 = fmaf(-8.0f * g1c * g2c * g3c, g1s * g2s * g3s, sh12 * sh32);
    const float q12 = fmaf(-2.0f * g3c, g3s, 4.0f * g1s * fmaf(g3c * g1s, g3s, g1c * g2c * g2s * sh32));
    const float q20 = 2.0f * g2c * g2s;
    const float q21 = -2.0f * g3c * sh22 * g3s;
    const float q22 = sh22 * sh32;
    memset(U, 0, sizeof(mat));
    U[0][0] = q00; U[0][1] = q10; U[0][2] = q20;
    U[1][0] = q01; U[1][1] = q11; U[1][2] = q21;
    U[2][0] = q02; U[2][1] = q12; U[2][2] = q22;
    S[0] = r00;
    S[1] = r11;
    S[2] = r22;
    memset(Vt, 0, sizeof(mat));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Vt[c][r] = V[r][c];
}

/* inv.wgsl:8-20 */
static void inv2(mat m, mat o) {
    mat adj;
    memset(adj, 0, sizeof(mat));
    adj[0][0] = m[1][1];
    adj[0][1] = -m[0][1];
    adj[1][0] = -m[1][0];
    adj[1][1] = m[0][0];
    const float det = m[0][0] * m[1][1] - m[1][0] * m[0][1];
    const float s = 1.0f / det;
    memset(o, 0, sizeof(mat));
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 2; ++r) o[c][r] = adj[c][r] * s;
}
/* inv.wgsl:26-45 */
static void inv3(mat m, mat o) {
    mat adj;
    memset(adj, 0, sizeof(mat));
    adj[0][0] = (m[1][1] * m[2][2] - m[2][1] * m[1][2]);
    adj[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]);
    adj[2][0] = (m[1][0] * m[2][1] - m[2][0] * m[1][1]);
    adj[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]);
    adj[1][1] = (m[0][0] * m[2][2] - m[2][0] * m[0][2]);
    adj[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]);
    adj[0][2] = (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    adj[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]);
    adj[2][2] = (m[0][0] * m[1][1] - m[1][0] * m[0][1]);
    const float det = (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]));
    const float s = 1.0f / det;
    memset(o, 0, sizeof(mat));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) o[c][r] = adj[c][r] * s;
}
/* inv.wgsl:52-88 */
static void inv4(mat m, mat o) {
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
    mat adj;
    adj[0][0] = (m[1][1] * sf00 - m[1][2] * sf01 + m[1][3] * sf02);
    adj[1][0] = -(m[1][0] * sf00 - m[1][2] * sf03 + m[1][3] * sf04);
    adj[2][0] = (m[1][0] * sf01 - m[1][1] * sf03 + m[1][3] * sf05);
    adj[3][0] = -(m[1][0] * sf02 - m[1][1] * sf04 + m[1][2] * sf05);
    adj[0][1] = -(m[0][1] * sf00 - m[0][2] * sf01 + m[0][3] * sf02);
    adj[1][1] = (m[0][0] * sf00 - m[0][2] * sf03 + m[0][3] * sf04);
    adj[2][1] = -(m[0][0] * sf01 - m[0][1] * sf03 + m[0][3] * sf05);
    adj[3][1] = (m[0][0] * sf02 - m[0][1] * sf04 + m[0][2] * sf05);
    adj[0][2] = (m[0][1] * sf06 - m[0][2] * sf07 + m[0][3] * sf08);
    adj[1][2] = -(m[0][0] * sf06 - m[0][2] * sf09 + m[0][3] * sf10);
    adj[2][2] = (m[0][0] * sf11 - m[0][1] * sf09 + m[0][3] * sf12);
    adj[3][2] = -(m[0][0] * sf08 - m[0][1] * sf10 + m[0][2] * sf12);
    adj[0][3] = -(m[0][1] * sf13 - m[0][2] * sf14 + m[0][3] * sf15);
    adj[1][3] = (m[0][0] * sf13 - m[0][2] * sf16 + m[0][3] * sf17);
    adj[2][3] = -(m[0][0] * sf14 - m[0][1] * sf16 + m[0][3] * sf18);
    adj[3][3] = (m[0][0] * sf15 - m[0][1] * sf17 + m[0][2] * sf18);
    const float det = (m[0][0] * adj[0][0] + m[0][1] * adj[1][0] + m[0][2] * adj[2][0] + m[0][3] * adj[3][0]);
    const float s = 1.0f / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) o[c][r] = adj[c][r] * s;
}

/* 4-byte words of one output element (0: unsupported combination) */
uint32_t orc_geom_out_words(int op, int dim) {
    if (dim < 2 || dim > 4) return 0;
    const int mw = mat_words(dim), cs = col_stride(dim);
    switch (op) {
    case GEOM_CHOLESKY:
    case GEOM_INV: return (uint32_t)mw;
    case GEOM_LU: return dim == 2 ? 10u : dim == 3 ? 20u : 28u;
    case GEOM_QR: return (uint32_t)(2 * mw);
    case GEOM_EIG: return (uint32_t)(mw + cs);
    case GEOM_SVD: return dim == 4 ? 0u : (uint32_t)(2 * mw + cs);
    default: return 0;
    }
}
uint32_t orc_geom_in_words(int dim) { return dim < 2 || dim > 4 ? 0u : (uint32_t)mat_words(dim); }

/* out[i] = f(in[i]) for i < n — the shape of every test kernel in geometry/ (e.g. cholesky.rs:53-63). */
int orc_geom_batch(int op, int dim, const float *in, void *out, uint64_t n) {
    const uint32_t ow = orc_geom_out_words(op, dim);
    if (!ow) return ORC_UNSUPPORTED;
    const int mw = mat_words(dim), cs = col_stride(dim);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        mat m, a, b;
        float v[4] = {0, 0, 0, 0};
        float *o = (float *)out + (uint64_t)i * ow;
        load_mat(m, in + (uint64_t)i * mw, dim);
        switch (op) {
        case GEOM_CHOLESKY:
            cholesky(m, dim);
            store_mat(o, m, dim);
            break;
        case GEOM_INV:
            if (dim == 2) inv2(m, a);
            else if (dim == 3) inv3(m, a);
            else inv4(m, a);
            store_mat(o, a, dim);
            break;
        case GEOM_LU: {
            uint32_t ia[4], ib[4], len;
            lu(m, ia, ib, &len, dim);
            store_mat(o, m, dim);
            uint32_t *p = (uint32_t *)(o + mw);
            for (uint32_t k = 0; k < ow - (uint32_t)mw; ++k) p[k] = 0;
            for (int k = 0; k < dim; ++k) {
                p[k] = ia[k];
                p[cs + k] = ib[k];
            }
            p[dim == 3 ? 7 : 2 * cs] = len; /* vec3<u32> ib is 12 bytes: len packs right behind it */
            break;
        }
        case GEOM_QR:
            qr(m, a, b, dim);
            store_mat(o, a, dim);
            store_mat(o + mw, b, dim);
            break;
        case GEOM_EIG:
            if (dim == 2) eig2(m, a, v);
            else eig_n(m, a, v, dim);
            store_mat(o, a, dim);
            store_vec(o + mw, v, dim);
            break;
        case GEOM_SVD:
            if (dim == 2) svd2(m, a, v, b);
            else svd3(m, a, v, b);
            store_mat(o, a, dim);
            store_vec(o + mw, v, dim);
            store_mat(o + mw + cs, b, dim);
            break;
        }
    }
    return ORC_OK;
}
