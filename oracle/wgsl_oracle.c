/*
 * wgsl_oracle.c — CPU restatement of the wgebra linalg WGSL kernels.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` leg may load this library, and only as
 * the checker / reported CPU baseline.  The product path (wgmath_b200/) never links,
 * imports or calls anything in oracle/.
 *
 * What it restates (all paths relative to /root/reference/crates/wgebra/src/linalg/):
 *   shape.wgsl:36-66      iv / im / it / div_ceil4 / with_vec4_elts (column-major branch; the ROW_MAJOR
 *                         addressing of :49-53 is restated by orc_gemm_ord / orc_gemv_ord)
 *   gemm.wgsl:81-113      gemm           (thread owns 4 rows x all N, 4x4 blocks over K)
 *   gemm.wgsl:29-78       gemm_fast      (64 threads split K by 256, 6-step mat4x4 tree)
 *   gemm.wgsl:116-148     gemm_tr
 *   gemm.wgsl:151-199     gemm_tr_fast
 *   gemv.wgsl:68-90       gemv
 *   gemv.wgsl:29-65       gemv_fast      (32 threads split K by 128, 5-step vec4 tree)
 *   gemv.wgsl:93-115      gemv_tr
 *   gemv.wgsl:118-154     gemv_tr_fast
 *   op_assign.wgsl:14-47  add/sub/mul/div/copy
 *   reduce.wgsl:12-96     128 strided partials, 7-step tree, init constants +-3.4e38
 * and the Rust-side dispatch rules:
 *   gemm.rs:78-126        dimension asserts, grid shape per variant
 *   gemv.rs:77-136        dimension asserts, GemvTrFast -> GemvTr fallback (:99-104),
 *                         assert out_nrows % 4 == 0 for the *_fast variants (:122)
 *   op_assign.rs:82-93    length assert, ceil(n/64) groups
 *   reduce.rs:100-113     exactly one workgroup of 128
 *   kernel.rs:140-148     zero-volume grid => dispatch silently skipped
 *
 * Arithmetic: IEEE binary32, products and sums evaluated left-to-right, no FMA contraction
 * (build with -ffp-contract=off).  WGSL leaves contraction and the order inside
 * mat4x4*mat4x4 implementation-defined, so bit-parity with "the" reference is undefined;
 * this file fixes one legal evaluation order.  PARITY PINNING: the reference holds no golden
 * vectors for this path; its four unit tests compare against nalgebra on unseeded random
 * data (gemm.rs:144-202, gemv.rs:153-197, reduce.rs:139-179) and on one deterministic
 * fixture (op_assign.rs:109-157).  tests/test_oracle.py replays those four tests against
 * this file with an independent f64 reference in nalgebra's place.  Beyond them: parity
 * unpinned.
 *
 * "Workgroups" and "invocations" become loops; the outermost loop over workgroups /
 * invocations is OpenMP-parallel so the same code serves as the CPU baseline on all cores.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint32_t nrows, ncols, nmats, stride, stride_mat, offset;
} shape_t; /* byte-identical to wgcore::shapes::ViewShape (crates/wgcore/src/shapes.rs:9-21) */

typedef struct { float x, y, z, w; } vec4;
typedef struct { vec4 c[4]; } mat4; /* column vectors, as WGSL mat4x4<f32> */

enum { ORC_OK = 0, ORC_DIM_MISMATCH = 2, ORC_UNSUPPORTED = 4 };

/* ---- shape.wgsl ------------------------------------------------------------------ */
static inline uint32_t div_ceil4(uint32_t a) { return (a + 3u) / 4u; }           /* :40-42 */
static inline uint32_t iv(shape_t v, uint32_t i) { return v.offset + i; }        /* :36-38 */
static inline uint32_t im(shape_t v, uint32_t i, uint32_t j) {                   /* :60-62 */
    return v.offset + i + j * v.stride;
}
static inline uint32_t it(shape_t v, uint32_t i, uint32_t j, uint32_t t) {       /* :45-47 */
    return t * v.stride_mat + im(v, i, j);
}
static inline shape_t with_vec4_elts(shape_t s) {                                /* :64-66 */
    shape_t r = { div_ceil4(s.nrows), s.ncols, s.nmats, div_ceil4(s.stride),
                  div_ceil4(s.stride_mat), s.offset / 4u };
    return r;
}

/* ---- WGSL vector / matrix arithmetic, one fixed legal evaluation order ------------ */
static inline vec4 v4_zero(void) { vec4 r = {0.f, 0.f, 0.f, 0.f}; return r; }
static inline vec4 v4_add(vec4 a, vec4 b) { vec4 r = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; return r; }
static inline mat4 m4_zero(void) { mat4 r; memset(&r, 0, sizeof r); return r; }
static inline mat4 m4_add(mat4 a, mat4 b) {
    mat4 r;
    for (int c = 0; c < 4; ++c) r.c[c] = v4_add(a.c[c], b.c[c]);
    return r;
}
/* mat4x4 * vec4 : linear combination of columns, left to right */
static inline vec4 m4_mul_v4(mat4 m, vec4 v) {
    vec4 r;
    r.x = m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z + m.c[3].x * v.w;
    r.y = m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z + m.c[3].y * v.w;
    r.z = m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z + m.c[3].z * v.w;
    r.w = m.c[0].w * v.x + m.c[1].w * v.y + m.c[2].w * v.z + m.c[3].w * v.w;
    return r;
}
static inline mat4 m4_mul(mat4 a, mat4 b) {
    mat4 r;
    for (int c = 0; c < 4; ++c) r.c[c] = m4_mul_v4(a, b.c[c]);
    return r;
}
static inline mat4 m4_transpose(mat4 a) {
    mat4 r;
    r.c[0].x = a.c[0].x; r.c[0].y = a.c[1].x; r.c[0].z = a.c[2].x; r.c[0].w = a.c[3].x;
    r.c[1].x = a.c[0].y; r.c[1].y = a.c[1].y; r.c[1].z = a.c[2].y; r.c[1].w = a.c[3].y;
    r.c[2].x = a.c[0].z; r.c[2].y = a.c[1].z; r.c[2].z = a.c[2].z; r.c[2].w = a.c[3].z;
    r.c[3].x = a.c[0].w; r.c[3].y = a.c[1].w; r.c[3].z = a.c[2].w; r.c[3].w = a.c[3].w;
    return r;
}
static inline mat4 load4(const vec4 *p, uint32_t ia, uint32_t stride) {
    mat4 r;
    r.c[0] = p[ia];
    r.c[1] = p[ia + stride];
    r.c[2] = p[ia + 2u * stride];
    r.c[3] = p[ia + 3u * stride];
    return r;
}
static inline void store4(vec4 *p, uint32_t io, uint32_t stride, mat4 m) {
    p[io] = m.c[0];
    p[io + stride] = m.c[1];
    p[io + stride * 2u] = m.c[2];
    p[io + stride * 3u] = m.c[3];
}

/* ---- gemm.rs:78-126 dispatch rules ------------------------------------------------ */
enum { GEMM = 0, GEMM_FAST = 1, GEMM_TR = 2, GEMM_TR_FAST = 3 };

static int gemm_check(shape_t so, shape_t s1, shape_t s2, int variant) {
    uint32_t m_rows, m_cols;
    if (variant == GEMM || variant == GEMM_FAST) { m_rows = s1.nrows; m_cols = s1.ncols; }
    else { m_rows = s1.ncols; m_cols = s1.nrows; }
    if (m_cols != s2.nrows) return ORC_DIM_MISMATCH;  /* gemm.rs:91 */
    if (m_rows != so.nrows) return ORC_DIM_MISMATCH;  /* :92 */
    if (so.ncols != s2.ncols) return ORC_DIM_MISMATCH; /* :93 */
    if (so.nmats != s1.nmats) return ORC_DIM_MISMATCH; /* :94 */
    if (so.nmats != s2.nmats) return ORC_DIM_MISMATCH; /* :95 */
    return ORC_OK;
}

/* gemm.wgsl:81-113 (one invocation) */
static void k_gemm(shape_t so, shape_t s1, shape_t s2, vec4 *out, const vec4 *m1, const vec4 *m2,
                   uint32_t gx, uint32_t gy) {
    if (gx < s1.nrows) {
        for (uint32_t k = 0; k < s2.ncols; k += 4u) {
            mat4 sum = m4_zero();
            for (uint32_t j = 0; j < s1.ncols; j += 4u) {
                mat4 a = load4(m1, it(s1, gx, j, gy), s1.stride);
                mat4 b = load4(m2, it(s2, j / 4u, k, gy), s2.stride);
                sum = m4_add(sum, m4_mul(a, b));
            }
            store4(out, it(so, gx, k, gy), so.stride, sum);
        }
    }
}

/* gemm.wgsl:116-148 (one invocation) */
static void k_gemm_tr(shape_t so, shape_t s1, shape_t s2, vec4 *out, const vec4 *m1, const vec4 *m2,
                      uint32_t gx, uint32_t gy) {
    if (gx < (s1.ncols + 3u) / 4u) {
        for (uint32_t k = 0; k < s2.ncols; k += 4u) {
            mat4 sum = m4_zero();
            for (uint32_t j = 0; j < s1.nrows; j++) {
                mat4 a = load4(m1, it(s1, j, gx * 4u, gy), s1.stride);
                mat4 b = load4(m2, it(s2, j, k, gy), s2.stride);
                sum = m4_add(sum, m4_mul(m4_transpose(a), b));
            }
            store4(out, it(so, gx, k, gy), so.stride, sum);
        }
    }
}

/* gemm.wgsl:29-78 and :151-199 (one workgroup of 64) */
static void k_gemm_fast_wg(shape_t so, shape_t s1, shape_t s2, vec4 *out, const vec4 *m1,
                           const vec4 *m2, uint32_t wx, uint32_t wy, int tr) {
    enum { WG = 64 };
    mat4 sketch[WG];
    for (uint32_t k = 0; k < s2.ncols; k += 4u) {
        for (uint32_t l = 0; l < WG; ++l) {
            mat4 sum = m4_zero();
            if (!tr) {
                for (uint32_t j = 0; j < s1.ncols; j += 4u * WG) {
                    mat4 a = load4(m1, it(s1, wx, j + l * 4u, wy), s1.stride);
                    mat4 b = load4(m2, it(s2, j / 4u + l, k, wy), s2.stride);
                    sum = m4_add(sum, m4_mul(a, b));
                }
            } else {
                for (uint32_t j = 0; j < s1.nrows; j += WG) {
                    mat4 a = load4(m1, it(s1, j + l, wx * 4u, wy), s1.stride);
                    mat4 b = load4(m2, it(s2, j + l, k, wy), s2.stride);
                    sum = m4_add(sum, m4_mul(m4_transpose(a), b));
                }
            }
            sketch[l] = sum;
        }
        for (uint32_t stride = 32u; stride >= 1u; stride >>= 1)           /* reduce_sum x6 */
            for (uint32_t l = 0; l < stride; ++l) sketch[l] = m4_add(sketch[l], sketch[l + stride]);
        store4(out, it(so, wx, k, wy), so.stride, sketch[0]);
    }
}

int orc_gemm(int variant, float *out, const shape_t *pso, const float *m1, const shape_t *ps1,
             const float *m2, const shape_t *ps2) {
    int rc = gemm_check(*pso, *ps1, *ps2, variant);
    if (rc) return rc;
    const uint32_t out_rows = pso->nrows, out_mats = pso->nmats;
    shape_t so = with_vec4_elts(*pso), s1 = with_vec4_elts(*ps1), s2 = with_vec4_elts(*ps2);
    vec4 *o = (vec4 *)out; const vec4 *a = (const vec4 *)m1, *b = (const vec4 *)m2;
    if (variant == GEMM || variant == GEMM_TR) {
        const int64_t ninv = (int64_t)((out_rows + 63u) / 64u) * 64;      /* gemm.rs:112 */
        if (ninv * out_mats == 0) return ORC_OK;                           /* kernel.rs:144 */
        for (uint32_t gy = 0; gy < out_mats; ++gy) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int64_t gx = 0; gx < ninv; ++gx) {
                if (variant == GEMM) k_gemm(so, s1, s2, o, a, b, (uint32_t)gx, gy);
                else k_gemm_tr(so, s1, s2, o, a, b, (uint32_t)gx, gy);
            }
        }
    } else {
        const int64_t nwg = (out_rows + 3u) / 4u;                          /* gemm.rs:114 */
        if (nwg * out_mats == 0) return ORC_OK;
        for (uint32_t gy = 0; gy < out_mats; ++gy) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int64_t wx = 0; wx < nwg; ++wx)
                k_gemm_fast_wg(so, s1, s2, o, a, b, (uint32_t)wx, gy, variant == GEMM_TR_FAST);
        }
    }
    return ORC_OK;
}

/* ---- gemv ---------------------------------------------------------------------- */
enum { GEMV = 0, GEMV_FAST = 1, GEMV_TR = 2, GEMV_TR_FAST = 3 };

/* gemv.wgsl:68-90 */
static void k_gemv(shape_t so, shape_t sm, shape_t sv, vec4 *out, const vec4 *m, const vec4 *v,
                   uint32_t gx, uint32_t gy, uint32_t gz) {
    if (gx < sm.nrows) {
        vec4 sum = v4_zero();
        for (uint32_t j = 0; j < sm.ncols; j += 4u) {
            mat4 a = load4(m, it(sm, gx, j, gz), sm.stride);
            sum = v4_add(sum, m4_mul_v4(a, v[it(sv, j / 4u, gy, gz)]));
        }
        out[it(so, gx, gy, gz)] = sum;
    }
}
/* gemv.wgsl:93-115 */
static void k_gemv_tr(shape_t so, shape_t sm, shape_t sv, vec4 *out, const vec4 *m, const vec4 *v,
                      uint32_t gx, uint32_t gy, uint32_t gz) {
    if (gx < (sm.ncols + 3u) / 4u) {
        vec4 sum = v4_zero();
        for (uint32_t j = 0; j < sm.nrows; j++) {
            mat4 a = load4(m, it(sm, j, gx * 4u, gz), sm.stride);
            sum = v4_add(sum, m4_mul_v4(m4_transpose(a), v[it(sv, j, gy, gz)]));
        }
        out[it(so, gx, gy, gz)] = sum;
    }
}
/* gemv.wgsl:29-65 and :118-154 (one workgroup of 32) */
static void k_gemv_fast_wg(shape_t so, shape_t sm, shape_t sv, vec4 *out, const vec4 *m,
                           const vec4 *v, uint32_t wx, uint32_t wy, uint32_t wz, int tr) {
    enum { WG = 32 };
    vec4 sketch[WG];
    for (uint32_t l = 0; l < WG; ++l) {
        vec4 sum = v4_zero();
        if (!tr) {
            for (uint32_t j = 0; j < sm.ncols; j += 4u * WG) {
                mat4 a = load4(m, it(sm, wx, j + l * 4u, wz), sm.stride);
                sum = v4_add(sum, m4_mul_v4(a, v[it(sv, j / 4u + l, wy, wz)]));
            }
        } else {
            for (uint32_t j = 0; j < sm.nrows; j += WG) {
                mat4 a = load4(m, it(sm, j + l, wx * 4u, wz), sm.stride);
                sum = v4_add(sum, m4_mul_v4(m4_transpose(a), v[it(sv, j + l, wy, wz)]));
            }
        }
        sketch[l] = sum;
    }
    for (uint32_t stride = 16u; stride >= 1u; stride >>= 1)               /* reduce_sum x5 */
        for (uint32_t l = 0; l < stride; ++l) sketch[l] = v4_add(sketch[l], sketch[l + stride]);
    out[it(so, wx, wy, wz)] = sketch[0];
}

/* Returns the variant actually run in *ran (the GemvTrFast -> GemvTr fallback, gemv.rs:99-104). */
int orc_gemv(int variant, float *out, const shape_t *pso, const float *m, const shape_t *psm,
             const float *v, const shape_t *psv, int *ran) {
    const uint32_t out_nrows = pso->nrows, out_ncols = pso->ncols, out_nmats = pso->nmats;
    uint32_t m_rows, m_cols;
    if (variant == GEMV || variant == GEMV_FAST) { m_rows = psm->nrows; m_cols = psm->ncols; }
    else { m_rows = psm->ncols; m_cols = psm->nrows; }
    if (m_cols != psv->nrows) return ORC_DIM_MISMATCH;                     /* gemv.rs:89 */
    if (m_rows != out_nrows) return ORC_DIM_MISMATCH;                      /* :90 */
    if (variant == GEMV_TR_FAST && psm->nrows % (32u * 4u) != 0) variant = GEMV_TR; /* :99-104 */
    if ((variant == GEMV_FAST || variant == GEMV_TR_FAST) && out_nrows % 4u != 0)
        return ORC_DIM_MISMATCH;                                           /* :122 assert */
    if (ran) *ran = variant;
    shape_t so = with_vec4_elts(*pso), sm = with_vec4_elts(*psm), sv = with_vec4_elts(*psv);
    vec4 *o = (vec4 *)out; const vec4 *a = (const vec4 *)m, *b = (const vec4 *)v;
    const int fast = (variant == GEMV_FAST || variant == GEMV_TR_FAST);
    const int64_t gxn = fast ? (int64_t)((out_nrows + 3u) / 4u)            /* :124 */
                             : (int64_t)((out_nrows + 31u) / 32u) * 32;    /* :118 */
    if (gxn * out_ncols * out_nmats == 0) return ORC_OK;
    for (uint32_t gz = 0; gz < out_nmats; ++gz)
        for (uint32_t gy = 0; gy < out_ncols; ++gy) {
#pragma omp parallel for schedule(static)
            for (int64_t gx = 0; gx < gxn; ++gx) {
                switch (variant) {
                case GEMV: k_gemv(so, sm, sv, o, a, b, (uint32_t)gx, gy, gz); break;
                case GEMV_TR: k_gemv_tr(so, sm, sv, o, a, b, (uint32_t)gx, gy, gz); break;
                case GEMV_FAST: k_gemv_fast_wg(so, sm, sv, o, a, b, (uint32_t)gx, gy, gz, 0); break;
                default: k_gemv_fast_wg(so, sm, sv, o, a, b, (uint32_t)gx, gy, gz, 1); break;
                }
            }
        }
    return ORC_OK;
}

/* ---- op_assign.wgsl:14-47, op_assign.rs:82-93 ------------------------------------- */
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_COPY = 4 };

int orc_op_assign(int op, float *a, const shape_t *sa, const float *b, const shape_t *sb) {
    if (sa->nrows != sb->nrows) return ORC_DIM_MISMATCH;                   /* op_assign.rs:82-86 */
    const int64_t n = sa->nrows;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t ia = iv(*sa, (uint32_t)i), ib = iv(*sb, (uint32_t)i);
        const float x = a[ia], y = b[ib];
        float r;
        switch (op) {
        case OP_ADD: r = x + y; break;
        case OP_SUB: r = x - y; break;
        case OP_MUL: r = x * y; break;
        case OP_DIV: r = x / y; break;
        default: r = y; break;
        }
        a[ia] = r;
    }
    return ORC_OK;
}

/* ---- reduce.wgsl:12-96, reduce.rs:30-60,100-113 ------------------------------------ */
enum { RED_MIN = 0, RED_MAX = 1, RED_SUM = 2, RED_PROD = 3, RED_SQNORM = 4 };

static inline float red_ws(int op, float acc, float x) {    /* workspace_fn, reduce.rs:40-48 */
    switch (op) {
    case RED_MIN: return fminf(acc, x);
    case RED_MAX: return fmaxf(acc, x);
    case RED_SUM: return acc + x;
    case RED_PROD: return acc * x;
    default: return acc + x * x;
    }
}
static inline float red_tree(int op, float acc, float x) {  /* reduce_fn, reduce.rs:50-58 */
    switch (op) {
    case RED_MIN: return fminf(acc, x);
    case RED_MAX: return fmaxf(acc, x);
    case RED_PROD: return acc * x;
    default: return acc + x;                                 /* Sum and SqNorm */
    }
}
static inline float red_init(int op) {                       /* init_fn, reduce.rs:30-38 */
    switch (op) {
    case RED_MIN: return 3.4e38f;                            /* reduce.wgsl:40-42 */
    case RED_MAX: return -3.4e38f;                           /* :44-46 */
    case RED_PROD: return 1.0f;
    default: return 0.0f;
    }
}

int orc_reduce(int op, const float *input, const shape_t *s, float *result) {
    enum { WG = 128 };
    float ws[WG];
    for (uint32_t t = 0; t < WG; ++t) {
        float acc = red_init(op);
        for (uint32_t i = t; i < s->nrows; i += WG) acc = red_ws(op, acc, input[iv(*s, i)]);
        ws[t] = acc;
    }
    for (uint32_t stride = 64u; stride >= 1u; stride >>= 1)
        for (uint32_t t = 0; t < stride; ++t) ws[t] = red_tree(op, ws[t], ws[t + stride]);
    *result = ws[0];
    return ORC_OK;
}

/* ---- float64 references on the same views (error measurement only) ---------------- */
void orc_gemm_f64(int tr, double *out /* dense col-major M x N per mat */, uint32_t M, uint32_t N,
                  uint32_t K, uint32_t nmats, const float *m1, const shape_t *s1, const float *m2,
                  const shape_t *s2) {
    for (uint32_t t = 0; t < nmats; ++t) {
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < (int64_t)N; ++n)
            for (uint32_t i = 0; i < M; ++i) {
                double acc = 0.0;
                for (uint32_t k = 0; k < K; ++k) {
                    const float a = tr ? m1[it(*s1, k, i, t)] : m1[it(*s1, i, k, t)];
                    acc += (double)a * (double)m2[it(*s2, k, (uint32_t)n, t)];
                }
                out[((size_t)t * N + (size_t)n) * M + i] = acc;
            }
    }
}

/* ---- row-major views (shape.wgsl:49-57, the ROW_MAJOR branch) ----------------------- */
/* No linalg shader of the reference is compiled with row_major_shader_defs() (shape.rs:13-15), so there is no vec4 kernel
 * body to restate for it: what the reference defines is the ADDRESS of element (i, j, t) of a row-major view
 * (im(): offset + i * stride + j, it(): t * stride_mat + im).  These two functions are gemm.wgsl:81-148 / gemv.wgsl:68-115
 * written per scalar element on top of that addressing, with an ordering flag (0 column-major, 1 row-major) per operand:
 * f32 products summed over k in increasing order, no contraction.  They are the checker for wgb_gemm_ord / wgb_gemv_ord. */
static inline uint32_t it_ord(shape_t v, int row_major, uint32_t i, uint32_t j, uint32_t t) {
    return t * v.stride_mat + v.offset + (row_major ? i * v.stride + j : i + j * v.stride);
}

int orc_gemm_ord(int variant, float *out, const shape_t *so, int out_rm, const float *m1, const shape_t *s1, int m1_rm,
                 const float *m2, const shape_t *s2, int m2_rm) {
    int rc = gemm_check(*so, *s1, *s2, variant);
    if (rc) return rc;
    const int tr = variant == GEMM_TR || variant == GEMM_TR_FAST;
    const uint32_t K = s2->nrows;
    for (uint32_t t = 0; t < so->nmats; ++t) {
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < (int64_t)so->ncols; ++n)
            for (uint32_t i = 0; i < so->nrows; ++i) {
                float acc = 0.f;
                for (uint32_t k = 0; k < K; ++k) {
                    const float a = tr ? m1[it_ord(*s1, m1_rm, k, i, t)] : m1[it_ord(*s1, m1_rm, i, k, t)];
                    acc = acc + a * m2[it_ord(*s2, m2_rm, k, (uint32_t)n, t)];
                }
                out[it_ord(*so, out_rm, i, (uint32_t)n, t)] = acc;
            }
    }
    return ORC_OK;
}

int orc_gemv_ord(int variant, float *out, const shape_t *so, const float *m, const shape_t *sm, int m_rm, const float *v,
                 const shape_t *sv) {
    const int tr = variant == 2 || variant == 3;
    const uint32_t rows = tr ? sm->ncols : sm->nrows, cols = tr ? sm->nrows : sm->ncols;
    if (cols != sv->nrows || rows != so->nrows) return ORC_DIM_MISMATCH;   /* gemv.rs:89-90 */
    for (uint32_t t = 0; t < so->nmats; ++t)
        for (uint32_t c = 0; c < so->ncols; ++c) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < (int64_t)rows; ++i) {
                float acc = 0.f;
                for (uint32_t k = 0; k < cols; ++k) {
                    const float a = tr ? m[it_ord(*sm, m_rm, k, (uint32_t)i, t)] : m[it_ord(*sm, m_rm, (uint32_t)i, k, t)];
                    acc = acc + a * v[it_ord(*sv, 0, k, c, t)];
                }
                out[it_ord(*so, 0, (uint32_t)i, c, t)] = acc;
            }
        }
    return ORC_OK;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core it can. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
