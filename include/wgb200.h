/*
 * wgb200.h — C ABI of libwgebra_b200.so: the B200-native (sm_100a CUDA) backend for the
 * wgebra dense-linear-algebra surface of wgmath.
 *
 * This is the drop-in boundary.  Every entry point replaces one piece of what sits under
 * the Rust signatures of the reference (paths relative to /root/reference/crates/):
 *
 *   wgb_ctx_*            <-> wgcore/src/gpu.rs:15-79          GpuInstance (device + queue)
 *   wgb_pass_*           <-> wgcore/src/kernel.rs:7-27        CommandEncoderExt::compute_pass
 *   wgb_submit / sync    <-> queue.submit / device.poll(wait) (wgebra/src/linalg/gemm.rs:190-192,
 *                                                              wgcore/src/tensor.rs:304-312)
 *   wgb_buffer_*         <-> wgcore/src/tensor.rs:65-187      TensorBuilder::build / build_init
 *                            wgcore/src/tensor.rs:227-265     copy_from / copy_from_view
 *                            wgcore/src/tensor.rs:300-384     read_bytes / read_to / read
 *   wgb_view_shape       <-> wgcore/src/shapes.rs:9-21        ViewShape (byte-identical, 24 B)
 *   wgb_gemm             <-> wgebra/src/linalg/gemm.rs:65-127 Gemm::dispatch_generic
 *                            + gemm.wgsl:29-199 (the four kernels)
 *   wgb_gemv             <-> wgebra/src/linalg/gemv.rs:64-137 Gemv::dispatch_generic
 *                            + gemv.wgsl:29-154
 *   wgb_op_assign        <-> wgebra/src/linalg/op_assign.rs:71-94 + op_assign.wgsl:14-47
 *   wgb_reduce           <-> wgebra/src/linalg/reduce.rs:100-113  + reduce.wgsl:12-96
 *   wgb_event_*          <-> wgcore/src/timestamps.rs:9-248   GpuTimestamps
 *   wgb_prefix_sum       <-> wgrapier/src/dynamics/prefix_sum.rs:49-99 WgPrefixSum::dispatch + prefix_sum.wgsl:35-147
 *   wgb_radix_sort       <-> wgparry/src/utils/radix_sort/mod.rs:111-223 RadixSort::dispatch + sort_*.wgsl
 *                            (the integer scan / sort primitives next to the linalg path)
 *   wgb_geometry_batch   <-> wgebra/src/geometry/{cholesky,lu,qr2,qr3,qr4,eig2,eig3,eig4,svd2,svd3,inv}.wgsl applied as
 *                            `out[i] = f(in[i])`, the kernel every test of that module builds (cholesky.rs:53-63,
 *                            lu.rs:101-111, qr2.rs:36-46, eig3.rs:36-46, svd3.rs:34-44) — SURVEY.md §8(f) 4, first half
 *
 * Extensions that have no reference counterpart (named by BASELINE.json north_star):
 *   wgb_gemm_ex            bf16 operands / bf16 output, f32 compute-mode selection
 *   wgb_dot                sum_i a[i]*b[i]   (the reference only has SqNorm = x.x)
 *   wgb_reduce_columns     one launch for Reduce over every column of a matrix view
 *                          (reference: one Reduce dispatch per GpuMatrix::column(j))
 *   wgb_comm_*, wgb_gemm_row_sharded   8-GPU row-sharded GEMM + all-gather of C (NCCL / NVLink)
 *   wgb_gemm_ord, wgb_gemv_ord   per-operand RowMajor / ColumnMajor ordering (tensor.rs:17-39, shape.wgsl:49-57)
 *   wgb_fill_uniform       seeded synthetic inputs generated in HBM (bench / tests)
 *
 * Conventions
 *   - Plain C: opaque handles, POD structs, pointers and sizes.  No torch / C++ types.
 *   - Every function returns a wgb_status.  The reference *panics* on a dimension
 *     mismatch (assert_eq! at gemm.rs:91-95, gemv.rs:89-90,122, op_assign.rs:82-86);
 *     here that is WGB_ERR_DIM_MISMATCH and the host shim turns it back into a panic /
 *     exception.  The reference silently *skips* a dispatch whose grid is empty or that
 *     binds a zero-sized buffer (kernel.rs:111-113,121-123,144); here that is WGB_OK
 *     with nothing enqueued.
 *   - Matrices are column-major; all shape fields are u32 *element* counts exactly as in
 *     ViewShape.  Kernels widen to 64-bit internally.
 *   - Work is enqueued asynchronously on the pass's CUDA stream, in order (wgpu executes
 *     the dispatches of a pass in order with storage hazards resolved, so one pass / one
 *     encoder <-> one stream).  Nothing is guaranteed complete until wgb_ctx_sync,
 *     wgb_buffer_read or wgb_event_elapsed_ms.
 *   - There is no CPU fallback: every entry point fails with WGB_ERR_CUDA /
 *     WGB_ERR_NO_DEVICE when no sm_100 device is usable.
 *   - Superset rule: the reference kernels are undefined unless every dimension, stride
 *     and offset is a multiple of 4 (shape.wgsl:64-66) and K is a multiple of 256 / 128
 *     for the *_fast variants (gemm.wgsl:40-41, gemv.wgsl:39-40).  This library computes
 *     the mathematically defined result for every well-formed view, and never touches
 *     memory outside the view.
 */
#ifndef WGB200_H
#define WGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGB200_ABI_VERSION 1

typedef struct wgb_ctx wgb_ctx;       /* device + in-order queue  (GpuInstance)            */
typedef struct wgb_pass wgb_pass;     /* command encoder + compute pass => a CUDA stream  */
typedef struct wgb_buffer wgb_buffer; /* wgpu::Buffer                                      */
typedef struct wgb_event wgb_event;   /* one GPU timestamp                                 */

typedef enum wgb_status {
    WGB_OK = 0,
    WGB_ERR_INVALID = 1,       /* null handle, unknown enum value, misuse                      */
    WGB_ERR_DIM_MISMATCH = 2,  /* the reference's assert_eq! "dimension mismatch" panics       */
    WGB_ERR_CUDA = 3,          /* a CUDA runtime / driver call failed                          */
    WGB_ERR_UNSUPPORTED = 4,   /* dtype / variant combination not implemented                  */
    WGB_ERR_OOM = 5,           /* device or pinned-host allocation failed                      */
    WGB_ERR_NCCL = 6,          /* NCCL missing or a collective failed                          */
    WGB_ERR_OUT_OF_BOUNDS = 7, /* a view reaches past the end of its buffer                    */
    WGB_ERR_NO_DEVICE = 8      /* no CUDA device with compute capability 10.x                  */
} wgb_status;

/* crates/wgcore/src/shapes.rs:9-21 — identical field order and size (24 bytes). */
typedef struct wgb_view_shape {
    uint32_t size[3];    /* rows, columns, matrices                                  */
    uint32_t stride;     /* elements between two columns                             */
    uint32_t stride_mat; /* elements between two matrices                            */
    uint32_t offset;     /* index of the first element in the underlying buffer      */
} wgb_view_shape;

typedef enum wgb_dtype { WGB_F32 = 0, WGB_BF16 = 1 } wgb_dtype;

/* wgebra/src/linalg/gemm.rs:25-35 (same order as the Rust enum) */
typedef enum wgb_gemm_variant {
    WGB_GEMM = 0,
    WGB_GEMM_FAST = 1,
    WGB_GEMM_TR = 2,
    WGB_GEMM_TR_FAST = 3
} wgb_gemm_variant;

/* wgebra/src/linalg/gemv.rs:24-34 */
typedef enum wgb_gemv_variant {
    WGB_GEMV = 0,
    WGB_GEMV_FAST = 1,
    WGB_GEMV_TR = 2,
    WGB_GEMV_TR_FAST = 3
} wgb_gemv_variant;

/* wgebra/src/linalg/op_assign.rs:15-26 */
typedef enum wgb_op_assign_variant {
    WGB_OP_ADD = 0,
    WGB_OP_SUB = 1,
    WGB_OP_MUL = 2,
    WGB_OP_DIV = 3,
    WGB_OP_COPY = 4
} wgb_op_assign_variant;

/* wgebra/src/linalg/reduce.rs:16-27 */
typedef enum wgb_reduce_op {
    WGB_RED_MIN = 0,
    WGB_RED_MAX = 1,
    WGB_RED_SUM = 2,
    WGB_RED_PROD = 3,
    WGB_RED_SQNORM = 4
} wgb_reduce_op;

/* wgebra/src/geometry/mod.rs:3-17: the factorization libraries (WgCholesky2..4, WgLU2..4, WgQR2..4, WgSymmetricEigen2..4,
 * WgSvd2 / WgSvd3, WgInv) */
typedef enum wgb_geom_op {
    WGB_GEOM_CHOLESKY = 0,
    WGB_GEOM_LU = 1,
    WGB_GEOM_QR = 2,
    WGB_GEOM_SYMMETRIC_EIGEN = 3,
    WGB_GEOM_SVD = 4, /* 2x2 and 3x3 only, like the reference */
    WGB_GEOM_INV = 5
} wgb_geom_op;

/* How an f32 x f32 GEMM is computed on the tensor cores. */
typedef enum wgb_f32_mode {
    WGB_F32_AUTO = 0,   /* 3xTF32 when the views are TMA-eligible, else SIMT FFMA        */
    WGB_F32_3XTF32 = 1, /* error-compensated hi/lo split, f32-level accuracy (parity gate)*/
    WGB_F32_TF32 = 2,   /* single-pass TF32: fast, ~1e-4 relative (informational)        */
    WGB_F32_SIMT = 3    /* FFMA on CUDA cores: exact f32 products                         */
} wgb_f32_mode;

/* wgpu::BufferUsages bit values, so a Rust shim can pass `usage.bits()` straight through. */
enum {
    WGB_USAGE_MAP_READ = 1 << 0, /* host-visible staging buffer: allocated as pinned host memory */
    WGB_USAGE_MAP_WRITE = 1 << 1,
    WGB_USAGE_COPY_SRC = 1 << 2,
    WGB_USAGE_COPY_DST = 1 << 3,
    WGB_USAGE_UNIFORM = 1 << 6,
    WGB_USAGE_STORAGE = 1 << 7
};

/* ------------------------------------------------------------------ library ---------- */
int wgb_abi_version(void);
/* Thread-local, human-readable description of the last non-OK status on this thread. */
const char *wgb_last_error_string(void);

/* ------------------------------------------------------------------ context ---------- */
/* gpu.rs:15-58 GpuInstance::new: pick the device, create its in-order queue (stream). */
wgb_status wgb_ctx_create(int device_ordinal, wgb_ctx **out);
/* Drains the queue and drops the caller's reference.  Like wgpu's handles (every Buffer keeps its Device alive), buffers,
 * events, passes, graphs and peer groups created from the context each hold a reference: they stay valid, and may be destroyed,
 * after wgb_ctx_destroy; the context's own resources are released with the last of them. */
wgb_status wgb_ctx_destroy(wgb_ctx *ctx);
/* device.poll(PollType::wait()) — tensor.rs:304-312: block until all submitted work is done. */
wgb_status wgb_ctx_sync(wgb_ctx *ctx);
wgb_status wgb_ctx_device_info(wgb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                               size_t *total_mem_bytes, char *name, size_t name_len);
/* Number of kernels this library has launched on this context (bench.py's gpu_launches). */
wgb_status wgb_ctx_launch_count(wgb_ctx *ctx, uint64_t *count);
/* The context's queue as a cudaStream_t (for interop with a host framework's allocator). */
wgb_status wgb_ctx_stream(wgb_ctx *ctx, void **cuda_stream);

/* ------------------------------------------------------------------ passes ----------- */
/* kernel.rs:15-26 compute_pass(label, timestamps): begin recording on the queue.  `begin_ts`
 * / `end_ts` may be NULL; when given they are recorded at pass begin / end (timestamps.rs:63-70). */
wgb_status wgb_pass_begin(wgb_ctx *ctx, const char *label, wgb_event *begin_ts, wgb_event *end_ts,
                          wgb_pass **out);
/* drop(pass): ends the pass.  The handle is invalid afterwards. */
wgb_status wgb_pass_end(wgb_pass *pass);
/* queue.submit(encoder.finish()): CUDA work is already in flight, this only flushes. */
wgb_status wgb_submit(wgb_ctx *ctx);

/* Record-once / replay-many (CUDA graphs): between capture_begin and capture_end every dispatch on the context's
 * queue is recorded instead of executed (like recording a wgpu command buffer); wgb_graph_launch replays the whole
 * recorded sequence with one launch, removing per-dispatch host cost for chains of small kernels.  Workspaces
 * must be warm (run the same sequence once before capturing): an allocation during capture fails with WGB_ERR_CUDA. */
typedef struct wgb_graph wgb_graph;
wgb_status wgb_graph_capture_begin(wgb_ctx *ctx);
wgb_status wgb_graph_capture_end(wgb_ctx *ctx, wgb_graph **out);
wgb_status wgb_graph_launch(wgb_graph *graph);
wgb_status wgb_graph_destroy(wgb_graph *graph);

/* ------------------------------------------------------------------ buffers ---------- */
/* tensor.rs:112-129 TensorBuilder::build: uninitialised buffer of `bytes` bytes. */
wgb_status wgb_buffer_create(wgb_ctx *ctx, size_t bytes, uint32_t usage, wgb_buffer **out);
/* tensor.rs:149-161 build_bytes / :175-186 build_init: create + upload from host memory. */
wgb_status wgb_buffer_create_init(wgb_ctx *ctx, const void *host_data, size_t bytes, uint32_t usage,
                                  wgb_buffer **out);
/* Wrap device memory owned by the caller (not freed on destroy). */
wgb_status wgb_buffer_wrap(wgb_ctx *ctx, void *device_ptr, size_t bytes, wgb_buffer **out);
wgb_status wgb_buffer_destroy(wgb_buffer *buf);
wgb_status wgb_buffer_size(const wgb_buffer *buf, size_t *bytes);
wgb_status wgb_buffer_device_ptr(const wgb_buffer *buf, void **ptr);
/* queue.write_buffer: asynchronous host -> device copy on the queue (host memory must stay
 * valid until the next sync; pinned memory makes it truly asynchronous). */
wgb_status wgb_buffer_write(wgb_ctx *ctx, wgb_buffer *dst, size_t dst_offset_bytes,
                            const void *host_src, size_t bytes);
/* tensor.rs:227-233 copy_from / :244-265 copy_from_view: buffer-to-buffer copy on the pass's
 * stream (pass may be NULL: the copy is then recorded on the context queue, as
 * CommandEncoder::copy_buffer_to_buffer is outside any compute pass). */
wgb_status wgb_buffer_copy(wgb_ctx *ctx, wgb_pass *pass, wgb_buffer *dst, size_t dst_offset_bytes,
                           const wgb_buffer *src, size_t src_offset_bytes, size_t bytes);
/* tensor.rs:300-384 read / read_to: blocks until all queued work is done, then copies out. */
wgb_status wgb_buffer_read(wgb_ctx *ctx, const wgb_buffer *src, size_t src_offset_bytes,
                           void *host_dst, size_t bytes);
/* Pinned host memory for callers that want asynchronous uploads / downloads. */
wgb_status wgb_host_alloc(size_t bytes, void **out);
wgb_status wgb_host_free(void *ptr);

/* ------------------------------------------------------------------ operators -------- */
/* gemm.rs:65-127.  out = m1 * m2 (GEMM, GEMM_FAST) or tr(m1) * m2 (GEMM_TR, GEMM_TR_FAST),
 * overwrite, batched over size[2].  f32 operands and output; equivalent to
 * wgb_gemm_ex(.., WGB_F32, WGB_F32, WGB_F32_AUTO).  The four variants give the same result
 * (they differ only in the reference's launch shape); the variant selects tr / non-tr. */
wgb_status wgb_gemm(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out,
                    const wgb_view_shape *out_shape, const wgb_buffer *m1,
                    const wgb_view_shape *m1_shape, const wgb_buffer *m2,
                    const wgb_view_shape *m2_shape);
/* in_dtype: element type of m1 and m2; out_dtype: element type of out (accumulation is f32). */
wgb_status wgb_gemm_ex(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out,
                       const wgb_view_shape *out_shape, const wgb_buffer *m1,
                       const wgb_view_shape *m1_shape, const wgb_buffer *m2,
                       const wgb_view_shape *m2_shape, wgb_dtype in_dtype, wgb_dtype out_dtype,
                       wgb_f32_mode f32_mode);
/* GEMM with the caller's next element-wise step fused into the epilogue: out = (m1 * m2) (op) operand,
 * i.e. Gemm::dispatch followed by OpAssign::dispatch(out, operand) (op_assign.rs:71-94) without the extra round trip of
 * `out` through HBM.  `operand` is a matrix view of out's element type with out's [rows, cols, mats]; its strides are free.
 * op == WGB_OP_COPY is rejected (it would discard the product). */
wgb_status wgb_gemm_op(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *out_shape,
                       const wgb_buffer *m1, const wgb_view_shape *m1_shape, const wgb_buffer *m2,
                       const wgb_view_shape *m2_shape, wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode f32_mode,
                       wgb_op_assign_variant op, const wgb_buffer *operand, const wgb_view_shape *operand_shape);

/* Matrix ordering of a view: wgcore/src/tensor.rs:17-39 (ColumnMajor / RowMajor: MatrixOrdering, the type parameter of
 * GpuTensorView) and the ROW_MAJOR branch of wgebra/src/linalg/shape.wgsl:49-57 switched on by row_major_shader_defs()
 * (shape.rs:13-15): element (i, j) of a row-major view lives at offset + i * stride + j, `stride` being the distance between
 * consecutive ROWS; size[] stays [rows, cols, mats] and stride_mat the distance between matrices.  No linalg shader of the
 * reference enables the define, so these entry points are the defined-by-shape.wgsl superset: every
 * combination of orderings is computed in place (no copy of a bf16 operand; an N-contiguous f32 m2 is transposed by the
 * 3xTF32 split that re-materialises it anyway). */
typedef enum wgb_ordering { WGB_COLUMN_MAJOR = 0, WGB_ROW_MAJOR = 1 } wgb_ordering;

/* wgb_gemm_ex / wgb_gemm_op with an ordering per operand.  op < 0: no fused element-wise step (operand ignored); otherwise
 * `operand` is a view ordered like `out`. */
wgb_status wgb_gemm_ord(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out, const wgb_view_shape *out_shape,
                        wgb_ordering out_ordering, const wgb_buffer *m1, const wgb_view_shape *m1_shape,
                        wgb_ordering m1_ordering, const wgb_buffer *m2, const wgb_view_shape *m2_shape,
                        wgb_ordering m2_ordering, wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode f32_mode, int op,
                        const wgb_buffer *operand, const wgb_view_shape *operand_shape);

/* Host-buffer GEMM: out_host = m1_host * m2_host (or tr(m1_host) * m2_host) with dense column-major host matrices
 * (leading dimensions = row counts).  Equivalent to build_init(m1), build_init(m2), dispatch, read — the sequence of
 * the reference's own tests (gemm.rs:156-193) — but pipelined: m2 is uploaded and the product downloaded in column
 * panels, so the PCIe download of panel j overlaps the upload of panel j+1 and the tensor-core work hides under both.
 * Blocking (returns when out_host is complete).  Pinned host memory (wgb_host_alloc) is needed for the overlap;
 * pageable memory works but serialises.  n_panels <= 0 selects the default. */
wgb_status wgb_gemm_host(wgb_ctx *ctx, wgb_gemm_variant variant, uint32_t M, uint32_t N, uint32_t K, void *out_host,
                         const void *m1_host, const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype,
                         wgb_f32_mode f32_mode, int n_panels);

/* The same product, enqueued: returns once the copies and kernels are queued (the wgpu model: queue.submit now, map_async /
 * poll later — tensor.rs:300-384).  out_host is complete after wgb_ctx_sync(), or — for work recorded later on the queue,
 * e.g. a timestamp — after wgb_gemm_host_flush(), which makes the queue wait (on the device, without blocking the host) for
 * every enqueued product's download.  Two device operand slots alternate, so the download of product i overlaps the upload of
 * product i+1; the host buffers of a product must stay untouched until it has completed. */
wgb_status wgb_gemm_host_enqueue(wgb_ctx *ctx, wgb_gemm_variant variant, uint32_t M, uint32_t N, uint32_t K, void *out_host,
                                 const void *m1_host, const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype,
                                 wgb_f32_mode f32_mode, int n_panels);
wgb_status wgb_gemm_host_flush(wgb_ctx *ctx);

/* WgPrefixSum::dispatch (wgrapier/src/dynamics/prefix_sum.rs:49-99): in-place EXCLUSIVE prefix sum of a u32 vector view,
 * data[i] <- data[0] + ... + data[i-1] (wrapping u32 adds, data[0] <- 0).  The reference's PrefixSumWorkspace (aux levels,
 * prefix_sum.rs:119-224) has no counterpart: the look-back descriptors live in the context.  An empty view is a no-op (the
 * reference's workspace sizing does not terminate for length 0, prefix_sum.rs:188-206). */
wgb_status wgb_prefix_sum(wgb_pass *pass, wgb_buffer *data, const wgb_view_shape *data_shape);

/* RadixSort::dispatch (wgparry/src/utils/radix_sort/mod.rs:111-223): the first n pairs of (input_keys, input_values), n =
 * min(*n_sort, length) with n_sort a device-resident u32 (GpuScalar<u32>; element 0 of the buffer), are written to
 * output_keys / output_values stably ordered by the low 4 * ceil(sorting_bits / 4) bits of the key.  Entries past n and the
 * inputs are left untouched; sorting_bits == 0 runs no pass (outputs untouched), sorting_bits > 32 is WGB_ERR_INVALID (the
 * reference's assert).  The key and value views must have equal lengths (its assert_eq!, WGB_ERR_DIM_MISMATCH), the output
 * views at least that length, and outputs must not overlap inputs.  Ping-pong buffers (RadixSortWorkspace, mod.rs:82-109)
 * live in the context. */
wgb_status wgb_radix_sort(wgb_pass *pass, const wgb_buffer *input_keys, const wgb_view_shape *input_keys_shape,
                          const wgb_buffer *input_values, const wgb_view_shape *input_values_shape, const wgb_buffer *n_sort,
                          uint32_t sorting_bits, wgb_buffer *output_keys, const wgb_view_shape *output_keys_shape,
                          wgb_buffer *output_values, const wgb_view_shape *output_values_shape);

/* Batched small-matrix factorizations: out[i] = f(in[i]) for i < n, one dim x dim f32 matrix per element (dim = 2, 3, 4).
 * In the reference these are WGSL function libraries other shaders import (geometry/mod.rs:3-17); the batched form is the
 * kernel each of its tests builds around them.  Elements use WGSL's storage layout, which is what the reference's Rust side
 * uploads and reads back (Matrix2 / Matrix4x3 / Matrix4, GpuLU*, GpuQR*, GpuSymmetricEigen*, GpuSvd2, GpuSvd3):
 *   matrix  = dim columns of CS floats, CS = 2 (dim 2) or 4 (dim 3 and 4; the 4th float of a vec3 column is padding)
 *   CHOLESKY, INV      -> matrix                                   16 / 48 / 64 bytes   (cholesky.wgsl:16-35, inv.wgsl:8-88)
 *   LU                 -> {lu: matrix, ia: uvecN, ib: uvecN, len}   40 / 80 / 112 bytes  (lu.wgsl:12-81; lu.rs:25-56)
 *   QR                 -> {q: matrix, r: matrix}                    32 / 96 / 128 bytes  (qr2.wgsl:7-107; qr3.rs:15-20)
 *   SYMMETRIC_EIGEN    -> {eigenvectors: matrix, eigenvalues: vecN} 24 / 64 / 80 bytes   (eig2.wgsl:7-40, eig3.wgsl:13-160)
 *   SVD (dim 2, 3)     -> {U: matrix, S: vecN, Vt: matrix}          40 / 112 bytes       (svd2.wgsl:5-39, svd3.wgsl:12-305)
 * Padding words are written as zero.  in_first / out_first are element indices into the buffers (a GpuVector view's offset).
 * In place (same buffer, same first element) is allowed for CHOLESKY and INV; any other overlap is WGB_ERR_INVALID.
 * n == 0 or a zero-sized buffer enqueues nothing (kernel.rs:111-113,144).  The QR sweeps of eig3 / eig4 are unbounded in the
 * reference (eig3.wgsl:77); here they stop after 256 sweeps.  Results are the reference's, quirks included: WGSL's sign(0) = 0 zeroes
 * eigenvector rows for inputs with an exactly trivial Householder step (diagonal / block-diagonal matrices, eig3.wgsl:48-67); the
 * eigenvalues are correct in every case. */
uint32_t wgb_geometry_in_bytes(int dim);
uint32_t wgb_geometry_out_bytes(wgb_geom_op op, int dim);
wgb_status wgb_geometry_batch(wgb_pass *pass, wgb_geom_op op, int dim, const wgb_buffer *in, uint64_t in_first, wgb_buffer *out,
                              uint64_t out_first, uint64_t n);

/* Diagnostics (no reference counterpart): per-cluster timeline of the most recent tcgen05 GEMM launch on this context.
 * enable != 0 switches tracing on for later launches (a few global stores per CTA); out, if non-null, receives up to
 * max_records records of 8 x u64 (one per CTA cluster, in cluster order) after synchronising the queue:
 *   [0] globaltimer ns at kernel entry      [1] ns after the programmatic-dependent-launch wait
 *   [2] ns when the first operand stage landed (first MMA issued)      [3] ns when the last MMA was issued
 *   [4] SM clock cycles between [2] and [3]  [5] k-blocks issued by this cluster  [6] ns when the leader CTA's epilogue finished
 *   [7] work units processed.  enable == 0 frees the trace buffer. */
wgb_status wgb_debug_tc_trace(wgb_ctx *ctx, int enable, unsigned long long *out, size_t max_records, size_t *n_records);

/* Gemm -> Reduce along one axis, fused: result[j] = reduce_op_i (m1 * m2)[i, j] (axis 1: one value per column,
 * what one Reduce::dispatch per GpuMatrix::column(j) of the product gives, reduce.rs:100-113 + tensor.rs:574-585) or
 * result[i] = reduce_op_j (m1 * m2)[i, j] (axis 2).  The product is never stored (no HBM round trip of C): the tensor-core epilogue
 * leaves per-32-row / per-32-column partial results that a fold kernel combines in index order — deterministic, and equal to the
 * two-dispatch chain up to f32 rounding (not bit-identical: the chain's reduction tree spans the whole column).  f32 results; f32 or
 * bf16 operands; one matrix per operand.  Products the tensor-core path does not take are stored in a scratch matrix and reduced. */
wgb_status wgb_gemm_reduce(wgb_pass *pass, wgb_gemm_variant variant, int axis, wgb_reduce_op reduce_op, wgb_buffer *result,
                           const wgb_view_shape *result_shape, const wgb_buffer *m1, const wgb_view_shape *m1_shape,
                           const wgb_buffer *m2, const wgb_view_shape *m2_shape, wgb_dtype in_dtype, wgb_f32_mode f32_mode);

/* Gemv -> Reduce fused: result = reduce_op over the elements of m * v (or tr(m) * v), the reference's
 * Gemv::dispatch + Reduce::dispatch chain (gemv.rs:64-137, reduce.rs:100-113) in one launch.  The product vector is never written to
 * a caller buffer; the scalar is bit-identical to the two-dispatch chain through a 16-byte aligned `out`.  One matrix, one vector. */
wgb_status wgb_gemv_reduce(wgb_pass *pass, wgb_gemv_variant variant, wgb_reduce_op reduce_op, wgb_buffer *result,
                           const wgb_buffer *m, const wgb_view_shape *m_shape, wgb_ordering m_ord, const wgb_buffer *v,
                           const wgb_view_shape *v_shape);

/* Which kernel family the last wgb_gemm* call on this pass dispatched to:
 * 0 none, 1 SIMT FFMA, 2 tcgen05 bf16, 3 tcgen05 tf32, 4 tcgen05 3xtf32. */
wgb_status wgb_pass_last_gemm_path(const wgb_pass *pass, int *path);
/* The tcgen05 kernel instantiation and plan of that call (all zero unless the path is 2..4), WGB_TC_CONFIG_WORDS ints:
 *   [0] operand kind (0 bf16, 1 tf32)  [1] A MN-major  [2] B MN-major  [3] BLOCK_N  [4] MMA passes (3 = 3xTF32)
 *   [5] output wgb_dtype  [6] CTAs per tile (cta_group)  [7] epilogue (0 per-lane stores, 1 TMA bulk stores)
 *   [8] tail N-split factor  [9] tail split-K factor  [10] destinations per output block (fused all-gather: ranks)  [11] work units
 *   [12] 3xTF32 operand split inside the GEMM (1) or by the split kernels (0) */
#define WGB_TC_CONFIG_WORDS 13
wgb_status wgb_pass_last_gemm_config(const wgb_pass *pass, int *config /* WGB_TC_CONFIG_WORDS */);

/* gemv.rs:64-137.  out = m * v or tr(m) * v; v / out may carry several columns and batches. */
wgb_status wgb_gemv(wgb_pass *pass, wgb_gemv_variant variant, wgb_buffer *out,
                    const wgb_view_shape *out_shape, const wgb_buffer *m,
                    const wgb_view_shape *m_shape, const wgb_buffer *v,
                    const wgb_view_shape *v_shape);

/* wgb_gemv with a row- or column-major m (v and out stay column-major: a contiguous vector is the same in both). */
wgb_status wgb_gemv_ord(wgb_pass *pass, wgb_gemv_variant variant, wgb_buffer *out, const wgb_view_shape *out_shape,
                        const wgb_buffer *m, const wgb_view_shape *m_shape, wgb_ordering m_ordering, const wgb_buffer *v,
                        const wgb_view_shape *v_shape);

/* GEMV with the caller's next element-wise step fused into the store: out = (m * v) (op) operand, i.e.
 * Gemv::dispatch followed by OpAssign::dispatch(out, operand) (gemv.rs:64-137, op_assign.rs:71-94) as one launch.  `operand` is
 * an f32 view with out's rows and at least out's columns / matrices; it may be the output view itself (out = m * v + out, the
 * residual update), any other overlap is WGB_ERR_INVALID.  op < 0: plain wgb_gemv_ord (operand ignored); WGB_OP_COPY is rejected
 * (it would discard the product).  A row mismatch is the reference's "Op-assign: dimension mismatch." panic. */
wgb_status wgb_gemv_op(wgb_pass *pass, wgb_gemv_variant variant, wgb_buffer *out, const wgb_view_shape *out_shape,
                       const wgb_buffer *m, const wgb_view_shape *m_shape, wgb_ordering m_ordering, const wgb_buffer *v,
                       const wgb_view_shape *v_shape, int op, const wgb_buffer *operand, const wgb_view_shape *operand_shape);

/* op_assign.rs:71-94.  a[i] = a[i] (op) b[i]; only size[0] and offset of the shapes are used
 * (shape.wgsl:36-38 iv()). */
wgb_status wgb_op_assign(wgb_pass *pass, wgb_op_assign_variant op, wgb_buffer *in_out_a,
                         const wgb_view_shape *a_shape, const wgb_buffer *in_b,
                         const wgb_view_shape *b_shape);

/* reduce.rs:100-113.  result[0] = reduce(op, value[offset .. offset+size[0]]). */
wgb_status wgb_reduce(wgb_pass *pass, wgb_reduce_op op, const wgb_buffer *value,
                      const wgb_view_shape *value_shape, wgb_buffer *result);

/* Extension: result[0] = sum_i a[i] * b[i]. */
wgb_status wgb_dot(wgb_pass *pass, const wgb_buffer *a, const wgb_view_shape *a_shape,
                   const wgb_buffer *b, const wgb_view_shape *b_shape, wgb_buffer *result);

/* Extension: out[offset + j] = reduce(op, column j of matrix t) for every column of the view,
 * in one launch (out is a vector view with size[0] == m.size[1] * m.size[2]). */
wgb_status wgb_reduce_columns(wgb_pass *pass, wgb_reduce_op op, const wgb_buffer *m,
                              const wgb_view_shape *m_shape, wgb_buffer *out,
                              const wgb_view_shape *out_shape);

/* Seeded U[0,1) fill of a matrix view: element (i, j) of matrix t gets
 * f(seed, row0 + i, col0 + j + t * size[1]); bit-identical to oracle.uniform().  dtype selects
 * f32 or bf16 (round-to-nearest-even of the f32 stream). */
wgb_status wgb_fill_uniform(wgb_pass *pass, wgb_buffer *buf, const wgb_view_shape *shape,
                            wgb_dtype dtype, uint64_t seed, uint32_t row0, uint32_t col0);

/* ------------------------------------------------------------------ timestamps ------- */
/* timestamps.rs: one wgb_event <-> one timestamp slot; elapsed <-> wait_for_results_ms. */
wgb_status wgb_event_create(wgb_ctx *ctx, wgb_event **out);
wgb_status wgb_event_destroy(wgb_event *ev);
wgb_status wgb_event_record(wgb_event *ev, wgb_pass *pass /* NULL: context queue */);
wgb_status wgb_event_elapsed_ms(wgb_event *begin, wgb_event *end, float *ms); /* blocks on `end` */

/* ------------------------------------------------------------------ multi-GPU -------- */
/* One process per GPU.  Rank 0 calls wgb_comm_get_unique_id and distributes the 128-byte id
 * (e.g. through torch.distributed); every rank then calls wgb_comm_init_rank. */
#define WGB_COMM_ID_BYTES 128
wgb_status wgb_comm_get_unique_id(void *id_out /* WGB_COMM_ID_BYTES */);
wgb_status wgb_comm_init_rank(wgb_ctx *ctx, int nranks, int rank, const void *id);
wgb_status wgb_comm_destroy(wgb_ctx *ctx);
/* Row-sharded GEMM: rank p owns rows [p*M/P, (p+1)*M/P) of m1 and of the
 * product.  m1_local is the local [M/P, K] row block (or [K, M/P] for the TR variants), m2 the
 * full [K, N] matrix (replicated).  out_gathered receives all P row-block panels back to back:
 * a cube view size = [M/P, N, P], stride = M/P, stride_mat = (M/P)*N, i.e. panel p is rows
 * [p*M/P, ..) of the product.  The all-gather is chunked by column panel and overlapped with
 * the remaining tiles of the local GEMM.  `n_chunks` <= 0 selects the default. */
wgb_status wgb_gemm_row_sharded(wgb_pass *pass, wgb_gemm_variant variant, wgb_buffer *out_gathered,
                                const wgb_buffer *m1_local, const wgb_view_shape *m1_local_shape,
                                const wgb_buffer *m2, const wgb_view_shape *m2_shape,
                                wgb_dtype in_dtype, wgb_dtype out_dtype, wgb_f32_mode f32_mode,
                                int n_chunks);

/* ---- fused GEMM + all-gather over peer memory (NVLink / NVSwitch), no NCCL on the data path ------------
 * Every rank owns a "gathered" buffer that its peers map through CUDA IPC.  The GEMM epilogue stores each
 * output element straight into the gathered buffer of every rank (local + P-1 peer stores over NVLink, overlapped
 * tile by tile with the tensor-core main loop); the last CTA then publishes a completion flag to each peer.
 * Set-up (once): create on every rank, export the 64-byte handle, exchange all handles through the launcher,
 * connect.  Every rank must call wgb_gemm_row_sharded_fused the same number of times (it is a collective). */
#define WGB_IPC_HANDLE_BYTES 64
typedef struct wgb_peer_gather wgb_peer_gather;
wgb_status wgb_peer_gather_create(wgb_ctx *ctx, int nranks, int rank, size_t gathered_bytes, wgb_peer_gather **out);
/* `depth` gathered buffers (1..3) that successive calls rotate through (call e uses buffer e mod depth):
 *   1  one buffer; a peer can start storing step e only when this rank has itself reached step e (lock step);
 *   2  peers may store step e + 1 as soon as this rank has started step e; the result of call e is valid until call e + 1 is issued;
 *   3  same look-ahead; the result of call e stays valid until call e + 2 is issued, so with WGB_GATHER_NO_WAIT the wait for (and
 *      the use of) step e may be queued after the GEMM of step e + 1: a slow rank no longer stalls the others at every step.
 * All ranks must use the same depth. */
wgb_status wgb_peer_gather_create_ex(wgb_ctx *ctx, int nranks, int rank, size_t gathered_bytes, int depth, wgb_peer_gather **out);
/* A group over memory the CALLER allocated symmetrically on every rank and mapped both ways: peer_bases[q] = this process's mapping
 * of rank q's region (peer_bases[rank] = the local one), multicast_base = the same region through an NVSwitch multicast object
 * bound on all ranks, or NULL.  With a multicast mapping the fused epilogue stores every output block ONCE (multimem.st) and the
 * switch delivers it to all ranks; without, it stores once per rank.  The region holds wgb_peer_gather_region_bytes() bytes on
 * every rank; the library never frees or unmaps it.  The ranks must synchronise (a barrier of their launcher) between creating
 * the group and the first product.  (cuMemCreate / cuMulticast* directly, or torch's symmetric memory in the Python mirror.) */
size_t wgb_peer_gather_region_bytes(size_t gathered_bytes, int depth);
wgb_status wgb_peer_gather_create_external(wgb_ctx *ctx, int nranks, int rank, size_t gathered_bytes, int depth,
                                           void *const *peer_bases /* nranks */, void *multicast_base, wgb_peer_gather **out);
/* Same-process connect: groups[q] is rank q's group (one context per rank; the ranks may share a device — the whole protocol then
 * runs on a single-GPU box — or sit on several devices driven by one process).  groups[rank] must be pg itself. */
wgb_status wgb_peer_gather_connect_local(wgb_peer_gather *pg, wgb_peer_gather *const *groups /* nranks */);
/* The gathered buffer written by the call `calls_back` calls ago (0 = the most recent), while it is still valid (see depth). */
wgb_status wgb_peer_gather_buffer_at(wgb_peer_gather *pg, int calls_back, wgb_buffer **out);
/* Queue a wait for the panels of every peer of the call `calls_back` calls ago (what wgb_gemm_row_sharded_fused does itself
 * unless WGB_GATHER_NO_WAIT is given).  Later work on the pass's queue sees the complete gathered result of that call. */
wgb_status wgb_peer_gather_wait(wgb_pass *pass, wgb_peer_gather *pg, int calls_back);
wgb_status wgb_peer_gather_export(wgb_peer_gather *pg, void *handle_out /* WGB_IPC_HANDLE_BYTES */);
wgb_status wgb_peer_gather_connect(wgb_peer_gather *pg, const void *handles /* nranks x WGB_IPC_HANDLE_BYTES */);
/* The local gathered buffer as a wgb_buffer (owned by the group; valid until wgb_peer_gather_destroy). */
wgb_status wgb_peer_gather_buffer(wgb_peer_gather *pg, wgb_buffer **out);
/* Diagnostics: link micro-benchmark.  The SMs stream `bytes` of `src` to ndst destinations `iters` times; mode 0 = st.global.v4 to
 * unicast addresses (local or peer mappings), 1 = multimem.st to one multicast address (dsts[0]), 2 = TMA bulk stores (16 KiB) to
 * unicast addresses.  ctas <= 0: two CTAs per SM.  Returns the time of one pass. */
wgb_status wgb_debug_link_stream(wgb_ctx *ctx, int mode, void *const *dsts, int ndst, const void *src, size_t bytes, int ctas,
                                 int iters, float *ms_per_iter);
/* Diagnostics: the flag block of this rank read on a private stream (usable while the queues are stuck behind a missing peer):
 * out[0..8) ready[q], out[8..16) done[q], out[16] CTA counter, out[17] calls made on this rank. */
wgb_status wgb_peer_gather_debug_flags(wgb_peer_gather *pg, unsigned int *out /* 18 words */);
/* Tear-down across processes is two-phase (CUDA IPC: importers unmap before the exporter frees): every rank disconnects, the
 * ranks meet at a barrier of their launcher, every rank destroys.  wgb_peer_gather_destroy alone disconnects first. */
wgb_status wgb_peer_gather_disconnect(wgb_peer_gather *pg);
wgb_status wgb_peer_gather_destroy(wgb_peer_gather *pg);
/* Same contract and output layout as wgb_gemm_row_sharded, the all-gather fused into the GEMM epilogue. */
wgb_status wgb_gemm_row_sharded_fused(wgb_pass *pass, wgb_gemm_variant variant, wgb_peer_gather *pg,
                                      const wgb_buffer *m1_local, const wgb_view_shape *m1_local_shape,
                                      const wgb_buffer *m2, const wgb_view_shape *m2_shape, wgb_dtype in_dtype,
                                      wgb_dtype out_dtype, wgb_f32_mode f32_mode);

/* flags: WGB_GATHER_NO_WAIT = return after this rank's GEMM + peer stores are queued, without waiting for the peers' panels;
 * pair with wgb_peer_gather_wait.  Flag waits between GPUs are bounded by WGB_PEER_TIMEOUT_MS (default 60000; 0 = unbounded):
 * a peer that never arrives surfaces as a CUDA error on this rank instead of a hung GPU. */
#define WGB_GATHER_NO_WAIT 1u
wgb_status wgb_gemm_row_sharded_fused_ex(wgb_pass *pass, wgb_gemm_variant variant, wgb_peer_gather *pg,
                                         const wgb_buffer *m1_local, const wgb_view_shape *m1_local_shape,
                                         const wgb_buffer *m2, const wgb_view_shape *m2_shape, wgb_dtype in_dtype,
                                         wgb_dtype out_dtype, wgb_f32_mode f32_mode, uint32_t flags);

/* The same collective with HOST operands, enqueued (the N > 1 counterpart of wgb_gemm_host_enqueue): this rank's dense
 * column-major A block ([M_local x K], or [K x M_local] for the transposed variants) and B ([K x N]) are uploaded into one of two
 * alternating device slots (when the context also has a communicator of the same ranks, wgb_comm_init_rank, each rank uploads only
 * its 1/P column slice of B and the slices are all-gathered over NVLink: B crosses the host links once per box; WGB_SHARD_B_UPLOAD=0
 * restores the whole-B upload), the fused GEMM + all-gather runs once they have landed, and the result is downloaded into out_host:
 * this rank's [M_local x N] panel (download_all == 0; the ranks of a box assemble C in host memory) or the whole gathered cube
 * [M_local x N x nranks].  Returns once everything is queued: the upload of the next product overlaps the GEMM and the download
 * of this one.  out_host is complete after wgb_ctx_sync(), or after wgb_gemm_host_flush() for later work on the queue.  Pinned
 * host memory (wgb_host_alloc) is needed for the overlap.  Collective: every rank must make the same sequence of calls. */
wgb_status wgb_gemm_row_sharded_fused_host_enqueue(wgb_ctx *ctx, wgb_gemm_variant variant, wgb_peer_gather *pg, uint32_t M_local,
                                                   uint32_t N, uint32_t K, void *out_host, const void *m1_local_host,
                                                   const void *m2_host, wgb_dtype in_dtype, wgb_dtype out_dtype,
                                                   wgb_f32_mode f32_mode, int download_all);

#ifdef __cplusplus
}
#endif
#endif /* WGB200_H */
