// wgebra_b200.hpp — header-only C++17 host mirror of the reference's Rust surface over the C ABI
// (include/wgb200.h).  Same names, argument order and error behaviour as
//   /root/reference/crates/wgcore/src/{gpu,tensor,shapes,kernel,timestamps}.rs and
//   /root/reference/crates/wgebra/src/linalg/{gemm,gemv,op_assign,reduce}.rs
// (the reference is Rust; no Rust toolchain exists in this image, so the compiled-language host side is C++).
//
//   wgb::GpuInstance gpu;                                   // GpuInstance::new().await
//   wgb::ViewShapeBuffers shapes;
//   auto m1 = wgb::TensorBuilder::matrix(r, c, wgb::STORAGE).build_init<float>(gpu.device(), data);
//   auto enc = gpu.device().create_command_encoder();
//   { auto pass = enc.compute_pass("test");                 // drop(pass) at scope end
//     wgb::Gemm::from_device(gpu.device()).dispatch(gpu.device(), shapes, pass, out, m1, m2); }
//   gpu.queue().submit(enc.finish());
//   std::vector<float> r = out.read(gpu.device());
//
// A dimension mismatch — a panic in the reference (assert_eq!, gemm.rs:91-95 etc.) — throws wgb::DimensionMismatch.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/wgb200.h"

namespace wgb {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};
struct DimensionMismatch : Error {
    using Error::Error;
};
inline void check(wgb_status s) {
    if (s == WGB_OK) return;
    const std::string msg = wgb_last_error_string();
    if (s == WGB_ERR_DIM_MISMATCH) throw DimensionMismatch(s, msg);
    throw Error(s, msg);
}

// wgpu::BufferUsages
enum : uint32_t { MAP_READ = 1u << 0, MAP_WRITE = 1u << 1, COPY_SRC = 1u << 2, COPY_DST = 1u << 3, UNIFORM = 1u << 6, STORAGE = 1u << 7 };

template <typename T> struct DType;
template <> struct DType<float> { static constexpr wgb_dtype value = WGB_F32; };
struct bf16 { uint16_t bits; };
template <> struct DType<bf16> { static constexpr wgb_dtype value = WGB_BF16; };
template <> struct DType<uint32_t> { static constexpr wgb_dtype value = WGB_F32; };   // 4-byte elements (GpuVector<u32> of the scan / sort primitives)

using ViewShape = wgb_view_shape;  // shapes.rs:9-21

// shapes.rs:46-116 — a cache of per-shape uniform buffers in the reference; a no-op here (shape = kernel parameter).
struct ViewShapeBuffers {
    template <typename Device> const ViewShape &get(const Device &, const ViewShape &s) const { return s; }
    void clear_tmp() const {}
};

class Device;
class ComputePass;

class GpuTimestamps {  // timestamps.rs:9-248
  public:
    GpuTimestamps(const Device &dev, uint32_t capacity);
    ~GpuTimestamps() { for (auto e : ev_) wgb_event_destroy(e); }
    std::pair<wgb_event *, wgb_event *> next_compute_pass_timestamp_writes() {
        if (len_ + 2 > ev_.size()) throw std::out_of_range("GpuTimestamps capacity exceeded");
        len_ += 2;
        return {ev_[len_ - 2], ev_[len_ - 1]};
    }
    void clear() { len_ = 0; }
    std::vector<float> wait_for_results_ms() const {
        std::vector<float> out(len_, 0.f);
        for (size_t i = 1; i < len_; ++i) check(wgb_event_elapsed_ms(ev_[0], ev_[i], &out[i]));
        return out;
    }
  private:
    std::vector<wgb_event *> ev_;
    size_t len_ = 0;
};

class ComputePass {  // kernel.rs:15-26; the destructor is Rust's drop(pass)
  public:
    ComputePass(wgb_ctx *ctx, const char *label, GpuTimestamps *ts) {
        wgb_event *b = nullptr, *e = nullptr;
        if (ts) std::tie(b, e) = ts->next_compute_pass_timestamp_writes();
        check(wgb_pass_begin(ctx, label, b, e, &h_));
    }
    ComputePass(ComputePass &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ComputePass(const ComputePass &) = delete;
    ~ComputePass() { if (h_) wgb_pass_end(h_); }
    wgb_pass *raw() const { return h_; }
    int last_gemm_path() const { int p = 0; check(wgb_pass_last_gemm_path(h_, &p)); return p; }
  private:
    wgb_pass *h_ = nullptr;
};

class CommandEncoder {
  public:
    explicit CommandEncoder(wgb_ctx *ctx) : ctx_(ctx) {}
    ComputePass compute_pass(const char *label, GpuTimestamps *timestamps = nullptr) { return ComputePass(ctx_, label, timestamps); }
    CommandEncoder &finish() { return *this; }
    wgb_ctx *ctx() const { return ctx_; }
  private:
    wgb_ctx *ctx_;
};

class Device {
  public:
    explicit Device(int ordinal) { check(wgb_ctx_create(ordinal, &h_)); }
    ~Device() { wgb_ctx_destroy(h_); }
    Device(const Device &) = delete;
    CommandEncoder create_command_encoder() const { return CommandEncoder(h_); }
    void poll_wait() const { check(wgb_ctx_sync(h_)); }     // device.poll(PollType::wait())
    uint64_t launch_count() const { uint64_t n = 0; check(wgb_ctx_launch_count(h_, &n)); return n; }
    wgb_ctx *raw() const { return h_; }
  private:
    wgb_ctx *h_ = nullptr;
};

inline GpuTimestamps::GpuTimestamps(const Device &dev, uint32_t capacity) {
    for (uint32_t i = 0; i < capacity; ++i) {
        wgb_event *e = nullptr;
        check(wgb_event_create(dev.raw(), &e));
        ev_.push_back(e);
    }
}

class Queue {
  public:
    explicit Queue(const Device &d) : d_(d) {}
    void submit(CommandEncoder &) const { check(wgb_submit(d_.raw())); }   // queue.submit(Some(encoder.finish()))
  private:
    const Device &d_;
};

class GpuInstance {  // gpu.rs:7-79
  public:
    explicit GpuInstance(int ordinal = 0) : device_(ordinal), queue_(device_) {}
    const Device &device() const { return device_; }
    const Queue &queue() const { return queue_; }
  private:
    Device device_;
    Queue queue_;
};

struct Buffer {  // wgpu::Buffer
    wgb_buffer *raw = nullptr;
    size_t bytes = 0;
    Buffer(wgb_buffer *r, size_t b) : raw(r), bytes(b) {}
    Buffer(const Buffer &) = delete;
    ~Buffer() { wgb_buffer_destroy(raw); }
    size_t size() const { return bytes; }
};

// tensor.rs:17-39 — MatrixOrdering markers.  The reference carries the ordering as a type parameter of the view; here it is
// a field of the view (set through as_view<RowMajor>() / reshape<D2, RowMajor>()), so the dispatch signatures stay as they are.
struct ColumnMajor { static constexpr wgb_ordering value = WGB_COLUMN_MAJOR; static constexpr bool is_row_major() { return false; } };
struct RowMajor { static constexpr wgb_ordering value = WGB_ROW_MAJOR; static constexpr bool is_row_major() { return true; } };

template <typename T, int DIM> class GpuTensorView {  // tensor.rs:416-511
  public:
    GpuTensorView(ViewShape s, const Buffer *b, wgb_ordering o = WGB_COLUMN_MAJOR) : view_shape_(s), buffer_(b), ordering_(o) {}
    ViewShape shape() const { return view_shape_; }
    wgb_ordering ordering() const { return ordering_; }
    const Buffer *buffer() const { return buffer_; }
    uint32_t len() const { return view_shape_.size[0]; }
    // GpuVectorView::rows (:445-462) / GpuMatrixView::rows (:498-510)
    GpuTensorView rows(uint32_t first_row, uint32_t nrows) const {
        ViewShape s = view_shape_;
        if (DIM == 1 && first_row + nrows > len()) throw std::out_of_range("Rows slice range out of bounds");
        s.size[0] = nrows;
        if (DIM == 1) { s.size[1] = 1; s.size[2] = 1; } else { s.size[2] = 1; }
        s.offset += (ordering_ == WGB_ROW_MAJOR && DIM > 1 ? s.stride : 1u) * first_row;   // shape.wgsl:49-53 for row-major views
        return GpuTensorView(s, buffer_, ordering_);
    }
    GpuTensorView columns(uint32_t first_col, uint32_t ncols) const {   // :484-496
        ViewShape s = view_shape_;
        s.size[1] = ncols; s.size[2] = 1;
        s.offset += (ordering_ == WGB_ROW_MAJOR ? 1u : s.stride) * first_col;
        return GpuTensorView(s, buffer_, ordering_);
    }
    GpuTensorView<T, 2> matrix(uint32_t matrix_id) const {            // :466-481
        if (matrix_id >= view_shape_.size[2]) throw std::out_of_range("matrix id out of range");
        ViewShape s = view_shape_;
        s.size[2] = 1;
        s.offset += s.stride_mat * matrix_id;
        s.stride_mat = 1;
        return GpuTensorView<T, 2>(s, buffer_, ordering_);
    }
    template <int D2> operator GpuTensorView<T, D2>() const { return GpuTensorView<T, D2>(view_shape_, buffer_, ordering_); }
  private:
    ViewShape view_shape_;
    const Buffer *buffer_;
    wgb_ordering ordering_;
};

template <typename T, int DIM> class GpuTensor {  // tensor.rs:192-399
  public:
    GpuTensor(std::array<uint32_t, DIM> shape, std::unique_ptr<Buffer> buf) : shape_(shape), buffer_(std::move(buf)) {}
    uint64_t len() const { uint64_t n = 1; for (auto s : shape_) n *= s; return n; }
    uint64_t bytes_len() const { return sizeof(T) * len(); }
    std::array<uint32_t, DIM> shape() const { return shape_; }
    const Buffer &buffer() const { return *buffer_; }

    // reshape (:514-541) with default strides, offset 0
    template <int D2, typename Ordering = ColumnMajor> GpuTensorView<T, D2> reshape(std::array<uint32_t, D2> shape) const {
        ViewShape s{{1, 1, 1}, 0, 0, 0};
        for (int i = 0; i < D2 && i < 3; ++i) s.size[i] = shape[i];
        const uint32_t s0 = D2 > 0 ? shape[0] : 1, s1 = D2 > 1 ? shape[1] : 1;
        s.stride = Ordering::is_row_major() ? s1 : s0;   // :525-529
        s.stride_mat = s0 * s1;
        return GpuTensorView<T, D2>(s, buffer_.get(), Ordering::value);
    }
    template <int D2 = 3, typename Ordering = ColumnMajor> GpuTensorView<T, D2> as_embedded_view() const {   // :287-297
        static_assert(D2 >= DIM, "Can only embed into a higher-order tensor view.");
        std::array<uint32_t, D2> e;
        e.fill(1);
        for (int i = 0; i < DIM; ++i) e[i] = shape_[i];
        return reshape<D2, Ordering>(e);
    }
    template <typename Ordering = ColumnMajor> GpuTensorView<T, DIM> as_view() const { return as_embedded_view<DIM, Ordering>(); }   // :282-284
    template <int D2> operator GpuTensorView<T, D2>() const { return as_embedded_view<(D2 > DIM ? D2 : DIM)>(); }   // :403-409

    GpuTensorView<T, 1> column(uint32_t i) const {                       // GpuMatrix::column :574-585
        return GpuTensorView<T, 1>(ViewShape{{shape_[0], 1, 1}, 1, 1, shape_[0] * i}, buffer_.get());
    }
    GpuTensorView<T, 2> columns(uint32_t first_col, uint32_t ncols) const {   // :600-612
        return GpuTensorView<T, 2>(ViewShape{{shape_[0], ncols, 1}, shape_[0], shape_[0] * shape_[1], first_col * shape_[0]}, buffer_.get());
    }
    GpuTensorView<T, (DIM == 1 ? 1 : 2)> rows(uint32_t first_row, uint32_t nrows) const {   // :614-626 / :669-681
        if (DIM == 1) return GpuTensorView<T, (DIM == 1 ? 1 : 2)>(ViewShape{{nrows, 1, 1}, shape_[0], shape_[0], first_row}, buffer_.get());
        return GpuTensorView<T, (DIM == 1 ? 1 : 2)>(ViewShape{{nrows, shape_[DIM > 1 ? 1 : 0], 1}, shape_[0], shape_[0] * shape_[DIM > 1 ? 1 : 0], first_row}, buffer_.get());
    }

    void copy_from(CommandEncoder &enc, const GpuTensor &src) const {    // :227-233
        if (len() != src.len()) throw Error(WGB_ERR_DIM_MISMATCH, "copy_from: length mismatch");
        check(wgb_buffer_copy(enc.ctx(), nullptr, buffer_->raw, 0, src.buffer_->raw, 0, bytes_len()));
    }
    std::vector<T> read(const Device &dev) const {                        // :375-384 (blocks like poll(wait))
        std::vector<T> out(len());
        check(wgb_buffer_read(dev.raw(), buffer_->raw, 0, out.data(), bytes_len()));
        return out;
    }
  private:
    std::array<uint32_t, DIM> shape_;
    std::unique_ptr<Buffer> buffer_;
};
template <typename T> using GpuScalar = GpuTensor<T, 0>;
template <typename T> using GpuVector = GpuTensor<T, 1>;
template <typename T> using GpuMatrix = GpuTensor<T, 2>;
template <typename T> using GpuCube = GpuTensor<T, 3>;
template <typename T> using GpuVectorView = GpuTensorView<T, 1>;
template <typename T> using GpuMatrixView = GpuTensorView<T, 2>;
template <typename T> using GpuCubeView = GpuTensorView<T, 3>;

template <int DIM> class TensorBuilder {  // tensor.rs:65-187
  public:
    TensorBuilder(std::array<uint32_t, DIM> shape, uint32_t usage) : shape_(shape), usage_(usage) {}
    uint64_t len() const { uint64_t n = 1; for (auto s : shape_) n *= s; return n; }
    template <typename T> GpuTensor<T, DIM> build(const Device &dev) const {                     // :112-129
        wgb_buffer *b = nullptr;
        check(wgb_buffer_create(dev.raw(), sizeof(T) * len(), usage_, &b));
        return GpuTensor<T, DIM>(shape_, std::make_unique<Buffer>(b, sizeof(T) * len()));
    }
    template <typename T> GpuTensor<T, DIM> build_init(const Device &dev, const std::vector<T> &data) const {   // :175-186
        if (data.size() < len()) throw Error(WGB_ERR_INVALID, "Incorrect number of elements provided for initializing Tensor.");
        wgb_buffer *b = nullptr;
        check(wgb_buffer_create_init(dev.raw(), data.data(), sizeof(T) * len(), usage_, &b));
        return GpuTensor<T, DIM>(shape_, std::make_unique<Buffer>(b, sizeof(T) * len()));
    }
  private:
    std::array<uint32_t, DIM> shape_;
    uint32_t usage_;
};
struct Tensors {   // TensorBuilder::{scalar, vector, matrix} (:72-90)
    static TensorBuilder<0> scalar(uint32_t usage) { return TensorBuilder<0>({}, usage); }
    static TensorBuilder<1> vector(uint32_t dim, uint32_t usage) { return TensorBuilder<1>({dim}, usage); }
    static TensorBuilder<2> matrix(uint32_t r, uint32_t c, uint32_t usage) { return TensorBuilder<2>({r, c}, usage); }
};

// ---- wgebra::linalg -------------------------------------------------------------------------
enum class GemmVariant { Gemm = 0, GemmFast = 1, GemmTr = 2, GemmTrFast = 3 };          // gemm.rs:25-35
enum class GemvVariant { Gemv = 0, GemvFast = 1, GemvTr = 2, GemvTrFast = 3 };          // gemv.rs:24-34
enum class OpAssignVariant { Add = 0, Sub = 1, Mul = 2, Div = 3, Copy = 4 };            // op_assign.rs:15-26
enum class ReduceOp { Min = 0, Max = 1, Sum = 2, Prod = 3, SqNorm = 4 };                // reduce.rs:16-27

class Gemm {  // gemm.rs:9-127
  public:
    static Gemm from_device(const Device &) { return Gemm(); }
    wgb_f32_mode f32_mode = WGB_F32_AUTO;
    template <typename T> void dispatch(const Device &d, const ViewShapeBuffers &s, ComputePass &p, GpuCubeView<T> out, GpuCubeView<T> m1, GpuCubeView<T> m2) const {
        dispatch_generic<T>(d, s, p, out, m1, m2, GemmVariant::Gemm);
    }
    template <typename T> void dispatch_tr(const Device &d, const ViewShapeBuffers &s, ComputePass &p, GpuCubeView<T> out, GpuCubeView<T> m1, GpuCubeView<T> m2) const {
        dispatch_generic<T>(d, s, p, out, m1, m2, GemmVariant::GemmTr);
    }
    template <typename T> void dispatch_generic(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuCubeView<T> out, GpuCubeView<T> m1,
                                                GpuCubeView<T> m2, GemmVariant variant) const {
        const ViewShape so = shapes.get(d, out.shape()), s1 = shapes.get(d, m1.shape()), s2 = shapes.get(d, m2.shape());   // gemm.rs:98-100
        if (out.ordering() != WGB_COLUMN_MAJOR || m1.ordering() != WGB_COLUMN_MAJOR || m2.ordering() != WGB_COLUMN_MAJOR) {
            check(wgb_gemm_ord(pass.raw(), (wgb_gemm_variant)variant, out.buffer()->raw, &so, out.ordering(), m1.buffer()->raw, &s1, m1.ordering(),
                               m2.buffer()->raw, &s2, m2.ordering(), DType<T>::value, DType<T>::value, f32_mode, -1, nullptr, nullptr));
            return;
        }
        check(wgb_gemm_ex(pass.raw(), (wgb_gemm_variant)variant, out.buffer()->raw, &so, m1.buffer()->raw, &s1, m2.buffer()->raw, &s2,
                          DType<T>::value, DType<T>::value, f32_mode));
    }

    // Extension (wgb_gemm_op): out = (m1 * m2) (op) operand — Gemm::dispatch + OpAssign::dispatch in one launch.
    template <typename T> void dispatch_op(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuCubeView<T> out, GpuCubeView<T> m1,
                                           GpuCubeView<T> m2, int op /* OpAssignVariant */, GpuCubeView<T> operand,
                                           GemmVariant variant = GemmVariant::Gemm) const {
        const ViewShape so = shapes.get(d, out.shape()), s1 = shapes.get(d, m1.shape()), s2 = shapes.get(d, m2.shape()), se = shapes.get(d, operand.shape());
        check(wgb_gemm_op(pass.raw(), (wgb_gemm_variant)variant, out.buffer()->raw, &so, m1.buffer()->raw, &s1, m2.buffer()->raw, &s2,
                          DType<T>::value, DType<T>::value, f32_mode, (wgb_op_assign_variant)op, operand.buffer()->raw, &se));
    }

    // Extension (wgb_gemm_reduce): result[j] = reduce_op over column j of m1 * m2 (axis 1) or result[i] over row i (axis 2); the
    // product is never stored — Gemm::dispatch + one Reduce::dispatch per GpuMatrix::column(j) in one pass over the operands.
    template <typename T> void dispatch_reduce(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuVectorView<float> result,
                                               GpuCubeView<T> m1, GpuCubeView<T> m2, int reduce_op /* ReduceOp */, int axis = 1,
                                               GemmVariant variant = GemmVariant::Gemm) const {
        const ViewShape rs = shapes.get(d, result.shape()), s1 = shapes.get(d, m1.shape()), s2 = shapes.get(d, m2.shape());
        check(wgb_gemm_reduce(pass.raw(), (wgb_gemm_variant)variant, axis, (wgb_reduce_op)reduce_op, result.buffer()->raw, &rs, m1.buffer()->raw,
                              &s1, m2.buffer()->raw, &s2, DType<T>::value, f32_mode));
    }
};

// Extension (wgb_graph_*): record a dispatch chain once, replay it with one launch.
class Graph {
  public:
    static void capture_begin(const Device &d) { check(wgb_graph_capture_begin(d.raw())); }
    static Graph capture_end(const Device &d) { Graph g; check(wgb_graph_capture_end(d.raw(), &g.h_)); return g; }
    Graph(Graph &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Graph(const Graph &) = delete;
    ~Graph() { if (h_) wgb_graph_destroy(h_); }
    void launch() const { check(wgb_graph_launch(h_)); }
  private:
    Graph() = default;
    wgb_graph *h_ = nullptr;
};

class Gemv {  // gemv.rs:9-137
  public:
    static Gemv from_device(const Device &) { return Gemv(); }
    template <typename T> void dispatch(const Device &d, const ViewShapeBuffers &s, ComputePass &p, GpuCubeView<T> out, GpuCubeView<T> m, GpuCubeView<T> v) const {
        dispatch_generic<T>(d, s, p, out, m, v, GemvVariant::Gemv);
    }
    template <typename T> void dispatch_tr(const Device &d, const ViewShapeBuffers &s, ComputePass &p, GpuCubeView<T> out, GpuCubeView<T> m, GpuCubeView<T> v) const {
        dispatch_generic<T>(d, s, p, out, m, v, GemvVariant::GemvTr);
    }
    template <typename T> void dispatch_generic(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuCubeView<T> out, GpuCubeView<T> m,
                                                GpuCubeView<T> v, GemvVariant variant) const {
        const ViewShape so = shapes.get(d, out.shape()), sm = shapes.get(d, m.shape()), sv = shapes.get(d, v.shape());
        check(wgb_gemv_ord(pass.raw(), (wgb_gemv_variant)variant, out.buffer()->raw, &so, m.buffer()->raw, &sm, m.ordering(), v.buffer()->raw, &sv));
    }
    // out = (m * v) (op) operand: Gemv::dispatch + OpAssign::dispatch(out, operand) as one launch (wgb_gemv_op); operand may be out
    template <typename T> void dispatch_op(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuCubeView<T> out, GpuCubeView<T> m,
                                           GpuCubeView<T> v, OpAssignVariant op, GpuCubeView<T> operand, GemvVariant variant = GemvVariant::Gemv) const {
        const ViewShape so = shapes.get(d, out.shape()), sm = shapes.get(d, m.shape()), sv = shapes.get(d, v.shape()), se = shapes.get(d, operand.shape());
        check(wgb_gemv_op(pass.raw(), (wgb_gemv_variant)variant, out.buffer()->raw, &so, m.buffer()->raw, &sm, m.ordering(), v.buffer()->raw, &sv,
                          (int)op, operand.buffer()->raw, &se));
    }
    // result = reduce_op(m * v): Gemv::dispatch + Reduce::dispatch(out, result) as one launch (wgb_gemv_reduce), bit-identical to the chain
    template <typename T> void dispatch_reduce(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, const GpuScalar<T> &result,
                                               GpuCubeView<T> m, GpuCubeView<T> v, int reduce_op /* ReduceOp */,
                                               GemvVariant variant = GemvVariant::Gemv) const {
        const ViewShape sm = shapes.get(d, m.shape()), sv = shapes.get(d, v.shape());
        check(wgb_gemv_reduce(pass.raw(), (wgb_gemv_variant)variant, (wgb_reduce_op)reduce_op, result.buffer()->raw, m.buffer()->raw, &sm,
                              m.ordering(), v.buffer()->raw, &sv));
    }
};

class OpAssign {  // op_assign.rs:43-94
  public:
    OpAssignVariant variant;
    static OpAssign make(const Device &, OpAssignVariant op) { return OpAssign{op}; }   // OpAssign::new
    template <typename T> void dispatch(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuVectorView<T> in_out_a, GpuVectorView<T> in_b) const {
        const ViewShape sa = shapes.get(d, in_out_a.shape()), sb = shapes.get(d, in_b.shape());
        check(wgb_op_assign(pass.raw(), (wgb_op_assign_variant)variant, in_out_a.buffer()->raw, &sa, in_b.buffer()->raw, &sb));
    }
};

class Reduce {  // reduce.rs:62-124
  public:
    ReduceOp op;
    static Reduce make(const Device &, ReduceOp op) { return Reduce{op}; }              // Reduce::new
    template <typename T> void dispatch(const Device &d, const ViewShapeBuffers &shapes, ComputePass &pass, GpuVectorView<T> value, const GpuScalar<T> &result) const {
        const ViewShape sv = shapes.get(d, value.shape());
        check(wgb_reduce(pass.raw(), (wgb_reduce_op)op, value.buffer()->raw, &sv, result.buffer().raw));
    }
};

// wgrapier/src/dynamics/prefix_sum.rs:119-224 — auxiliary levels in the reference; only its bookkeeping here.
struct PrefixSumWorkspace {
    std::vector<uint32_t> stages;
    static PrefixSumWorkspace make() { return {}; }
    static PrefixSumWorkspace with_capacity(const Device &d, uint32_t buffer_len) { PrefixSumWorkspace w; w.reserve(d, buffer_len); return w; }
    void reserve(const Device &, uint32_t buffer_len) {   // :185-224 (the reference never terminates for 0)
        stages.clear();
        uint32_t stage_len = (buffer_len + 255) / 256;
        while (stage_len > 1) { stages.push_back(stage_len); stage_len = (stage_len + 255) / 256; }
        stages.push_back(1);
    }
};

class WgPrefixSum {  // prefix_sum.rs:22-117: in-place exclusive prefix sum of a GpuVector<u32>
  public:
    static WgPrefixSum from_device(const Device &) { return WgPrefixSum(); }
    void dispatch(const Device &d, ComputePass &pass, PrefixSumWorkspace &workspace, const GpuVector<uint32_t> &data) const {
        dispatch(d, pass, workspace, data.as_view());
    }
    void dispatch(const Device &d, ComputePass &pass, PrefixSumWorkspace &workspace, GpuVectorView<uint32_t> data) const {
        workspace.reserve(d, data.len());
        const ViewShape s = data.shape();
        check(wgb_prefix_sum(pass.raw(), data.buffer()->raw, &s));
    }
    static void eval_cpu(std::vector<uint32_t> &v) {   // :101-117
        uint32_t run = 0;
        for (auto &x : v) { const uint32_t t = x; x = run; run += t; }
    }
};

struct RadixSortWorkspace {  // wgparry/src/utils/radix_sort/mod.rs:82-109 (buffers live in the context here)
    static RadixSortWorkspace make(const Device &) { return {}; }
};

class RadixSort {  // radix_sort/mod.rs:67-223
  public:
    static RadixSort from_device(const Device &) { return RadixSort(); }
    void dispatch(const Device &, ComputePass &pass, RadixSortWorkspace &, const GpuVector<uint32_t> &input_keys,
                  const GpuVector<uint32_t> &input_values, const GpuScalar<uint32_t> &n_sort, uint32_t sorting_bits,
                  const GpuVector<uint32_t> &output_keys, const GpuVector<uint32_t> &output_values) const {
        if (input_keys.len() != input_values.len())     // assert_eq! :121-125
            throw DimensionMismatch(WGB_ERR_DIM_MISMATCH, "Input keys and values must have the same number of elements");
        if (sorting_bits > 32) throw Error(WGB_ERR_INVALID, "Can only sort up to 32 bits");   // assert! :126
        const ViewShape sk = input_keys.as_view().shape(), sv = input_values.as_view().shape(), so = output_keys.as_view().shape(),
                        sw = output_values.as_view().shape();
        check(wgb_radix_sort(pass.raw(), input_keys.buffer().raw, &sk, input_values.buffer().raw, &sv, n_sort.buffer().raw, sorting_bits,
                             output_keys.buffer().raw, &so, output_values.buffer().raw, &sw));
    }
};

// ---- wgebra::geometry (factorization libraries) -----------------------------
// Element types in WGSL storage layout: what the reference's tests upload / read back (Matrix2 / Matrix4x3 / Matrix4,
// cholesky.rs:151-153; GpuLU* lu.rs:25-56; GpuQR* qr3.rs:15-20; GpuSymmetricEigen* eig3.rs:16-21; GpuSvd2 svd2.rs:12-19;
// GpuSvd3 svd3.rs:15-22).  m[c][r] = column c, row r; 3x3 matrices have 4-float columns.
template <int D> struct GpuMat { float m[D][D == 2 ? 2 : 4]; };
template <int D> struct GpuLU;
template <> struct GpuLU<2> { GpuMat<2> lu; uint32_t ia[2], ib[2], len, _pad; };
template <> struct GpuLU<3> { GpuMat<3> lu; uint32_t ia[4], ib[3], len; };
template <> struct GpuLU<4> { GpuMat<4> lu; uint32_t ia[4], ib[4], len, _pad[3]; };
template <int D> struct GpuQR { GpuMat<D> q, r; };
template <int D> struct GpuSymmetricEigen { GpuMat<D> eigenvectors; float eigenvalues[D == 2 ? 2 : 4]; };
template <int D> struct GpuSvd { GpuMat<D> u; float s[D == 2 ? 2 : 4]; GpuMat<D> vt; };
static_assert(sizeof(GpuMat<2>) == 16 && sizeof(GpuMat<3>) == 48 && sizeof(GpuMat<4>) == 64, "WGSL matrix layout");
static_assert(sizeof(GpuLU<2>) == 40 && sizeof(GpuLU<3>) == 80 && sizeof(GpuLU<4>) == 112, "lu.wgsl:12-34 layout");
static_assert(sizeof(GpuQR<2>) == 32 && sizeof(GpuQR<3>) == 96 && sizeof(GpuQR<4>) == 128, "qr2.wgsl:7-12 layout");
static_assert(sizeof(GpuSymmetricEigen<2>) == 24 && sizeof(GpuSymmetricEigen<3>) == 64 && sizeof(GpuSymmetricEigen<4>) == 80, "eig2.wgsl:7-12");
static_assert(sizeof(GpuSvd<2>) == 40 && sizeof(GpuSvd<3>) == 112, "svd2.wgsl:5-9 / svd3.wgsl:12-16 layout");

// One class per reference `Shader` struct.  The reference builds exactly one kind of kernel from each — out[i] = f(in[i])
// (cholesky.rs:53-63 ...) — and that is `dispatch`: KernelDispatch::new(device, &mut pass, &pipeline).bind0([in, out]).dispatch(len).
template <wgb_geom_op OP, int D, typename Out> class GeometryShader {
  public:
    static GeometryShader from_device(const Device &) { return GeometryShader(); }
    void dispatch(const Device &, ComputePass &pass, const GpuVector<GpuMat<D>> &inputs, const GpuVector<Out> &outputs) const {
        dispatch(pass, inputs.as_view(), outputs.as_view());
    }
    void dispatch(ComputePass &pass, GpuVectorView<GpuMat<D>> inputs, GpuVectorView<Out> outputs) const {
        static_assert(OP != WGB_GEOM_SVD || D < 4, "the reference has no 4x4 SVD");
        check(wgb_geometry_batch(pass.raw(), OP, D, inputs.buffer()->raw, inputs.shape().offset, outputs.buffer()->raw,
                                 outputs.shape().offset, inputs.len()));
    }
};
template <int D> using WgCholesky = GeometryShader<WGB_GEOM_CHOLESKY, D, GpuMat<D>>;        // cholesky.rs:21-38 WgCholesky2/3/4
template <int D> using WgLU = GeometryShader<WGB_GEOM_LU, D, GpuLU<D>>;                     // lu.rs:65-79
template <int D> using WgQR = GeometryShader<WGB_GEOM_QR, D, GpuQR<D>>;                     // qr2.rs:22-27 ...
template <int D> using WgSymmetricEigen = GeometryShader<WGB_GEOM_SYMMETRIC_EIGEN, D, GpuSymmetricEigen<D>>;   // eig2.rs:23-27 ...
template <int D> using WgSvd = GeometryShader<WGB_GEOM_SVD, D, GpuSvd<D>>;                  // svd2.rs:21-23, svd3.rs:24-28
template <int D> using WgInv = GeometryShader<WGB_GEOM_INV, D, GpuMat<D>>;                  // inv.rs:3-8 (inv2 / inv3 / inv4)

}  // namespace wgb
