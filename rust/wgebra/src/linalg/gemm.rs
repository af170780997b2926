//! crates/wgebra/src/linalg/gemm.rs:9-127 — same struct, enum and `dispatch*` signatures; one FFI call underneath.
use super::{dtype_of, ComposerError};
use bytemuck::Pod;
use wgcore::shapes::ViewShapeBuffers;
use wgcore::tensor::{GpuCubeView, MatrixOrdering};
use wgpu::{sys, ComputePass, ComputePipeline, Device};

/// Shader for computing the product of two matrices.
pub struct Gemm {
    pub gemm: ComputePipeline,
    pub gemm_fast: ComputePipeline,
    pub gemm_tr: ComputePipeline,
    pub gemm_tr_fast: ComputePipeline,
}

// `#[derive(Shader)] #[shader(derive(Shape), src = "gemm.wgsl", composable = false)]` of gemm.rs:9-11, written out
wgcore::impl_shader!(Gemm, "wgebra/src/linalg/gemm.wgsl", "wgmath_b200/csrc/gemm_tc_kernel.cuh, gemm_simt.cu", |_device| Gemm {
    gemm: ComputePipeline("gemm"),
    gemm_fast: ComputePipeline("gemm_fast"),
    gemm_tr: ComputePipeline("gemm_tr"),
    gemm_tr_fast: ComputePipeline("gemm_tr_fast"),
});

#[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
pub enum GemmVariant { Gemm, GemmFast, GemmTr, GemmTrFast }   // discriminants == wgb_gemm_variant

impl Gemm {
    /// Inherent twin of `Shader::from_device`, so the call compiles with or without the trait in scope.
    pub fn from_device(device: &Device) -> Result<Self, ComposerError> {
        <Self as wgcore::Shader>::from_device(device)
    }

    /// `out = m1 * m2`
    pub fn dispatch<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, T>>, m1: impl Into<GpuCubeView<'b, T>>, m2: impl Into<GpuCubeView<'b, T>>) {
        self.dispatch_generic(device, shapes, pass, out, m1, m2, GemmVariant::Gemm)
    }
    /// `out = tr(m1) * m2`
    pub fn dispatch_tr<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, T>>, m1: impl Into<GpuCubeView<'b, T>>, m2: impl Into<GpuCubeView<'b, T>>) {
        self.dispatch_generic(device, shapes, pass, out, m1, m2, GemmVariant::GemmTr)
    }
    pub fn dispatch_generic<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, T>>, m1: impl Into<GpuCubeView<'b, T>>, m2: impl Into<GpuCubeView<'b, T>>, variant: GemmVariant) {
        let (out, m1, m2) = (out.into(), m1.into(), m2.into());
        // the dimension asserts of gemm.rs:81-96 live in the library: status 2 -> panic!("Gemm: dimension mismatch. …");
        // the uniform buffers of gemm.rs:98-100 are still fetched so the cache behaves as in the reference
        let _ = (shapes.get(device, out.shape()), shapes.get(device, m1.shape()), shapes.get(device, m2.shape()));
        let (so, s1, s2) = (out.shape().into(), m1.shape().into(), m2.shape().into());
        let dt = dtype_of::<T>();
        sys::check(unsafe { sys::wgb_gemm_ex(pass.raw(), variant as i32, out.buffer().raw(), &so, m1.buffer().raw(), &s1,
                                             m2.buffer().raw(), &s2, dt, dt, /* WGB_F32_AUTO: parity-gated 3xTF32 */ 0) });
    }

    /// Extension: the same product on views of any `MatrixOrdering` (tensor.rs:17-39; addressing of shape.wgsl:49-57).  The
    /// reference's `dispatch*` only accept `ColumnMajor` views (gemm.rs:65-74) and no shader is built with
    /// `row_major_shader_defs()`; the library computes every ordering combination in place (`wgb_gemm_ord`).
    pub fn dispatch_ordered<'a, 'b, T: Pod, O: MatrixOrdering + 'b, A: MatrixOrdering + 'b, B: MatrixOrdering + 'b>(
        &'a self, _device: &Device, _shapes: &ViewShapeBuffers, pass: &mut ComputePass, out: impl Into<GpuCubeView<'b, T, O>>,
        m1: impl Into<GpuCubeView<'b, T, A>>, m2: impl Into<GpuCubeView<'b, T, B>>, variant: GemmVariant) {
        let (out, m1, m2) = (out.into(), m1.into(), m2.into());
        let (so, s1, s2) = (out.shape().into(), m1.shape().into(), m2.shape().into());
        let dt = dtype_of::<T>();
        sys::check(unsafe { sys::wgb_gemm_ord(pass.raw(), variant as i32, out.buffer().raw(), &so, O::is_row_major() as i32,
                                              m1.buffer().raw(), &s1, A::is_row_major() as i32, m2.buffer().raw(), &s2,
                                              B::is_row_major() as i32, dt, dt, 0, -1, std::ptr::null(), std::ptr::null()) });
    }
}
