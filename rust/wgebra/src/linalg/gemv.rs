//! crates/wgebra/src/linalg/gemv.rs:9-137
use super::{require_f32, ComposerError};
use bytemuck::Pod;
use wgcore::shapes::ViewShapeBuffers;
use wgcore::tensor::GpuCubeView;
use wgpu::{sys, ComputePass, ComputePipeline, Device};

pub struct Gemv {
    pub gemv: ComputePipeline,
    pub gemv_fast: ComputePipeline,
    pub gemv_tr: ComputePipeline,
    pub gemv_tr_fast: ComputePipeline,
}

// `#[derive(Shader)] #[shader(derive(Shape), src = "gemv.wgsl", composable = false)]` of gemv.rs:9-11, written out
wgcore::impl_shader!(Gemv, "wgebra/src/linalg/gemv.wgsl", "wgmath_b200/csrc/gemv.cu", |_device| Gemv {
    gemv: ComputePipeline("gemv"),
    gemv_fast: ComputePipeline("gemv_fast"),
    gemv_tr: ComputePipeline("gemv_tr"),
    gemv_tr_fast: ComputePipeline("gemv_tr_fast"),
});

#[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
pub enum GemvVariant { Gemv, GemvFast, GemvTr, GemvTrFast }

impl Gemv {
    /// Inherent twin of `Shader::from_device`, so the call compiles with or without the trait in scope.
    pub fn from_device(device: &Device) -> Result<Self, ComposerError> {
        <Self as wgcore::Shader>::from_device(device)
    }
    pub fn dispatch<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, T>>, m: impl Into<GpuCubeView<'b, T>>, v: impl Into<GpuCubeView<'b, T>>) {
        self.dispatch_generic(device, shapes, pass, out, m, v, GemvVariant::Gemv)
    }
    pub fn dispatch_tr<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, T>>, m: impl Into<GpuCubeView<'b, T>>, v: impl Into<GpuCubeView<'b, T>>) {
        self.dispatch_generic(device, shapes, pass, out, m, v, GemvVariant::GemvTr)
    }
    /// The `GemvTrFast -> GemvTr` fallback (gemv.rs:99-104) and the `out_nrows % 4` assert of the fast variants (:122) are
    /// applied by the library, which runs one kernel pair valid for every shape.
    pub fn dispatch_generic<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, T>>, m: impl Into<GpuCubeView<'b, T>>, v: impl Into<GpuCubeView<'b, T>>, variant: GemvVariant) {
        require_f32::<T>("Gemv");
        let (out, m, v) = (out.into(), m.into(), v.into());
        let _ = (shapes.get(device, out.shape()), shapes.get(device, m.shape()), shapes.get(device, v.shape()));
        let (so, sm, sv) = (out.shape().into(), m.shape().into(), v.shape().into());
        sys::check(unsafe { sys::wgb_gemv(pass.raw(), variant as i32, out.buffer().raw(), &so, m.buffer().raw(), &sm, v.buffer().raw(), &sv) });
    }
}
