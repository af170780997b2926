//! crates/wgebra/src/linalg/gemv.rs:9-137
use super::ComposerError;
use wgcore::shapes::ViewShapeBuffers;
use wgcore::tensor::GpuCubeView;
use wgpu::{sys, ComputePass, ComputePipeline, Device};

pub struct Gemv {
    pub gemv: ComputePipeline,
    pub gemv_fast: ComputePipeline,
    pub gemv_tr: ComputePipeline,
    pub gemv_tr_fast: ComputePipeline,
}

#[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
pub enum GemvVariant { Gemv, GemvFast, GemvTr, GemvTrFast }

impl Gemv {
    pub fn from_device(_device: &Device) -> Result<Self, ComposerError> {
        Ok(Self { gemv: ComputePipeline("gemv"), gemv_fast: ComputePipeline("gemv_fast"), gemv_tr: ComputePipeline("gemv_tr"), gemv_tr_fast: ComputePipeline("gemv_tr_fast") })
    }
    pub fn dispatch<'a, 'b>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, f32>>, m: impl Into<GpuCubeView<'b, f32>>, v: impl Into<GpuCubeView<'b, f32>>) {
        self.dispatch_generic(device, shapes, pass, out, m, v, GemvVariant::Gemv)
    }
    pub fn dispatch_tr<'a, 'b>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, f32>>, m: impl Into<GpuCubeView<'b, f32>>, v: impl Into<GpuCubeView<'b, f32>>) {
        self.dispatch_generic(device, shapes, pass, out, m, v, GemvVariant::GemvTr)
    }
    /// The `GemvTrFast -> GemvTr` fallback (gemv.rs:99-104) and the `out_nrows % 4` assert of the fast variants (:122) are
    /// applied by the library, which runs one kernel pair valid for every shape.
    pub fn dispatch_generic<'a, 'b>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        out: impl Into<GpuCubeView<'b, f32>>, m: impl Into<GpuCubeView<'b, f32>>, v: impl Into<GpuCubeView<'b, f32>>, variant: GemvVariant) {
        let (out, m, v) = (out.into(), m.into(), v.into());
        let (so, sm, sv) = (shapes.get(device, out.shape()), shapes.get(device, m.shape()), shapes.get(device, v.shape()));
        sys::check(unsafe { sys::wgb_gemv(pass.raw(), variant as i32, out.buffer().raw(), &so, m.buffer().raw(), &sm, v.buffer().raw(), &sv) });
    }
}
