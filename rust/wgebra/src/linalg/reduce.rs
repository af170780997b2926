//! crates/wgebra/src/linalg/reduce.rs:13-124
use super::ComposerError;
use wgcore::shapes::ViewShapeBuffers;
use wgcore::tensor::{GpuScalar, GpuVectorView};
use wgpu::{sys, ComputePass, ComputePipeline, Device};

#[derive(Copy, Clone, PartialEq, Eq, Debug)]
#[non_exhaustive]
pub enum ReduceOp { Min, Max, Sum, Prod, SqNorm }

/// A GPU kernel for performing the operation described by [`ReduceOp`].
pub struct Reduce(pub ComputePipeline, pub ReduceOp);

impl Reduce {
    pub const SRC: &'static str = "(precompiled CUDA: wgmath_b200/csrc/level1.cu)";
    pub const FILE_PATH: &'static str = "wgebra/src/reduce.wgsl";
    pub fn new(_device: &Device, op: ReduceOp) -> Result<Self, ComposerError> { Ok(Self(ComputePipeline("reduce"), op)) }
    /// `result = reduce(value)`; the whole machine instead of the reference's single 128-thread workgroup (reduce.rs:112).
    pub fn dispatch<'a>(&self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        value: impl Into<GpuVectorView<'a, f32>>, result: &GpuScalar<f32>) {
        let value = value.into();
        let sv = shapes.get(device, value.shape());
        sys::check(unsafe { sys::wgb_reduce(pass.raw(), self.1 as i32, value.buffer().raw(), &sv, result.buffer().raw()) });
    }
    /// reduce.rs:116-124 on a plain slice (the reference takes a nalgebra `DVector`).
    #[doc(hidden)]
    pub fn eval_cpu(&self, val: &[f32]) -> f32 {
        match self.1 {
            ReduceOp::Min => val.iter().copied().fold(f32::INFINITY, f32::min),
            ReduceOp::Max => val.iter().copied().fold(f32::NEG_INFINITY, f32::max),
            ReduceOp::Prod => val.iter().product(),
            ReduceOp::Sum => val.iter().sum(),
            ReduceOp::SqNorm => val.iter().map(|x| x * x).sum(),
        }
    }
}
