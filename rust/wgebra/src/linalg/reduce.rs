//! crates/wgebra/src/linalg/reduce.rs:13-124
use super::{require_f32, ComposerError};
use bytemuck::Pod;
use nalgebra::DVector;
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
    pub fn dispatch<'a, T: Pod>(&self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        value: impl Into<GpuVectorView<'a, T>>, result: &GpuScalar<T>) {
        require_f32::<T>("Reduce");
        let value = value.into();
        let _ = shapes.get(device, value.shape());
        let sv = value.shape().into();
        sys::check(unsafe { sys::wgb_reduce(pass.raw(), self.1 as i32, value.buffer().raw(), &sv, result.buffer().raw()) });
    }
    /// reduce.rs:116-124
    #[doc(hidden)]
    pub fn eval_cpu(&self, val: &DVector<f32>) -> f32 {
        match self.1 {
            ReduceOp::Min => val.min(),
            ReduceOp::Max => val.max(),
            ReduceOp::Prod => val.product(),
            ReduceOp::Sum => val.sum(),
            ReduceOp::SqNorm => val.norm_squared(),
        }
    }
}
