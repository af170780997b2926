//! crates/wgebra/src/linalg/op_assign.rs:12-94
use super::{require_f32, ComposerError};
use bytemuck::Pod;
use wgcore::shapes::ViewShapeBuffers;
use wgcore::tensor::GpuVectorView;
use wgpu::{sys, ComputePass, ComputePipeline, Device};

#[derive(Copy, Clone, PartialEq, Eq, Debug)]
#[non_exhaustive]
pub enum OpAssignVariant { Add, Sub, Mul, Div, Copy }

/// A GPU kernel for performing the operation described by [`OpAssignVariant`].
pub struct OpAssign(pub ComputePipeline, pub OpAssignVariant);

impl OpAssign {
    pub const SRC: &'static str = "(precompiled CUDA: wgmath_b200/csrc/level1.cu)";
    pub const FILE_PATH: &'static str = "wgebra/src/op_assign.wgsl";
    pub fn new(_device: &Device, op: OpAssignVariant) -> Result<Self, ComposerError> { Ok(OpAssign(ComputePipeline("op_assign"), op)) }
    /// `in_out_a ?= in_b`; panics with "Op-assign: dimension mismatch." like op_assign.rs:82-86.
    pub fn dispatch<'a, 'b, T: Pod>(&'a self, device: &Device, shapes: &ViewShapeBuffers, pass: &mut ComputePass,
        in_out_a: impl Into<GpuVectorView<'b, T>>, in_b: impl Into<GpuVectorView<'b, T>>) {
        require_f32::<T>("OpAssign");
        let (a, b) = (in_out_a.into(), in_b.into());
        let _ = (shapes.get(device, a.shape()), shapes.get(device, b.shape()));
        let (sa, sb) = (a.shape().into(), b.shape().into());
        sys::check(unsafe { sys::wgb_op_assign(pass.raw(), self.1 as i32, a.buffer().raw(), &sa, b.buffer().raw(), &sb) });
    }
}
