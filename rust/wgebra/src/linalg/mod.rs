//! Fundamental linear-algebra matrix/vector operations (linalg/mod.rs:1-13).
mod gemm;
mod gemv;
mod op_assign;
mod reduce;

pub use gemm::{Gemm, GemmVariant};
pub use gemv::{Gemv, GemvVariant};
pub use op_assign::{OpAssign, OpAssignVariant};
pub use reduce::{Reduce, ReduceOp};

/// Construction can only fail for "no usable sm_100 device"; the reference's `ComposerError` (shader compile errors,
/// gemm.rs derive / op_assign.rs:52) has no analogue, so the error type is a string newtype under the same name.
#[derive(Debug)]
pub struct ComposerError(pub String);

/// Element types the kernels accept: f32 (the reference's only type) and bf16 bit patterns for GEMM operands.
pub trait B200Scalar: wgcore::Pod { const DTYPE: i32; }
impl B200Scalar for f32 { const DTYPE: i32 = 0; }
#[derive(Copy, Clone, Default, Debug, PartialEq, Eq)]
#[repr(transparent)]
pub struct Bf16(pub u16);
unsafe impl bytemuck::Zeroable for Bf16 {}
unsafe impl bytemuck::Pod for Bf16 {}
impl B200Scalar for Bf16 { const DTYPE: i32 = 1; }
