//! Fundamental linear-algebra matrix/vector operations (linalg/mod.rs:1-13).
mod gemm;
mod gemv;
mod op_assign;
mod reduce;

pub use gemm::{Gemm, GemmVariant};
pub use gemv::{Gemv, GemvVariant};
pub use op_assign::{OpAssign, OpAssignVariant};
pub use reduce::{Reduce, ReduceOp};

/// The error type of every constructor, under the reference's own path (gemm.rs derive, op_assign.rs:52, reduce.rs:71).
pub use naga_oil::compose::ComposerError;

/// The reference's `dispatch*` are generic over `T: Pod` (gemm.rs:39, gemv.rs:38, op_assign.rs:71, reduce.rs:100) although its
/// shaders only exist for f32.  The bound stays; the element type the kernels are told about is resolved from the size of `T`:
/// 4 bytes = f32 (the reference's only type), 2 bytes = bf16 bit patterns (GEMM operands / output only, an extension).
pub(crate) fn dtype_of<T: wgcore::Pod>() -> i32 {
    match core::mem::size_of::<T>() {
        4 => 0, // WGB_F32
        2 => 1, // WGB_BF16
        n => panic!("wgebra (B200): unsupported element size {n} bytes (f32, or bf16 bit patterns for Gemm)"),
    }
}
pub(crate) fn require_f32<T: wgcore::Pod>(what: &str) {
    assert_eq!(core::mem::size_of::<T>(), 4, "{what}: the kernels behind this shader exist for f32 elements only");
}

/// bf16 bit pattern (an extension: the reference has no 16-bit element type); `Pod`, 2 bytes, so `Gemm::dispatch::<Bf16>` runs the
/// tcgen05 `kind::f16` kernels.
#[derive(Copy, Clone, Default, Debug, PartialEq, Eq)]
#[repr(transparent)]
pub struct Bf16(pub u16);
unsafe impl bytemuck::Zeroable for Bf16 {}
unsafe impl bytemuck::Pod for Bf16 {}
