//! crates/wgebra/src/geometry/{cholesky,lu,qr2,qr3,qr4,eig2,eig3,eig4,svd2,svd3,inv}.rs over `wgb_geometry_batch`.
//!
//! In the reference each struct is a `Shader` whose WGSL functions other shaders import; the only kernels built from them are
//! the per-module test kernels `out[i] = f(in[i])` (cholesky.rs:53-63, lu.rs:101-111, qr2.rs:36-46, eig3.rs:36-46,
//! svd3.rs:34-44).  Here the struct names stay and gain `dispatch`, which is that kernel; CUDA kernels that want the functions
//! themselves include `wgmath_b200/csrc/geometry.cuh`.  Element types are the reference's (same bytes as `GpuVector::init` /
//! `encase` upload today).  NOT COMPILED here (../../README.md).
use nalgebra::{Matrix2, Matrix4, Matrix4x3, SVector, Vector2, Vector4};
use wgcore::tensor::GpuVector;
use wgpu::{sys, ComputePass, Device};

use crate::linalg::ComposerError;

const CHOLESKY: i32 = 0;
const LU: i32 = 1;
const QR: i32 = 2;
const SYMMETRIC_EIGEN: i32 = 3;
const SVD: i32 = 4;
const INV: i32 = 5;

/// lu.rs:25-56 `gpu_output_types!` — flattened to the WGSL storage layout (a `vec3<u32>` is 12 bytes in a 16-byte slot and
/// `len` packs right behind `ib`).
#[repr(C)] #[derive(Copy, Clone, PartialEq)] pub struct GpuLU2 { pub lu: Matrix2<f32>, pub ia: SVector<u32, 2>, pub ib: SVector<u32, 2>, pub len: u32, _pad: u32 }
#[repr(C)] #[derive(Copy, Clone, PartialEq)] pub struct GpuLU3 { pub lu: Matrix4x3<f32>, pub ia: SVector<u32, 4>, pub ib: SVector<u32, 3>, pub len: u32 }
#[repr(C)] #[derive(Copy, Clone, PartialEq)] pub struct GpuLU4 { pub lu: Matrix4<f32>, pub ia: SVector<u32, 4>, pub ib: SVector<u32, 4>, pub len: u32, _pad: [u32; 3] }
/// qr2.rs:9-20 / qr3.rs:9-20 / qr4.rs:9-20
#[repr(C)] #[derive(Copy, Clone, Debug)] pub struct GpuQR2 { pub q: Matrix2<f32>, pub r: Matrix2<f32> }
#[repr(C)] #[derive(Copy, Clone, Debug)] pub struct GpuQR3 { pub q: Matrix4x3<f32>, pub r: Matrix4x3<f32> }
#[repr(C)] #[derive(Copy, Clone, Debug)] pub struct GpuQR4 { pub q: Matrix4<f32>, pub r: Matrix4<f32> }
/// eig2.rs:10-19 / eig3.rs:11-21 / eig4.rs:12-22
#[repr(C)] #[derive(Copy, Clone, Debug)] pub struct GpuSymmetricEigen2 { pub eigenvectors: Matrix2<f32>, pub eigenvalues: Vector2<f32> }
#[repr(C)] #[derive(Copy, Clone, Debug)] pub struct GpuSymmetricEigen3 { pub eigenvectors: Matrix4x3<f32>, pub eigenvalues: Vector4<f32> }
#[repr(C)] #[derive(Copy, Clone, Debug)] pub struct GpuSymmetricEigen4 { pub eigenvectors: Matrix4<f32>, pub eigenvalues: Vector4<f32> }
/// svd2.rs:9-19, svd3.rs:10-23
#[repr(C)] #[derive(Copy, Clone)] pub struct GpuSvd2 { pub u: Matrix2<f32>, pub s: Vector2<f32>, pub vt: Matrix2<f32> }
#[repr(C)] #[derive(Copy, Clone)] pub struct GpuSvd3 { pub u: Matrix4x3<f32>, pub s: Vector4<f32>, pub vt: Matrix4x3<f32> }

macro_rules! geometry_shader {
    ($(#[$doc:meta])* $name:ident, $op:expr, $dim:expr, $mat:ty, $out:ty) => {
        $(#[$doc])*
        pub struct $name;
        // what `#[derive(Shader)]` generates in the reference (cholesky.rs:21-34 ...): the `wgcore::Shader` impl
        wgcore::impl_shader!($name, concat!("wgebra/src/geometry/", stringify!($name), ".wgsl"), "wgmath_b200/csrc/geometry.cuh", |_device| {
            debug_assert_eq!(unsafe { sys::wgb_geometry_in_bytes($dim) } as usize, core::mem::size_of::<$mat>());
            debug_assert_eq!(unsafe { sys::wgb_geometry_out_bytes($op, $dim) } as usize, core::mem::size_of::<$out>());
            $name
        });
        impl $name {
            /// Inherent twin of `Shader::from_device`, so the call compiles with or without the trait in scope.
            pub fn from_device(device: &Device) -> Result<Self, ComposerError> {
                <Self as wgcore::Shader>::from_device(device)
            }
            /// `outputs[i] = f(inputs[i])`: `KernelDispatch::new(device, pass, &pipeline).bind0([inputs.buffer(),
            /// outputs.buffer()]).dispatch(inputs.len())` of the reference's tests.
            pub fn dispatch(&self, _device: &Device, pass: &mut ComputePass, inputs: &GpuVector<$mat>, outputs: &GpuVector<$out>) {
                sys::check(unsafe { sys::wgb_geometry_batch(pass.raw(), $op, $dim, inputs.buffer().raw(), 0, outputs.buffer().raw(), 0, inputs.len()) });
            }
        }
    };
}

geometry_shader!(/// cholesky.rs:21-24
    WgCholesky2, CHOLESKY, 2, Matrix2<f32>, Matrix2<f32>);
geometry_shader!(/// cholesky.rs:26-29
    WgCholesky3, CHOLESKY, 3, Matrix4x3<f32>, Matrix4x3<f32>);
geometry_shader!(/// cholesky.rs:31-34
    WgCholesky4, CHOLESKY, 4, Matrix4<f32>, Matrix4<f32>);
geometry_shader!(/// lu.rs:65-68
    WgLU2, LU, 2, Matrix2<f32>, GpuLU2);
geometry_shader!(/// lu.rs:70-73
    WgLU3, LU, 3, Matrix4x3<f32>, GpuLU3);
geometry_shader!(/// lu.rs:75-78
    WgLU4, LU, 4, Matrix4<f32>, GpuLU4);
geometry_shader!(/// qr2.rs:22-25
    WgQR2, QR, 2, Matrix2<f32>, GpuQR2);
geometry_shader!(/// qr3.rs:22-25
    WgQR3, QR, 3, Matrix4x3<f32>, GpuQR3);
geometry_shader!(/// qr4.rs:22-25
    WgQR4, QR, 4, Matrix4<f32>, GpuQR4);
geometry_shader!(/// eig2.rs:23-26
    WgSymmetricEigen2, SYMMETRIC_EIGEN, 2, Matrix2<f32>, GpuSymmetricEigen2);
geometry_shader!(/// eig3.rs:23-26
    WgSymmetricEigen3, SYMMETRIC_EIGEN, 3, Matrix4x3<f32>, GpuSymmetricEigen3);
geometry_shader!(/// eig4.rs:24-27
    WgSymmetricEigen4, SYMMETRIC_EIGEN, 4, Matrix4<f32>, GpuSymmetricEigen4);
geometry_shader!(/// svd2.rs:21-24
    WgSvd2, SVD, 2, Matrix2<f32>, GpuSvd2);
geometry_shader!(/// svd3.rs:24-27
    WgSvd3, SVD, 3, Matrix4x3<f32>, GpuSvd3);

/// inv.rs:3-8: one shader holding inv2 / inv3 / inv4 (inv.wgsl:8-88).
pub struct WgInv;
wgcore::impl_shader!(WgInv, "wgebra/src/geometry/inv.wgsl", "wgmath_b200/csrc/geometry.cuh", |_device| WgInv);
impl WgInv {
    /// Inherent twin of `Shader::from_device`.
    pub fn from_device(device: &Device) -> Result<Self, ComposerError> { <Self as wgcore::Shader>::from_device(device) }
    pub fn dispatch2(&self, pass: &mut ComputePass, inputs: &GpuVector<Matrix2<f32>>, outputs: &GpuVector<Matrix2<f32>>) {
        sys::check(unsafe { sys::wgb_geometry_batch(pass.raw(), INV, 2, inputs.buffer().raw(), 0, outputs.buffer().raw(), 0, inputs.len()) });
    }
    pub fn dispatch3(&self, pass: &mut ComputePass, inputs: &GpuVector<Matrix4x3<f32>>, outputs: &GpuVector<Matrix4x3<f32>>) {
        sys::check(unsafe { sys::wgb_geometry_batch(pass.raw(), INV, 3, inputs.buffer().raw(), 0, outputs.buffer().raw(), 0, inputs.len()) });
    }
    pub fn dispatch4(&self, pass: &mut ComputePass, inputs: &GpuVector<Matrix4<f32>>, outputs: &GpuVector<Matrix4<f32>>) {
        sys::check(unsafe { sys::wgb_geometry_batch(pass.raw(), INV, 4, inputs.buffer().raw(), 0, outputs.buffer().raw(), 0, inputs.len()) });
    }
}
