//! Source-compatible subset of `wgcore` (reference: crates/wgcore/src/lib.rs:5-33) for the linalg hot path:
//! `gpu`, `shapes`, `tensor`, `kernel`, `timestamps`.  Shader composition (`shader`, `composer`, `utils`, the derive
//! macro, hot reloading) is intentionally absent: the CUDA kernels are precompiled.  NOT COMPILED here (../README.md).
pub mod gpu;
pub mod kernel;
pub mod shapes;
pub mod tensor;
pub mod timestamps;

pub use bytemuck::Pod;

/// Third-party re-exports, as `wgcore::re_exports` in the reference (lib.rs:23-33).
pub mod re_exports {
    pub use bytemuck;
    pub use wgpu::{self, Device};
}
