//! Source-compatible subset of `wgcore` (reference: crates/wgcore/src/lib.rs:5-33) for the linalg hot path:
//! `gpu`, `shapes`, `tensor`, `kernel`, `timestamps`, and the `Shader` trait (`shader`, with the `hot_reloading` state type its
//! signatures name).  Shader *composition* (`composer`, `utils`, the derive macro) is absent: the CUDA kernels are precompiled;
//! `impl_shader!` writes out what the derive would generate.  NOT COMPILED here (../README.md).
pub mod gpu;
pub mod hot_reloading;
pub mod kernel;
pub mod shader;
pub mod shapes;
pub mod tensor;
pub mod timestamps;

pub use bytemuck::Pod;
pub use shader::{Shader, ShaderRegistry};

/// Third-party re-exports, as `wgcore::re_exports` in the reference (lib.rs:23-33).
pub mod re_exports {
    pub use bytemuck;
    pub use wgpu::{self, Device};
}
