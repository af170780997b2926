//! `wgebra::linalg` over the C ABI (reference: crates/wgebra/src/lib.rs:1-7, linalg/mod.rs:1-13).  The `geometry` and
//! `utils` WGSL function libraries are out of scope.  NOT COMPILED here (../README.md).
pub mod linalg;
pub use linalg::*;
