//! `wgebra::linalg` and the factorization libraries of `wgebra::geometry` over the C ABI (reference:
//! crates/wgebra/src/lib.rs:1-7, linalg/mod.rs:1-13, geometry/mod.rs:3-17).  The transform types of `geometry` (rot2, quat,
//! sim2, sim3) and `utils` are out of scope.  NOT COMPILED here (../README.md).
pub mod geometry;
pub mod linalg;
pub use geometry::*;
pub use linalg::*;
