//! Facade of `naga_oil` 0.19 for the B200 build of wgebra: `wgcore::Shader::from_device`, `OpAssign::new`, `Reduce::new` … return
//! `Result<_, naga_oil::compose::ComposerError>` in the reference (crates/wgcore/src/shader.rs:65-72,
//! crates/wgebra/src/linalg/op_assign.rs:52, reduce.rs:71).  Downstream code names that type, so it exists here under the same
//! path; the only failure it can carry on this backend is "no usable sm_100 device".  NOT COMPILED here (../README.md).
pub mod compose {
    /// Same name and path as naga_oil's error; `inner` holds the library's message.
    #[derive(Debug, Clone)]
    pub struct ComposerError {
        pub inner: String,
    }
    impl ComposerError {
        pub fn new(msg: impl Into<String>) -> Self { Self { inner: msg.into() } }
    }
    impl std::fmt::Display for ComposerError {
        fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "{}", self.inner) }
    }
    impl std::error::Error for ComposerError {}

    /// WGSL module composer: there is no WGSL on this backend, so the composer holds nothing; it exists for the signatures of
    /// `Shader::compose` / `Shader::composer`.
    #[derive(Debug, Default, Clone)]
    pub struct Composer;
}
