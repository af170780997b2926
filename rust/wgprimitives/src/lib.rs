//! The two integer primitives next to the linalg path over the C ABI.  NOT COMPILED here (../README.md).
//!
//! `prefix_sum` replaces crates/wgrapier/src/dynamics/prefix_sum.rs (module `wgrapier::dynamics::prefix_sum`), `radix_sort`
//! replaces crates/wgparry/src/utils/radix_sort/mod.rs (module `wgparry::utils::radix_sort`): same type names, same `dispatch`
//! signatures, same panics.  A maintainer drops each file over the original module; nothing else in those crates changes.
pub mod prefix_sum;
pub mod radix_sort;
