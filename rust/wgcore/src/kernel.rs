//! crates/wgcore/src/kernel.rs:7-27 — only `CommandEncoderExt::compute_pass`; `KernelDispatch` (bind groups) has no
//! counterpart: one C-ABI call replaces `KernelDispatch::new().bind0([..]).dispatch(..)`.
use crate::timestamps::GpuTimestamps;
use wgpu::{CommandEncoder, ComputePass};

pub trait CommandEncoderExt {
    fn compute_pass<'encoder>(&'encoder mut self, label: &str, timestamps: Option<&mut GpuTimestamps>) -> ComputePass<'encoder>;
}
impl CommandEncoderExt for CommandEncoder {
    fn compute_pass<'encoder>(&'encoder mut self, label: &str, timestamps: Option<&mut GpuTimestamps>) -> ComputePass<'encoder> {
        let (b, e) = timestamps.and_then(|ts| ts.next_compute_pass_timestamp_writes()).map(|w| (w.begin_event, w.end_event))
            .unwrap_or((std::ptr::null_mut(), std::ptr::null_mut()));
        self.begin_compute_pass(label, b, e)
    }
}
