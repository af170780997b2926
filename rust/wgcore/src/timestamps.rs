//! crates/wgcore/src/timestamps.rs:9-248 over CUDA events.
use wgpu::sys::{self, wgb_event};
use wgpu::{CommandEncoder, Device, Queue};

pub struct GpuTimestamps { events: Vec<*mut wgb_event>, len: usize }
impl GpuTimestamps {
    pub fn new(device: &Device, capacity: u32) -> Self {
        let events = (0..capacity).map(|_| { let mut e = std::ptr::null_mut(); sys::check(unsafe { sys::wgb_event_create(device.raw(), &mut e) }); e }).collect();
        Self { events, len: 0 }
    }
    pub fn clear(&mut self) { self.len = 0; }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    /// timestamps.rs:63-70: reserve a begin and an end slot for one compute pass.
    pub fn next_compute_pass_timestamp_writes(&mut self) -> Option<(*mut wgb_event, *mut wgb_event)> {
        if self.len + 2 > self.events.len() { return None; }
        self.len += 2;
        Some((self.events[self.len - 2], self.events[self.len - 1]))
    }
    pub fn resolve(&self, _encoder: &mut CommandEncoder) {}
    /// timestamps.rs:226-230: slot times in ms relative to slot 0.
    pub fn wait_for_results_ms(&self, _device: &Device, _queue: &Queue) -> Vec<f64> {
        (0..self.len).map(|i| if i == 0 { 0.0 } else { let mut ms = 0f32; sys::check(unsafe { sys::wgb_event_elapsed_ms(self.events[0], self.events[i], &mut ms) }); ms as f64 }).collect()
    }
}
impl Drop for GpuTimestamps { fn drop(&mut self) { for e in &self.events { unsafe { sys::wgb_event_destroy(*e); } } } }
