//! crates/wgcore/src/timestamps.rs:9-248 over CUDA events: one event per timestamp slot; "ticks" are nanoseconds since slot 0.
use wgpu::sys::{self, wgb_event};
use wgpu::{BufferAsyncError, ComputePass, ComputePassTimestampWrites, Device, QuerySet, Queue};

pub struct GpuTimestamps { set: QuerySet, events: Vec<*mut wgb_event>, len: u32 }
impl GpuTimestamps {
    pub fn new(device: &wgpu::Device, capacity: u32) -> Self {
        let events = (0..capacity).map(|_| { let mut e = std::ptr::null_mut(); sys::check(unsafe { sys::wgb_event_create(device.raw(), &mut e) }); e }).collect();
        Self { set: QuerySet, events, len: 0 }
    }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn len(&self) -> usize { self.len as usize }
    pub fn query_set(&self) -> &QuerySet { &self.set }
    /// timestamps.rs:63-70: reserve a begin and an end slot for one compute pass.
    pub fn next_compute_pass_timestamp_writes(&mut self) -> Option<ComputePassTimestampWrites<'_>> {
        let [i0, i1] = self.next_query_indices()?;
        Some(ComputePassTimestampWrites { query_set: &self.set, beginning_of_pass_write_index: Some(i0), end_of_pass_write_index: Some(i1),
                                          begin_event: self.events[i0 as usize], end_event: self.events[i1 as usize] })
    }
    pub fn next_query_index(&mut self) -> Option<u32> { self.next_query_indices::<1>().map(|idx| idx[0]) }
    pub fn next_query_indices<const COUNT: usize>(&mut self) -> Option<[u32; COUNT]> {
        if self.len as usize + COUNT > self.events.len() { return None; }
        let first = self.len;
        self.len += COUNT as u32;
        Some(core::array::from_fn(|i| first + i as u32))
    }
    /// timestamps.rs:99-116: a timestamp inside a pass = an event recorded on the pass's stream at this point.
    pub fn write_next_timestamp(&mut self, compute_pass: &mut ComputePass) -> Option<u32> {
        let id = self.next_query_index()?;
        self.write_timestamp_at(compute_pass, id).then_some(id)
    }
    pub fn write_timestamp_at(&mut self, compute_pass: &mut ComputePass, query_index: u32) -> bool {
        if (query_index as usize) < self.events.len() {
            sys::check(unsafe { sys::wgb_event_record(self.events[query_index as usize], compute_pass.raw()) });
            true
        } else {
            false
        }
    }
    /// timestamps.rs:119-134: nothing to resolve, events are read directly.
    pub fn resolve(&self, _encoder: &mut wgpu::CommandEncoder) {}
    pub async fn wait_for_results_async(&self, device: &Device) -> Result<Vec<u64>, BufferAsyncError> { Ok(self.wait_for_results(device)) }
    pub async fn wait_for_results_ms_async(&self, queue: &Queue, device: &Device) -> Result<Vec<f64>, BufferAsyncError> { Ok(self.wait_for_results_ms(device, queue)) }
    /// timestamps.rs:201-224: raw ticks (here: nanoseconds relative to slot 0; the period is 1 ns).
    pub fn wait_for_results(&self, _device: &wgpu::Device) -> Vec<u64> {
        (0..self.len as usize).map(|i| if i == 0 { 0 } else { let mut ms = 0f32; sys::check(unsafe { sys::wgb_event_elapsed_ms(self.events[0], self.events[i], &mut ms) }); (ms as f64 * 1.0e6) as u64 }).collect()
    }
    /// timestamps.rs:226-230
    pub fn wait_for_results_ms(&self, device: &Device, _queue: &Queue) -> Vec<f64> { Self::timestamps_to_ms(&self.wait_for_results(device), 1.0) }
    /// timestamps.rs:235-243
    pub fn timestamps_to_ms(timestamps: &[u64], timestamp_period: f32) -> Vec<f64> {
        timestamps.iter().map(|t| *t as f64 * timestamp_period as f64 / 1_000_000.0).collect()
    }
    pub fn clear(&mut self) { self.len = 0; }
}
impl Drop for GpuTimestamps { fn drop(&mut self) { for e in &self.events { unsafe { sys::wgb_event_destroy(*e); } } } }
