//! crates/wgcore/src/shapes.rs:9-116
use wgpu::{Device, Queue};

/// shapes.rs:9-21 — `#[repr(C)]`, 24 bytes; the C ABI's `wgb_view_shape` is this very struct.
pub type ViewShape = wgpu::sys::wgb_view_shape;

/// shapes.rs:46-116.  The reference caches one uniform buffer per shape ("emulated push constants"); CUDA passes the
/// shape as a kernel parameter, so `get` hands the shape back and the cache holds nothing.
#[derive(Default)]
pub struct ViewShapeBuffers;
impl ViewShapeBuffers {
    pub fn new() -> Self { Self }
    pub fn clear_tmp(&self) {}
    pub fn put_tmp(&self, _device: &Device, _queue: &Queue, _shape: ViewShape) {}
    pub fn contains(&self, _shape: ViewShape) -> bool { true }
    pub fn get(&self, _device: &Device, shape: ViewShape) -> ViewShape { shape }
}
