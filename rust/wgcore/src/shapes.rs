//! crates/wgcore/src/shapes.rs:9-116
use crate::tensor::MatrixOrdering;
use std::collections::HashMap;
use std::sync::{Arc, Mutex};
use wgpu::{Buffer, BufferUsages, Device, Queue};

/// shapes.rs:9-21 — `#[repr(C)]`, 24 bytes, `Pod`; byte-identical to the C ABI's `wgb_view_shape` (include/wgb200.h).
#[derive(Debug, Copy, Clone, PartialEq, Eq, Hash, bytemuck::Pod, bytemuck::Zeroable)]
#[repr(C)]
pub struct ViewShape {
    pub size: [u32; 3],
    pub stride: u32,
    pub stride_mat: u32,
    pub offset: u32,
}

impl ViewShape {
    /// shapes.rs:25-40 (kept for callers that compute it; the CUDA kernels take f32 shapes and vectorise internally).
    pub fn f32_to_vec4<Ordering: MatrixOrdering>(self) -> Self {
        let size = if Ordering::is_column_major() {
            [self.size[0] / 4, self.size[1], self.size[2]]
        } else {
            [self.size[0], self.size[1] / 4, self.size[2]]
        };
        Self { size, stride: self.stride / 4, stride_mat: self.stride_mat / 4, offset: self.offset / 4 }
    }
}

impl From<ViewShape> for wgpu::sys::wgb_view_shape {
    fn from(s: ViewShape) -> Self {
        wgpu::sys::wgb_view_shape { size: s.size, stride: s.stride, stride_mat: s.stride_mat, offset: s.offset }
    }
}

/// shapes.rs:46-116: a map between a `ViewShape` and a 24-byte uniform buffer holding it ("emulated push constants").  The CUDA
/// kernels receive the shape as a kernel parameter, so nothing ever reads these buffers; they are still created and cached with
/// the reference's semantics (`get` inserts on first use, `put_tmp` / `clear_tmp` recycle) for callers that bind them themselves.
#[derive(Default)]
pub struct ViewShapeBuffers {
    buffers: Mutex<HashMap<ViewShape, Arc<Buffer>>>,
    tmp_buffers: Mutex<HashMap<ViewShape, Arc<Buffer>>>,
    recycled: Mutex<Vec<Arc<Buffer>>>,
}

impl ViewShapeBuffers {
    pub fn new() -> Self { Self::default() }

    fn make_buffer(device: &Device, shape: ViewShape) -> Arc<Buffer> {
        Arc::new(device.create_buffer_init(bytemuck::cast_slice(&[shape]), BufferUsages::UNIFORM | BufferUsages::COPY_DST))
    }

    /// shapes.rs:65-71
    pub fn clear_tmp(&self) {
        let mut recycled = self.recycled.lock().unwrap();
        for (_, buffer) in self.tmp_buffers.lock().unwrap().drain() {
            recycled.push(buffer);
        }
    }

    /// shapes.rs:73-92
    pub fn put_tmp(&self, device: &Device, queue: &Queue, shape: ViewShape) {
        if self.contains(shape) {
            return;
        }
        let recycled = self.recycled.lock().unwrap().pop();
        let buffer = if let Some(buffer) = recycled {
            queue.write_buffer(&buffer, 0, bytemuck::cast_slice(&[shape]));
            buffer
        } else {
            Self::make_buffer(device, shape)
        };
        self.tmp_buffers.lock().unwrap().insert(shape, buffer);
    }

    /// shapes.rs:94-96
    pub fn contains(&self, shape: ViewShape) -> bool {
        self.buffers.lock().unwrap().contains_key(&shape) || self.tmp_buffers.lock().unwrap().contains_key(&shape)
    }

    /// shapes.rs:107-116
    pub fn get(&self, device: &Device, shape: ViewShape) -> Arc<Buffer> {
        if let Some(buffer) = self.tmp_buffers.lock().unwrap().get(&shape) {
            return buffer.clone();
        }
        self.buffers.lock().unwrap().entry(shape).or_insert_with(|| Self::make_buffer(device, shape)).clone()
    }
}
