//! Facade: the wgpu names that appear in the signatures of wgcore / wgebra's linalg path, backed by the C ABI of
//! `libwgebra_b200.so` (`include/wgb200.h`).  One `Device` = one CUDA device + one in-order stream (the wgpu queue).
//! NOT COMPILED in the authoring environment (no Rust toolchain) — see ../README.md.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr, CString};
use std::marker::PhantomData;
use std::sync::Arc;

/// Raw bindings (include/wgb200.h).  Status codes: 0 = OK, 2 = dimension mismatch (the reference's `assert_eq!` panics).
pub mod sys {
    use super::*;
    #[repr(C)] pub struct wgb_ctx { _p: [u8; 0] }
    #[repr(C)] pub struct wgb_pass { _p: [u8; 0] }
    #[repr(C)] pub struct wgb_buffer { _p: [u8; 0] }
    #[repr(C)] pub struct wgb_event { _p: [u8; 0] }
    /// Byte-identical to `wgcore::shapes::ViewShape` (shapes.rs:9-21).
    #[repr(C)] #[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
    pub struct wgb_view_shape { pub size: [u32; 3], pub stride: u32, pub stride_mat: u32, pub offset: u32 }

    extern "C" {
        pub fn wgb_last_error_string() -> *const c_char;
        pub fn wgb_ctx_create(device_ordinal: c_int, out: *mut *mut wgb_ctx) -> c_int;
        pub fn wgb_ctx_destroy(ctx: *mut wgb_ctx) -> c_int;
        pub fn wgb_ctx_sync(ctx: *mut wgb_ctx) -> c_int;
        pub fn wgb_pass_begin(ctx: *mut wgb_ctx, label: *const c_char, begin_ts: *mut wgb_event, end_ts: *mut wgb_event,
                              out: *mut *mut wgb_pass) -> c_int;
        pub fn wgb_pass_end(pass: *mut wgb_pass) -> c_int;
        pub fn wgb_submit(ctx: *mut wgb_ctx) -> c_int;
        pub fn wgb_buffer_create(ctx: *mut wgb_ctx, bytes: usize, usage: u32, out: *mut *mut wgb_buffer) -> c_int;
        pub fn wgb_buffer_create_init(ctx: *mut wgb_ctx, data: *const c_void, bytes: usize, usage: u32,
                                      out: *mut *mut wgb_buffer) -> c_int;
        pub fn wgb_buffer_destroy(buf: *mut wgb_buffer) -> c_int;
        pub fn wgb_buffer_write(ctx: *mut wgb_ctx, dst: *mut wgb_buffer, dst_off: usize, src: *const c_void, bytes: usize) -> c_int;
        pub fn wgb_buffer_copy(ctx: *mut wgb_ctx, pass: *mut wgb_pass, dst: *mut wgb_buffer, dst_off: usize,
                               src: *const wgb_buffer, src_off: usize, bytes: usize) -> c_int;
        pub fn wgb_buffer_read(ctx: *mut wgb_ctx, src: *const wgb_buffer, src_off: usize, dst: *mut c_void, bytes: usize) -> c_int;
        pub fn wgb_gemm_ex(pass: *mut wgb_pass, variant: c_int, out: *mut wgb_buffer, out_shape: *const wgb_view_shape,
                           m1: *const wgb_buffer, m1_shape: *const wgb_view_shape, m2: *const wgb_buffer,
                           m2_shape: *const wgb_view_shape, in_dtype: c_int, out_dtype: c_int, f32_mode: c_int) -> c_int;
        pub fn wgb_gemv(pass: *mut wgb_pass, variant: c_int, out: *mut wgb_buffer, out_shape: *const wgb_view_shape,
                        m: *const wgb_buffer, m_shape: *const wgb_view_shape, v: *const wgb_buffer,
                        v_shape: *const wgb_view_shape) -> c_int;
        /// orderings: 0 = ColumnMajor, 1 = RowMajor (wgb_ordering); op < 0: no fused element-wise step
        pub fn wgb_gemm_ord(pass: *mut wgb_pass, variant: c_int, out: *mut wgb_buffer, out_shape: *const wgb_view_shape, out_ord: c_int,
                            m1: *const wgb_buffer, m1_shape: *const wgb_view_shape, m1_ord: c_int, m2: *const wgb_buffer,
                            m2_shape: *const wgb_view_shape, m2_ord: c_int, in_dtype: c_int, out_dtype: c_int, f32_mode: c_int,
                            op: c_int, operand: *const wgb_buffer, operand_shape: *const wgb_view_shape) -> c_int;
        pub fn wgb_gemv_ord(pass: *mut wgb_pass, variant: c_int, out: *mut wgb_buffer, out_shape: *const wgb_view_shape,
                            m: *const wgb_buffer, m_shape: *const wgb_view_shape, m_ord: c_int, v: *const wgb_buffer,
                            v_shape: *const wgb_view_shape) -> c_int;
        pub fn wgb_op_assign(pass: *mut wgb_pass, op: c_int, a: *mut wgb_buffer, a_shape: *const wgb_view_shape,
                             b: *const wgb_buffer, b_shape: *const wgb_view_shape) -> c_int;
        pub fn wgb_reduce(pass: *mut wgb_pass, op: c_int, value: *const wgb_buffer, value_shape: *const wgb_view_shape,
                          result: *mut wgb_buffer) -> c_int;
        pub fn wgb_prefix_sum(pass: *mut wgb_pass, data: *mut wgb_buffer, data_shape: *const wgb_view_shape) -> c_int;
        pub fn wgb_radix_sort(pass: *mut wgb_pass, input_keys: *const wgb_buffer, input_keys_shape: *const wgb_view_shape,
                              input_values: *const wgb_buffer, input_values_shape: *const wgb_view_shape, n_sort: *const wgb_buffer,
                              sorting_bits: u32, output_keys: *mut wgb_buffer, output_keys_shape: *const wgb_view_shape,
                              output_values: *mut wgb_buffer, output_values_shape: *const wgb_view_shape) -> c_int;
        pub fn wgb_gemv_op(pass: *mut wgb_pass, variant: c_int, out: *mut wgb_buffer, out_shape: *const wgb_view_shape,
                           m: *const wgb_buffer, m_shape: *const wgb_view_shape, m_ordering: c_int, v: *const wgb_buffer,
                           v_shape: *const wgb_view_shape, op: c_int, operand: *const wgb_buffer,
                           operand_shape: *const wgb_view_shape) -> c_int;
        pub fn wgb_geometry_in_bytes(dim: c_int) -> u32;
        pub fn wgb_geometry_out_bytes(op: c_int, dim: c_int) -> u32;
        pub fn wgb_geometry_batch(pass: *mut wgb_pass, op: c_int, dim: c_int, input: *const wgb_buffer, in_first: u64,
                                  output: *mut wgb_buffer, out_first: u64, n: u64) -> c_int;
        pub fn wgb_event_create(ctx: *mut wgb_ctx, out: *mut *mut wgb_event) -> c_int;
        pub fn wgb_event_destroy(ev: *mut wgb_event) -> c_int;
        pub fn wgb_event_elapsed_ms(begin: *mut wgb_event, end: *mut wgb_event, ms: *mut f32) -> c_int;
        pub fn wgb_event_record(ev: *mut wgb_event, pass: *mut wgb_pass) -> c_int;
        pub fn wgb_gemv_reduce(pass: *mut wgb_pass, variant: c_int, reduce_op: c_int, result: *mut wgb_buffer, m: *const wgb_buffer,
                               m_shape: *const wgb_view_shape, m_ord: c_int, v: *const wgb_buffer, v_shape: *const wgb_view_shape) -> c_int;
    }

    /// Non-zero status -> panic with the library's message ("Gemm: dimension mismatch. …" for status 2, as gemm.rs:91).
    #[track_caller]
    pub fn check(status: c_int) {
        if status != 0 {
            let msg = unsafe { CStr::from_ptr(wgb_last_error_string()) }.to_string_lossy().into_owned();
            panic!("{msg}");
        }
    }
}

pub type BufferAddress = u64;

bitflags::bitflags! {
    /// Same bit values as wgpu::BufferUsages (passed to the C ABI unchanged; MAP_READ => pinned host staging buffer).
    #[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
    pub struct BufferUsages: u32 {
        const MAP_READ = 1 << 0; const MAP_WRITE = 1 << 1; const COPY_SRC = 1 << 2; const COPY_DST = 1 << 3;
        const INDEX = 1 << 4; const VERTEX = 1 << 5; const UNIFORM = 1 << 6; const STORAGE = 1 << 7;
        const INDIRECT = 1 << 8; const QUERY_RESOLVE = 1 << 9;
    }
}

struct CtxHandle(*mut sys::wgb_ctx);
unsafe impl Send for CtxHandle {}
unsafe impl Sync for CtxHandle {}
impl Drop for CtxHandle { fn drop(&mut self) { unsafe { sys::wgb_ctx_destroy(self.0); } } }

/// wgpu::Device: owns the context; cheap to clone (Arc), Send + Sync like the original.
#[derive(Clone)]
pub struct Device { ctx: Arc<CtxHandle> }
impl Device {
    /// `adapter.request_device` of gpu.rs:36-52: fails only if no sm_100 CUDA device is usable (there is no CPU fallback).
    pub fn open(ordinal: i32) -> Result<(Device, Queue), String> {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { sys::wgb_ctx_create(ordinal, &mut raw) };
        if st != 0 {
            return Err(unsafe { CStr::from_ptr(sys::wgb_last_error_string()) }.to_string_lossy().into_owned());
        }
        let dev = Device { ctx: Arc::new(CtxHandle(raw)) };
        Ok((dev.clone(), Queue { device: dev }))
    }
    pub fn raw(&self) -> *mut sys::wgb_ctx { self.ctx.0 }
    pub fn create_command_encoder(&self, _desc: &CommandEncoderDescriptor) -> CommandEncoder { CommandEncoder { device: self.clone() } }
    pub fn create_buffer(&self, desc: &BufferDescriptor) -> Buffer {
        let mut raw = std::ptr::null_mut();
        sys::check(unsafe { sys::wgb_buffer_create(self.raw(), desc.size as usize, desc.usage.bits(), &mut raw) });
        Buffer { raw, size: desc.size, device: self.clone() }
    }
    /// util::DeviceExt::create_buffer_init (tensor.rs:150-154)
    pub fn create_buffer_init(&self, contents: &[u8], usage: BufferUsages) -> Buffer {
        let mut raw = std::ptr::null_mut();
        sys::check(unsafe { sys::wgb_buffer_create_init(self.raw(), contents.as_ptr() as *const c_void, contents.len(), usage.bits(), &mut raw) });
        Buffer { raw, size: contents.len() as u64, device: self.clone() }
    }
    /// device.poll(PollType::wait()) (tensor.rs:304-312)
    pub fn poll_wait(&self) { sys::check(unsafe { sys::wgb_ctx_sync(self.raw()) }); }
}

#[derive(Default)] pub struct CommandEncoderDescriptor;
pub struct BufferDescriptor<'a> { pub label: Option<&'a str>, pub size: BufferAddress, pub usage: BufferUsages, pub mapped_at_creation: bool }

pub struct Queue { device: Device }
impl Queue {
    /// queue.submit(Some(encoder.finish())): dispatches were enqueued when recorded; this only flushes.
    pub fn submit<I: IntoIterator<Item = CommandBuffer>>(&self, _buffers: I) { sys::check(unsafe { sys::wgb_submit(self.device.raw()) }); }
    pub fn write_buffer(&self, buffer: &Buffer, offset: BufferAddress, data: &[u8]) {
        sys::check(unsafe { sys::wgb_buffer_write(self.device.raw(), buffer.raw, offset as usize, data.as_ptr() as *const c_void, data.len()) });
        self.device.poll_wait(); // `data` may be dropped by the caller right after this returns
    }
}

pub struct Buffer { raw: *mut sys::wgb_buffer, size: u64, device: Device }
unsafe impl Send for Buffer {}
unsafe impl Sync for Buffer {}
impl Buffer {
    pub fn raw(&self) -> *mut sys::wgb_buffer { self.raw }
    pub fn size(&self) -> u64 { self.size }
    pub fn device(&self) -> &Device { &self.device }
    /// Blocking read-back (the map_async + poll(wait) + get_mapped_range sequence of tensor.rs:300-325).
    pub fn read_bytes(&self) -> Vec<u8> {
        let mut out = vec![0u8; self.size as usize];
        sys::check(unsafe { sys::wgb_buffer_read(self.device.raw(), self.raw, 0, out.as_mut_ptr() as *mut c_void, out.len()) });
        out
    }
}
impl Drop for Buffer { fn drop(&mut self) { unsafe { sys::wgb_buffer_destroy(self.raw); } } }

/// wgpu::BufferView: the mapped bytes of a buffer; here an owned copy made by one blocking read (tensor.rs:300-325).
pub struct BufferView<'a> { bytes: Vec<u8>, _buffer: PhantomData<&'a Buffer> }
impl<'a> BufferView<'a> { pub fn new(bytes: Vec<u8>) -> Self { Self { bytes, _buffer: PhantomData } } }
impl std::ops::Deref for BufferView<'_> { type Target = [u8]; fn deref(&self) -> &[u8] { &self.bytes } }
impl AsRef<[u8]> for BufferView<'_> { fn as_ref(&self) -> &[u8] { &self.bytes } }

bitflags::bitflags! {
    /// wgpu::Backends (gpu.rs:20-29 `with_backends`): accepted and ignored, there is one backend.
    #[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
    pub struct Backends: u32 { const VULKAN = 1; const GL = 2; const METAL = 4; const DX12 = 8; const BROWSER_WEBGPU = 16; }
}
/// wgpu::Adapter (gpu.rs:69-71): what `nvidia-smi -L` would say about the device.
#[derive(Clone, Debug)]
pub struct Adapter { pub name: String, pub ordinal: i32 }
/// wgpu::QuerySet (timestamps.rs:44-46): the event slots live in `GpuTimestamps`; the set is a tag.
#[derive(Debug, Default)]
pub struct QuerySet;
/// wgpu::ComputePassTimestampWrites (timestamps.rs:63-70): the begin / end slots of one pass.
pub struct ComputePassTimestampWrites<'a> {
    pub query_set: &'a QuerySet,
    pub beginning_of_pass_write_index: Option<u32>,
    pub end_of_pass_write_index: Option<u32>,
    pub begin_event: *mut sys::wgb_event,
    pub end_event: *mut sys::wgb_event,
}
/// wgpu::BufferAsyncError (timestamps.rs:196-224): never produced, a read is one blocking copy.
#[derive(Debug, Clone, Copy)]
pub struct BufferAsyncError;

pub struct CommandBuffer;
pub struct CommandEncoder { device: Device }
impl CommandEncoder {
    pub fn device(&self) -> &Device { &self.device }
    pub fn begin_compute_pass(&mut self, label: &str, begin_ts: *mut sys::wgb_event, end_ts: *mut sys::wgb_event) -> ComputePass<'_> {
        let label = CString::new(label).unwrap_or_default();
        let mut raw = std::ptr::null_mut();
        sys::check(unsafe { sys::wgb_pass_begin(self.device.raw(), label.as_ptr(), begin_ts, end_ts, &mut raw) });
        ComputePass { raw, _encoder: PhantomData }
    }
    pub fn copy_buffer_to_buffer(&mut self, src: &Buffer, src_off: BufferAddress, dst: &Buffer, dst_off: BufferAddress, size: BufferAddress) {
        sys::check(unsafe { sys::wgb_buffer_copy(self.device.raw(), std::ptr::null_mut(), dst.raw, dst_off as usize, src.raw, src_off as usize, size as usize) });
    }
    pub fn finish(self) -> CommandBuffer { CommandBuffer }
}

/// A compute pass = the context's in-order stream while it is borrowed; `drop(pass)` ends it (kernel.rs:15-26).
pub struct ComputePass<'encoder> { raw: *mut sys::wgb_pass, _encoder: PhantomData<&'encoder mut CommandEncoder> }
impl ComputePass<'_> { pub fn raw(&self) -> *mut sys::wgb_pass { self.raw } }
impl Drop for ComputePass<'_> { fn drop(&mut self) { unsafe { sys::wgb_pass_end(self.raw); } } }

/// The CUDA kernels are compiled into the library: a "pipeline" is only a tag naming the kernel family.
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct ComputePipeline(pub &'static str);
