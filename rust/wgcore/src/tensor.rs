//! crates/wgcore/src/tensor.rs (line numbers below refer to it): buffers, builders, views.  The view arithmetic is the
//! reference's, verbatim in meaning; only buffer creation / copy / read go through the C ABI.
use crate::gpu::GpuInstance;
use crate::shapes::ViewShape;
use bytemuck::Pod;
use encase::internal::{CreateFrom, ReadFrom, WriteInto};
use encase::{ShaderSize, ShaderType, StorageBuffer};
use nalgebra::{Dim, IsContiguous, Matrix, Storage};
use std::marker::PhantomData;
use std::mem::size_of;
use wgpu::{Buffer, BufferAddress, BufferDescriptor, BufferUsages, BufferView, CommandEncoder, Device};

#[derive(Copy, Clone)] pub struct ColumnMajor;
#[derive(Copy, Clone)] pub struct RowMajor;
pub trait MatrixOrdering: Copy + Clone { fn is_row_major() -> bool; fn is_column_major() -> bool { !Self::is_row_major() } }
impl MatrixOrdering for ColumnMajor { fn is_row_major() -> bool { false } }
impl MatrixOrdering for RowMajor { fn is_row_major() -> bool { true } }

pub type GpuScalar<T> = GpuTensor<T, 0>;
pub type GpuVector<T> = GpuTensor<T, 1>;
pub type GpuMatrix<T> = GpuTensor<T, 2>;
pub type GpuCube<T> = GpuTensor<T, 3>;
pub type GpuScalarView<'a, T, Ordering = ColumnMajor> = GpuTensorView<'a, T, Ordering, 0>;
pub type GpuVectorView<'a, T, Ordering = ColumnMajor> = GpuTensorView<'a, T, Ordering, 1>;
pub type GpuMatrixView<'a, T, Ordering = ColumnMajor> = GpuTensorView<'a, T, Ordering, 2>;
pub type GpuCubeView<'a, T, Ordering = ColumnMajor> = GpuTensorView<'a, T, Ordering, 3>;

/// :65-187
pub struct TensorBuilder<const DIM: usize> { shape: [u32; DIM], usage: BufferUsages, label: Option<String> }
impl TensorBuilder<0> { pub fn scalar(usage: BufferUsages) -> Self { Self::tensor([], usage) } }
impl TensorBuilder<1> { pub fn vector(dim: u32, usage: BufferUsages) -> Self { Self::tensor([dim], usage) } }
impl TensorBuilder<2> { pub fn matrix(nrows: u32, ncols: u32, usage: BufferUsages) -> Self { Self::tensor([nrows, ncols], usage) } }
impl<const DIM: usize> TensorBuilder<DIM> {
    pub fn tensor(shape: [u32; DIM], usage: BufferUsages) -> Self { Self { shape, usage, label: None } }
    fn len(&self) -> u64 { self.shape.into_iter().map(|s| s as u64).product() }
    pub fn label(mut self, label: String) -> Self { self.label = Some(label); self }
    pub fn build<T: Pod>(self, device: &Device) -> GpuTensor<T, DIM> {                         // :112-129
        let buffer = device.create_buffer(&BufferDescriptor { label: self.label.as_deref(), size: size_of::<T>() as u64 * self.len(), usage: self.usage, mapped_at_creation: false });
        GpuTensor { shape: self.shape, buffer, phantom: PhantomData }
    }
    pub fn build_uninit_encased<T: ShaderType>(self, device: &Device) -> GpuTensor<T, DIM> {   // :132-146
        let buffer = device.create_buffer(&BufferDescriptor { label: self.label.as_deref(), size: T::min_size().get() * self.len(), usage: self.usage, mapped_at_creation: false });
        GpuTensor { shape: self.shape, buffer, phantom: PhantomData }
    }
    pub fn build_encase<T>(self, device: &Device, data: impl AsRef<[T]>) -> GpuTensor<T, DIM> where T: ShaderType + ShaderSize + WriteInto {   // :164-173
        let mut bytes = vec![];
        let mut buffer = StorageBuffer::new(&mut bytes);
        buffer.write(data.as_ref()).unwrap();
        self.build_bytes(device, &bytes)
    }
    pub fn build_bytes<T>(self, device: &Device, data: &[u8]) -> GpuTensor<T, DIM> {          // :149-161
        GpuTensor { shape: self.shape, buffer: device.create_buffer_init(data, self.usage), phantom: PhantomData }
    }
    pub fn build_init<T: Pod>(self, device: &Device, data: &[T]) -> GpuTensor<T, DIM> {       // :175-186
        assert!(data.len() as u64 >= self.len(), "Incorrect number of elements provided for initializing Tensor.Expected at least {}, found {}", self.len(), data.len());
        let len = self.len();
        self.build_bytes::<T>(device, bytemuck::cast_slice(&data[..len as usize]))
    }
}

/// :192-399
pub struct GpuTensor<T, const DIM: usize> { shape: [u32; DIM], buffer: Buffer, phantom: PhantomData<T> }
impl<T, const DIM: usize> GpuTensor<T, DIM> {
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn len(&self) -> u64 { self.shape.into_iter().map(|s| s as u64).product() }
    pub fn bytes_len(&self) -> u64 where T: Pod { size_of::<T>() as u64 * self.len() }
    pub fn bytes_len_encased(&self) -> u64 where T: ShaderType { T::min_size().get() * self.len() }       // :217-222
    pub fn copy_from_encased(&self, encoder: &mut CommandEncoder, source: &GpuTensor<T, DIM>) where T: ShaderType {   // :235-241
        assert_eq!(self.len(), source.len());
        encoder.copy_buffer_to_buffer(&source.buffer, 0, &self.buffer, 0, self.bytes_len_encased())
    }
    pub fn copy_from(&self, encoder: &mut CommandEncoder, source: &GpuTensor<T, DIM>) where T: Pod {      // :227-233
        assert_eq!(self.len(), source.len());
        encoder.copy_buffer_to_buffer(&source.buffer, 0, &self.buffer, 0, self.bytes_len())
    }
    pub fn copy_from_view<'a, Ordering>(&self, encoder: &mut CommandEncoder, source: impl Into<GpuTensorView<'a, T, Ordering, DIM>>) where T: Pod + 'a, Ordering: 'a {   // :244-265
        let source = source.into();
        assert_eq!(source.view_shape.size[0], if DIM == 0 { 1 } else { self.shape[0] });
        encoder.copy_buffer_to_buffer(source.buffer, source.view_shape.offset as BufferAddress * size_of::<T>() as BufferAddress, &self.buffer, 0, self.bytes_len())
    }
    pub fn shape(&self) -> [u32; DIM] { self.shape }
    pub fn buffer(&self) -> &Buffer { &self.buffer }
    pub fn into_inner(self) -> Buffer { self.buffer }
    pub fn as_view<Ordering: MatrixOrdering>(&self) -> GpuTensorView<'_, T, Ordering, DIM> { self.into() }
    pub fn as_embedded_view<Ordering: MatrixOrdering, const DIM2: usize>(&self) -> GpuTensorView<'_, T, Ordering, DIM2> {   // :287-297
        assert!(DIM2 >= DIM, "Can only embed into a higher-order tensor view.");
        let mut embedded_shape = [1; DIM2];
        embedded_shape[..DIM].copy_from_slice(&self.shape[..DIM]);
        self.reshape(embedded_shape, None, None)
    }
    /// :300-325 — the map_async + poll(wait) + get_mapped_range sequence is one blocking device-to-host copy here; the view owns
    /// the bytes.  `async` kept for source compatibility.
    pub async fn read_bytes<'a>(&'a self, _device: &'a Device) -> anyhow::Result<BufferView<'a>> {
        Ok(BufferView::new(self.buffer.read_bytes()))
    }
    /// :375-384
    pub async fn read(&self, device: &Device) -> anyhow::Result<Vec<T>> where T: Pod {
        let data = self.read_bytes(device).await?;
        Ok(bytemuck::try_cast_slice(&data).map_err(|e| anyhow::anyhow!("{e}"))?.to_vec())
    }
    /// :387-399
    pub async fn read_encased(&self, device: &Device) -> anyhow::Result<Vec<T>> where T: ShaderType + ReadFrom + ShaderSize + CreateFrom {
        let data = self.read_bytes(device).await?;
        let mut result = vec![];
        let bytes: &[u8] = data.as_ref();
        StorageBuffer::new(&bytes).read(&mut result)?;
        Ok(result)
    }
    pub async fn read_to(&self, device: &Device, out: &mut [T]) -> anyhow::Result<()> where T: Pod { out.copy_from_slice(&self.read(device).await?); Ok(()) }
    pub async fn slow_read(&self, gpu: &GpuInstance) -> Vec<T> where T: Pod { self.read(gpu.device()).await.unwrap() }   // :340-355: no staging copy needed
    pub async fn slow_read_encased(&self, gpu: &GpuInstance) -> Vec<T> where T: ShaderType + ReadFrom + ShaderSize + CreateFrom { self.read_encased(gpu.device()).await.unwrap() }   // :357-372
    pub fn reshape<Ordering: MatrixOrdering, const DIM2: usize>(&self, shape: [u32; DIM2], stride: Option<u32>, stride_mat: Option<u32>) -> GpuTensorView<'_, T, Ordering, DIM2> {   // :514-541
        assert!(shape.iter().product::<u32>() <= self.shape.iter().product::<u32>());
        let mut size = [1; 3];
        size[..DIM2].copy_from_slice(&shape[..DIM2]);
        let default_stride = if Ordering::is_column_major() { shape.first().copied().unwrap_or(1) } else { shape.get(1).copied().unwrap_or(1) };
        GpuTensorView { view_shape: ViewShape { size, stride: stride.unwrap_or(default_stride), stride_mat: stride_mat.unwrap_or(shape.first().copied().unwrap_or(1) * shape.get(1).copied().unwrap_or(1)), offset: 0 }, buffer: &self.buffer, phantom: PhantomData }
    }
}
impl<'a, T, Ordering: MatrixOrdering, const DIM1: usize, const DIM2: usize> From<&'a GpuTensor<T, DIM1>> for GpuTensorView<'a, T, Ordering, DIM2> {   // :403-409
    fn from(val: &'a GpuTensor<T, DIM1>) -> Self { val.as_embedded_view() }
}

/// :416-420
#[derive(Copy, Clone)]
pub struct GpuTensorView<'a, T, Ordering, const DIM: usize> { view_shape: ViewShape, buffer: &'a Buffer, phantom: PhantomData<(T, Ordering)> }
impl<'a, T, Ordering, const DIM: usize> GpuTensorView<'a, T, Ordering, DIM> {
    pub fn shape(&self) -> ViewShape { self.view_shape }
    pub fn buffer(&self) -> &'a Buffer { self.buffer }
    fn with(&self, view_shape: ViewShape) -> GpuTensorView<'a, T, Ordering, DIM> { GpuTensorView { view_shape, buffer: self.buffer, phantom: PhantomData } }
}
impl<T> GpuVectorView<'_, T> {                                                                        // :434-463
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn len(&self) -> u32 { self.view_shape.size[0] }
    pub fn rows(&self, i: u32, nrows: u32) -> Self {
        assert!(i + nrows <= self.len(), "Rows slice range out of bounds: {}..{}", i, i + nrows);
        self.with(ViewShape { size: [nrows, 1, 1], stride: self.view_shape.stride, stride_mat: self.view_shape.stride_mat, offset: self.view_shape.offset + i })
    }
}
impl<'a, T, Ordering> GpuCubeView<'a, T, Ordering> {                                                   // :465-481
    pub fn matrix(&self, matrix_id: u32) -> GpuMatrixView<'a, T, Ordering> {
        let [nrows, ncols, nmats] = self.view_shape.size;
        assert!(matrix_id < nmats);
        GpuTensorView { view_shape: ViewShape { size: [nrows, ncols, 1], stride: self.view_shape.stride, stride_mat: 1, offset: self.view_shape.offset + self.view_shape.stride_mat * matrix_id }, buffer: self.buffer, phantom: PhantomData }
    }
}
impl<T, Ordering> GpuMatrixView<'_, T, Ordering> {                                                     // :483-511
    pub fn columns(&self, first_col: u32, ncols: u32) -> Self {
        let s = self.view_shape;
        self.with(ViewShape { size: [s.size[0], ncols, 1], stride: s.stride, stride_mat: s.stride_mat, offset: s.offset + s.stride * first_col })
    }
    pub fn rows(&self, first_row: u32, nrows: u32) -> Self {
        let s = self.view_shape;
        self.with(ViewShape { size: [nrows, s.size[1], 1], stride: s.stride, stride_mat: s.stride_mat, offset: s.offset + first_row })
    }
}
impl<T> GpuMatrix<T> {                                                                                // :544-626
    pub fn uninit(device: &Device, nrows: u32, ncols: u32, usage: BufferUsages) -> Self where T: Pod { TensorBuilder::matrix(nrows, ncols, usage).build(device) }
    pub fn uninit_encased(device: &Device, nrows: u32, ncols: u32, usage: BufferUsages) -> Self where T: ShaderType { TensorBuilder::matrix(nrows, ncols, usage).build_uninit_encased(device) }   // :553-558
    /// :561-571 — a contiguous (column-major) nalgebra matrix, uploaded as it lies in memory.
    pub fn init<R: Dim, C: Dim, S: Storage<T, R, C> + IsContiguous>(device: &Device, matrix: &Matrix<T, R, C, S>, usage: BufferUsages) -> Self where T: Pod + nalgebra::Scalar {
        TensorBuilder::matrix(matrix.nrows() as u32, matrix.ncols() as u32, usage).build_init(device, matrix.as_slice())
    }
    /// Extension: the same from a plain column-major slice.
    pub fn init_slice(device: &Device, nrows: u32, ncols: u32, column_major: &[T], usage: BufferUsages) -> Self where T: Pod { TensorBuilder::matrix(nrows, ncols, usage).build_init(device, column_major) }
    pub fn column(&self, i: u32) -> GpuVectorView<'_, T> {
        GpuTensorView { view_shape: ViewShape { size: [self.shape[0], 1, 1], stride: 1, stride_mat: 1, offset: self.shape[0] * i }, buffer: &self.buffer, phantom: PhantomData }
    }
    pub fn columns(&self, first_col: u32, ncols: u32) -> GpuMatrixView<'_, T> {
        let nrows = self.shape[0];
        GpuTensorView { view_shape: ViewShape { size: [nrows, ncols, 1], stride: nrows, stride_mat: self.shape[0] * self.shape[1], offset: first_col * nrows }, buffer: &self.buffer, phantom: PhantomData }
    }
    pub fn rows(&self, first_row: u32, nrows: u32) -> GpuMatrixView<'_, T> {
        GpuTensorView { view_shape: ViewShape { size: [nrows, self.shape[1], 1], stride: self.shape[0], stride_mat: self.shape[0] * self.shape[1], offset: first_row }, buffer: &self.buffer, phantom: PhantomData }
    }
    /// :587-598 with the offset computed from the *parent's* row count (the reference uses the slice's, which addresses
    /// the wrong column for j > 0 unless the slice spans all rows).
    pub fn slice(&self, (i, j): (u32, u32), (nrows, ncols): (u32, u32)) -> GpuMatrixView<'_, T> {
        GpuTensorView { view_shape: ViewShape { size: [nrows, ncols, 1], stride: self.shape[0], stride_mat: self.shape[0] * self.shape[1], offset: i + j * self.shape[0] }, buffer: &self.buffer, phantom: PhantomData }
    }
}
impl<T> GpuVector<T> {                                                                                // :629-682
    pub fn encase(device: &Device, vector: impl AsRef<[T]>, usage: BufferUsages) -> Self where T: ShaderType + ShaderSize + WriteInto { let v = vector.as_ref(); TensorBuilder::vector(v.len() as u32, usage).build_encase(device, v) }   // :633-639
    pub fn uninit_encased(device: &Device, len: u32, usage: BufferUsages) -> Self where T: ShaderType { TensorBuilder::vector(len, usage).build_uninit_encased(device) }   // :650-655
    pub fn uninit(device: &Device, len: u32, usage: BufferUsages) -> Self where T: Pod { TensorBuilder::vector(len, usage).build(device) }
    pub fn init(device: &Device, vector: impl AsRef<[T]>, usage: BufferUsages) -> Self where T: Pod { let v = vector.as_ref(); TensorBuilder::vector(v.len() as u32, usage).build_init(device, v) }
    pub fn rows(&self, first_row: u32, num_rows: u32) -> GpuVectorView<'_, T> {
        GpuTensorView { view_shape: ViewShape { size: [num_rows, 1, 1], stride: self.shape[0], stride_mat: self.shape[0], offset: first_row }, buffer: &self.buffer, phantom: PhantomData }
    }
}
impl<T> GpuScalar<T> {                                                                                // :684-705
    pub fn uninit_encased(device: &Device, usage: BufferUsages) -> Self where T: ShaderType { TensorBuilder::scalar(usage).build_uninit_encased(device) }   // :692-697
    pub fn uninit(device: &Device, usage: BufferUsages) -> Self where T: Pod { TensorBuilder::scalar(usage).build(device) }
    pub fn init(device: &Device, value: T, usage: BufferUsages) -> Self where T: Pod { TensorBuilder::scalar(usage).build_init(device, &[value]) }
}
