//! crates/wgparry/src/utils/radix_sort/mod.rs:67-223 over `wgb_radix_sort`.
use crate::prefix_sum::ComposerError;
use wgcore::tensor::{ColumnMajor, GpuScalar, GpuVector};
use wgpu::{sys, ComputePass, Device};

/// mod.rs:82-109.  The pass uniforms, count / reduced buffers, indirect-dispatch sizes and ping-pong outputs of the reference
/// live in the library's context; the type stays so that call sites keep compiling.
pub struct RadixSortWorkspace;

impl RadixSortWorkspace {
    pub fn new(_device: &Device) -> Self { RadixSortWorkspace }
}

/// mod.rs:67-80.
pub struct RadixSort;

impl RadixSort {
    pub fn from_device(_device: &Device) -> Result<Self, ComposerError> { Ok(RadixSort) }

    /// mod.rs:111-223: stable LSD sort of the first `*n_sort` (device-resident) pairs by the low `4 * ceil(sorting_bits / 4)` key
    /// bits into `output_keys` / `output_values`.  Panics like the reference: unequal key / value lengths (:121-125), more than
    /// 32 bits (:126).
    #[allow(clippy::too_many_arguments)]
    pub fn dispatch(&self, _device: &Device, pass: &mut ComputePass, _workspace: &mut RadixSortWorkspace, input_keys: &GpuVector<u32>,
                    input_values: &GpuVector<u32>, n_sort: &GpuScalar<u32>, sorting_bits: u32, output_keys: &GpuVector<u32>,
                    output_values: &GpuVector<u32>) {
        assert_eq!(input_keys.len(), input_values.len(), "Input keys and values must have the same number of elements");
        assert!(sorting_bits <= 32, "Can only sort up to 32 bits");
        let (sk, sv): (sys::wgb_view_shape, sys::wgb_view_shape) = (input_keys.as_view::<ColumnMajor>().shape().into(), input_values.as_view::<ColumnMajor>().shape().into());
        let (so, sw): (sys::wgb_view_shape, sys::wgb_view_shape) = (output_keys.as_view::<ColumnMajor>().shape().into(), output_values.as_view::<ColumnMajor>().shape().into());
        sys::check(unsafe {
            sys::wgb_radix_sort(pass.raw(), input_keys.buffer().raw(), &sk, input_values.buffer().raw(), &sv, n_sort.buffer().raw(),
                                sorting_bits, output_keys.buffer().raw(), &so, output_values.buffer().raw(), &sw)
        });
    }
}
