//! crates/wgrapier/src/dynamics/prefix_sum.rs:22-224 over `wgb_prefix_sum`.
use wgcore::tensor::GpuVector;
use wgpu::{sys, ComputePass, ComputePipeline, Device};

/// Construction can only fail for "no usable sm_100 device"; kept for signature compatibility with `#[derive(Shader)]`.
#[derive(Debug)]
pub struct ComposerError(pub String);

/// prefix_sum.rs:22-31.  The two pipelines of the reference (up/down-sweep and add-back) are kernel-family tags here.
pub struct WgPrefixSum {
    #[allow(dead_code)]
    prefix_sum: ComputePipeline,
    #[allow(dead_code)]
    add_data_grp: ComputePipeline,
}

impl WgPrefixSum {
    const THREADS: u32 = 256;

    pub fn from_device(_device: &Device) -> Result<Self, ComposerError> {
        Ok(Self { prefix_sum: ComputePipeline("prefix_sum"), add_data_grp: ComputePipeline("add_data_grp") })
    }

    /// prefix_sum.rs:49-99: in-place EXCLUSIVE prefix sum (wrapping u32 adds).  One library call instead of 2 * levels dispatches;
    /// the workspace only keeps the reference's capacity bookkeeping (its auxiliary levels live in the context).
    pub fn dispatch(&self, device: &Device, pass: &mut ComputePass, workspace: &mut PrefixSumWorkspace, data: &GpuVector<u32>) {
        workspace.reserve(device, data.len() as u32);
        let shape: sys::wgb_view_shape = data.as_view::<wgcore::tensor::ColumnMajor>().shape().into();
        sys::check(unsafe { sys::wgb_prefix_sum(pass.raw(), data.buffer().raw(), &shape) });
    }

    /// prefix_sum.rs:101-117 on a plain slice (the reference takes a nalgebra `DVector<u32>`).
    pub fn eval_cpu(&self, v: &mut [u32]) {
        let mut run = 0u32;
        for x in v.iter_mut() {
            let t = *x;
            *x = run;
            run = run.wrapping_add(t);
        }
    }
}

/// prefix_sum.rs:119-224.  `stages` mirrors the level lengths the reference would allocate (ceil(n / 256) ... 1).
#[derive(Default)]
pub struct PrefixSumWorkspace {
    pub stages: Vec<u32>,
    pub num_stages: usize,
}

impl PrefixSumWorkspace {
    pub fn new() -> Self { Self::default() }
    pub fn with_capacity(device: &Device, buffer_len: u32) -> Self {
        let mut w = Self::default();
        w.reserve(device, buffer_len);
        w
    }
    /// :185-224 (the reference does not terminate for `buffer_len == 0`; here that is one empty level)
    pub fn reserve(&mut self, _device: &Device, buffer_len: u32) {
        self.stages.clear();
        let mut stage_len = buffer_len.div_ceil(WgPrefixSum::THREADS);
        while stage_len > 1 {
            self.stages.push(stage_len);
            stage_len = stage_len.div_ceil(WgPrefixSum::THREADS);
        }
        self.stages.push(1);
        self.num_stages = self.stages.len();
    }
}
