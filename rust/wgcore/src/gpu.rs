//! crates/wgcore/src/gpu.rs:7-79
use std::sync::Arc;
use wgpu::{Device, Queue};

pub struct GpuInstance { device: Arc<Device>, queue: Queue }

impl GpuInstance {
    /// `GpuInstance::new().await` — kept `async` for source compatibility; completes immediately.
    pub async fn new() -> anyhow::Result<Self> { Self::with_ordinal(0) }
    pub async fn without_gl() -> anyhow::Result<Self> { Self::with_ordinal(0) }
    /// One process per GPU: rank r opens ordinal r.
    pub fn with_ordinal(ordinal: i32) -> anyhow::Result<Self> {
        let (device, queue) = Device::open(ordinal).map_err(|e| anyhow::anyhow!("Failed to initialize gpu adapter: {e}"))?;
        Ok(Self { device: Arc::new(device), queue })
    }
    pub fn device(&self) -> &Device { &self.device }
    pub fn device_arc(&self) -> Arc<Device> { self.device.clone() }
    pub fn queue(&self) -> &Queue { &self.queue }
}
