//! crates/wgcore/src/gpu.rs:7-79
use std::sync::Arc;
use wgpu::{Adapter, Backends, Device, Queue};

pub struct GpuInstance { adapter: Adapter, device: Arc<Device>, queue: Queue }

impl GpuInstance {
    /// `GpuInstance::new().await` — kept `async` for source compatibility; completes immediately.
    pub async fn new() -> anyhow::Result<Self> { Self::with_backends(Backends::all()).await }
    pub async fn without_gl() -> anyhow::Result<Self> { Self::with_backends(Backends::all() & (!Backends::GL)).await }
    /// gpu.rs:24-58.  There is one backend (CUDA on sm_100a): `backends` is accepted and ignored.  The ordinal comes from
    /// `WGEBRA_B200_DEVICE` (default 0); one process per GPU sets it to its local rank.
    pub async fn with_backends(_backends: Backends) -> anyhow::Result<Self> {
        let ordinal = std::env::var("WGEBRA_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Self::with_ordinal(ordinal)
    }
    /// Extension: one process per GPU, rank r opens ordinal r.
    pub fn with_ordinal(ordinal: i32) -> anyhow::Result<Self> {
        let (device, queue) = Device::open(ordinal).map_err(|e| anyhow::anyhow!("Failed to initialize gpu adapter: {e}"))?;
        Ok(Self { adapter: Adapter { name: "NVIDIA B200 (sm_100a) through libwgebra_b200".into(), ordinal }, device: Arc::new(device), queue })
    }
    pub fn adapter(&self) -> &Adapter { &self.adapter }
    pub fn device(&self) -> &Device { &self.device }
    pub fn device_arc(&self) -> Arc<Device> { self.device.clone() }
    pub fn queue(&self) -> &Queue { &self.queue }
}
