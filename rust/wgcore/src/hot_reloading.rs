//! crates/wgcore/src/hot_reloading.rs — only the state type named by `Shader::{watch_sources, needs_reload, reload_if_changed}`:
//! there are no shader sources to watch, so the state never reports a change.  NOT COMPILED here (../../README.md).
#[derive(Default)]
pub struct HotReloadState;

impl HotReloadState {
    pub fn new() -> notify::Result<Self> {
        Ok(Self)
    }
    pub fn update_changes(&mut self) {}
}
