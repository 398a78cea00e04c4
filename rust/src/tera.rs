//! Stand-in for the two `tera` types the entry points take (upstream: `use tera::{Context, Tera}`, src/gemm.rs:1).  Nothing is
//! rendered any more -- the constants Tera injected into the WGSL are template parameters of the CUDA kernels -- but the
//! `Context` still records them, because `test_harness` reads M, N, K, the workgroup size and `absmax` back from it.
use std::collections::BTreeMap;

#[derive(Default)]
pub struct Tera;

#[derive(Default, Clone, Debug)]
pub struct Context {
    ints: BTreeMap<String, i64>,
    floats: BTreeMap<String, f64>,
}

/// Values an entry point may `insert` (the upstream code inserts usize, u32, i32 and one f32).
pub trait ContextValue {
    fn put(&self, key: &str, ctx: &mut Context);
}
macro_rules! int_value {
    ($($t:ty),*) => {$(impl ContextValue for $t {
        fn put(&self, key: &str, ctx: &mut Context) { ctx.ints.insert(key.to_string(), *self as i64); }
    })*};
}
int_value!(usize, u32, i32, u64, i64);
impl ContextValue for f32 {
    fn put(&self, key: &str, ctx: &mut Context) {
        ctx.floats.insert(key.to_string(), *self as f64);
    }
}

impl Context {
    pub fn new() -> Self {
        Self::default()
    }
    pub fn insert<V: ContextValue>(&mut self, key: &str, value: &V) {
        value.put(key, self);
    }
    pub fn int(&self, key: &str) -> Option<i64> {
        self.ints.get(key).copied()
    }
    pub fn float(&self, key: &str) -> Option<f64> {
        self.floats.get(key).copied()
    }
    /// `{{ M }}`-style lookup that panics like a failed `tera.render(..).unwrap()` would
    pub fn require(&self, key: &str) -> i64 {
        self.int(key).unwrap_or_else(|| panic!("Variable `{}` not found in context while rendering", key))
    }
}

impl Tera {
    /// upstream: `tera.add_raw_template(name, include_str!(..)).unwrap()`; there is no text to register any more
    pub fn add_raw_template(&mut self, _name: &str, _source: &str) -> Result<(), String> {
        Ok(())
    }
}
