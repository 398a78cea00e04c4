//! Safe RAII wrappers over the C ABI: what `wgpu::Device` + `wgpu::Queue`, `wgpu::Buffer` and `wgpu::ComputePipeline` were to
//! the upstream harness.  Handles are freed on drop, as the wgpu objects were at the end of `test_harness`.
use crate::ffi::*;
use crate::{WorkgroupCount, Workload};
use std::ffi::CStr;
use std::os::raw::c_void;

/// replaces `(wgpu::Device, wgpu::Queue)` from gpu_handle() (src/harness.rs:87-101): one device, one in-order stream
pub struct Device {
    ctx: *mut b200mm_ctx,
}

/// replaces `wgpu::Buffer` (STORAGE | COPY_SRC)
pub struct Buffer<'d> {
    dev: &'d Device,
    raw: *mut b200mm_buffer,
    len_bytes: usize,
}

/// replaces shader module + `wgpu::ComputePipeline` (src/harness.rs:179-191): a kernel id with the shape baked in
pub struct Pipeline<'d> {
    dev: &'d Device,
    raw: *mut b200mm_kernel,
}

impl Device {
    /// `wgpu::util::initialize_adapter_from_env_or_default(..).expect("No GPU found given preference")`; the device ordinal
    /// comes from B200MM_DEVICE (default 0) the way the backend came from WGPU_BACKEND
    pub fn new() -> Device {
        let ordinal = std::env::var("B200MM_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut ctx = std::ptr::null_mut();
        unsafe { check(std::ptr::null(), b200mm_ctx_create(ordinal, &mut ctx)) };
        Device { ctx }
    }
    pub fn raw(&self) -> *mut b200mm_ctx {
        self.ctx
    }
    pub fn name(&self) -> String {
        let mut buf = [0i8; 256];
        unsafe {
            check(self.ctx, b200mm_ctx_device_info(self.ctx, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(),
                                                    std::ptr::null_mut(), buf.as_mut_ptr() as *mut _, buf.len()));
            CStr::from_ptr(buf.as_ptr() as *const _).to_string_lossy().into_owned()
        }
    }
    /// `device.create_buffer_init(&BufferInitDescriptor { contents, .. })` (src/harness.rs:135,158)
    pub fn create_buffer_init<T: Copy>(&self, contents: &[T]) -> Buffer<'_> {
        let len_bytes = std::mem::size_of_val(contents);
        let mut raw = std::ptr::null_mut();
        unsafe { check(self.ctx, b200mm_buffer_create_init(self.ctx, contents.as_ptr() as *const c_void, len_bytes, &mut raw)) };
        Buffer { dev: self, raw, len_bytes }
    }
    /// `create_shader_module_unchecked` + `create_compute_pipeline`: `shader` is the kernel name an entry point returned
    pub fn create_compute_pipeline(&self, shader: &str, dims: (usize, usize, usize), workload: &Workload, absmax: f32) -> Pipeline<'_> {
        let (m, n, k) = dims;
        let params = b200mm_kernel_params {
            workgroup_size: [workload.size().0, workload.size().1, workload.size().2],
            absmax,
            batch: 1,
            ..Default::default()
        };
        let mut raw = std::ptr::null_mut();
        unsafe { check(self.ctx, b200mm_kernel_get(self.ctx, crate::gemm::kernel_id_of(shader), m, n, k, &params, &mut raw)) };
        Pipeline { dev: self, raw }
    }
    /// `queue.submit(vec![mm(..)])` for one dispatch (src/harness.rs:250-287): asynchronous, in order on the device's stream
    pub fn mm(&self, pipeline: &Pipeline, a: &Buffer, b: &Buffer, c: &Buffer, workgroup_count: &WorkgroupCount) {
        let grid = [workgroup_count.0, workgroup_count.1, workgroup_count.2];
        unsafe { check(self.ctx, b200mm_launch(self.ctx, pipeline.raw, a.raw, b.raw, c.raw, grid.as_ptr())) };
    }
    /// `device.poll(wgpu::Maintain::Wait)`
    pub fn poll_wait(&self) {
        unsafe { check(self.ctx, b200mm_sync(self.ctx)) };
    }
    /// CUDA-event timing around a region of the stream (the upstream harness has only `Instant::now()`)
    pub fn timed<R>(&self, f: impl FnOnce() -> R) -> (R, f32) {
        let mut ms = 0f32;
        unsafe { check(self.ctx, b200mm_timer_begin(self.ctx)) };
        let r = f();
        unsafe { check(self.ctx, b200mm_timer_end(self.ctx, &mut ms)) };
        (r, ms)
    }
}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe { b200mm_ctx_destroy(self.ctx) };
    }
}

impl<'d> Buffer<'d> {
    pub fn len_bytes(&self) -> usize {
        self.len_bytes
    }
    /// `to_cpu` (src/harness.rs:289-302): blocking read-back of the whole buffer as f32
    pub fn to_cpu(&self) -> Vec<f32> {
        let mut out = vec![0f32; self.len_bytes / 4];
        unsafe { check(self.dev.ctx, b200mm_buffer_read(self.dev.ctx, self.raw, 0, out.as_mut_ptr() as *mut c_void, self.len_bytes)) };
        out
    }
}

impl<'d> Drop for Buffer<'d> {
    fn drop(&mut self) {
        unsafe { b200mm_buffer_free(self.dev.ctx, self.raw) };
    }
}

impl<'d> Pipeline<'d> {
    pub fn geometry(&self) -> ([u32; 3], [u32; 3]) {
        let (mut g, mut b) = ([0u32; 3], [0u32; 3]);
        unsafe { check(self.dev.ctx, b200mm_kernel_geometry(self.raw, g.as_mut_ptr(), b.as_mut_ptr())) };
        (g, b)
    }
}

impl<'d> Drop for Pipeline<'d> {
    fn drop(&mut self) {
        unsafe { b200mm_kernel_free(self.dev.ctx, self.raw) };
    }
}
