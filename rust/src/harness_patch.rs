//! How src/harness.rs changes (sketch, UNCOMPILED): the five wgpu touch-points become C-ABI calls; mm_ref, check's
//! tolerance gate, generate_weight_data, the 8 warm-up / 10 timed structure and the GFLOPS print stay as they are.
use crate::ffi::*;
use crate::{WorkgroupCount, Workload};

pub struct Gpu { ctx: *mut b200mm_ctx }

impl Gpu {
    /// replaces gpu_handle() (src/harness.rs:87-101)
    pub fn new() -> Self {
        let mut ctx = std::ptr::null_mut();
        unsafe { check(std::ptr::null(), b200mm_ctx_create(0, &mut ctx)) };
        Gpu { ctx }
    }
    /// replaces device.create_buffer_init (src/harness.rs:135,158)
    pub fn buffer_init<T: bytemuck::Pod>(&self, data: &[T]) -> *mut b200mm_buffer {
        let mut b = std::ptr::null_mut();
        let bytes: &[u8] = bytemuck::cast_slice(data);
        unsafe { check(self.ctx, b200mm_buffer_create_init(self.ctx, bytes.as_ptr() as _, bytes.len(), &mut b)) };
        b
    }
    /// replaces create_shader_module_unchecked + create_compute_pipeline (src/harness.rs:179-191);
    /// `shader` is now a kernel id + the constants Tera used to inject
    pub fn pipeline(&self, kernel_id: i32, dims: (usize, usize, usize), workload: &Workload, absmax: f32) -> *mut b200mm_kernel {
        let (m, n, k) = dims;
        let p = b200mm_kernel_params {
            workgroup_size: [workload.size().0, workload.size().1, workload.size().2],
            absmax, batch: 1, ..Default::default()
        };
        let mut kern = std::ptr::null_mut();
        unsafe { check(self.ctx, b200mm_kernel_get(self.ctx, kernel_id, m, n, k, &p, &mut kern)) };
        kern
    }
    /// replaces mm() + queue.submit (src/harness.rs:250-287): asynchronous, in order
    pub fn mm(&self, kern: *mut b200mm_kernel, a: *mut b200mm_buffer, b: *mut b200mm_buffer, c: *mut b200mm_buffer, count: &WorkgroupCount) {
        let grid = [count.0, count.1, count.2];
        unsafe { check(self.ctx, b200mm_launch(self.ctx, kern, a, b, c, grid.as_ptr())) };
    }
    /// replaces to_cpu() (src/harness.rs:289-302): blocking read-back
    pub fn to_cpu(&self, buf: *mut b200mm_buffer, len: usize) -> Vec<f32> {
        let mut out = vec![0f32; len];
        unsafe { check(self.ctx, b200mm_buffer_read(self.ctx, buf, 0, out.as_mut_ptr() as _, len * 4)) };
        out
    }
}
