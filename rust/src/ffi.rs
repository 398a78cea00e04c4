//! Raw bindings of include/b200mm.h (link with `-L wgpu_mm_b200/lib -l b200mm`).  UNCOMPILED in this repo (no Rust toolchain).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct b200mm_ctx { _p: [u8; 0] }
#[repr(C)] pub struct b200mm_buffer { _p: [u8; 0] }
#[repr(C)] pub struct b200mm_kernel { _p: [u8; 0] }

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct b200mm_kernel_params {
    pub workgroup_size: [u32; 3],
    pub absmax: f32,
    pub batch: u32,
    pub flags: u32,
    pub tune: [u32; 4],
    pub group_k: u32,
}

pub const B200MM_K_GEMM_1: c_int = 1;
pub const B200MM_K_GEMM_1V: c_int = 2;
pub const B200MM_K_GEMM_2: c_int = 3;
pub const B200MM_K_GEMM_3: c_int = 4;
pub const B200MM_K_GEMM_4: c_int = 5;
pub const B200MM_K_GEMM_5: c_int = 6;
pub const B200MM_K_QGEMV_1: c_int = 11;
pub const B200MM_K_SGEMM_SIMT: c_int = 32;
pub const B200MM_K_SGEMM_TC3X: c_int = 33;
pub const B200MM_K_GEMV_F32: c_int = 34;
pub const B200MM_K_QGEMV_SINT8: c_int = 35;

extern "C" {
    pub fn b200mm_last_error(ctx: *const b200mm_ctx) -> *const c_char;
    pub fn b200mm_ctx_create(device_ordinal: c_int, out: *mut *mut b200mm_ctx) -> c_int;            // gpu_handle
    pub fn b200mm_ctx_destroy(ctx: *mut b200mm_ctx) -> c_int;
    pub fn b200mm_sync(ctx: *mut b200mm_ctx) -> c_int;
    pub fn b200mm_buffer_create_init(ctx: *mut b200mm_ctx, host: *const c_void, bytes: usize,
                                     out: *mut *mut b200mm_buffer) -> c_int;                         // create_buffer_init
    pub fn b200mm_buffer_read(ctx: *mut b200mm_ctx, buf: *const b200mm_buffer, offset: usize,
                              host: *mut c_void, bytes: usize) -> c_int;                             // to_cpu
    pub fn b200mm_buffer_free(ctx: *mut b200mm_ctx, buf: *mut b200mm_buffer) -> c_int;
    pub fn b200mm_kernel_get(ctx: *mut b200mm_ctx, kernel_id: c_int, m: usize, n: usize, k: usize,
                             params: *const b200mm_kernel_params, out: *mut *mut b200mm_kernel) -> c_int; // shader module + pipeline
    pub fn b200mm_kernel_free(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel) -> c_int;
    pub fn b200mm_launch(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel, a: *const b200mm_buffer,
                         b: *const b200mm_buffer, c: *mut b200mm_buffer, grid: *const u32) -> c_int;   // mm
    pub fn b200mm_timer_begin(ctx: *mut b200mm_ctx) -> c_int;
    pub fn b200mm_timer_end(ctx: *mut b200mm_ctx, elapsed_ms: *mut f32) -> c_int;
}

/// Every non-zero status becomes a panic, preserving the reference's error convention.
pub unsafe fn check(ctx: *const b200mm_ctx, rc: c_int) {
    if rc != 0 {
        let msg = std::ffi::CStr::from_ptr(b200mm_last_error(ctx)).to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}
