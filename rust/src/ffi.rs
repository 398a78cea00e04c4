//! Raw bindings of include/b200mm.h and include/wgpu_mm_c.h -- every exported function, in header order.
//! UNCOMPILED in this repository (no Rust toolchain); tests/test_host.py checks that every `B200MM_API` / `WGPUMM_API`
//! symbol is bound here and that the kernel-id / status / flag constants equal the header's.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_uint, c_void};

#[repr(C)]
pub struct b200mm_ctx {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct b200mm_buffer {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct b200mm_kernel {
    _opaque: [u8; 0],
}

/// b200mm_kernel_params: what Tera used to inject into the WGSL text (src/gemm.rs:24-29, src/gemv.rs:26-31)
#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct b200mm_kernel_params {
    pub workgroup_size: [u32; 3],
    pub absmax: c_float,
    pub batch: u32,
    pub flags: u32,
    pub tune: [u32; 4],
    pub group_k: u32,
}

/// wgpumm_report (include/wgpu_mm_c.h)
#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct wgpumm_report {
    pub max_abs_err: c_double,
    pub max_rel_err_f64: c_double,
    pub kernel_ms: c_double,
    pub wall_ns: c_double,
    pub gflops: c_double,
    pub kernel_gflops: c_double,
    pub kernel_gbps: c_double,
    pub seed: u64,
    pub grid: [u32; 3],
    pub block: [u32; 3],
    pub rotated: c_int,
}

// b200mm_status
pub const B200MM_OK: c_int = 0;
pub const B200MM_ERR_INVALID: c_int = -1;
pub const B200MM_ERR_CUDA: c_int = -2;
pub const B200MM_ERR_NO_DEVICE: c_int = -3;
pub const B200MM_ERR_LIMITS: c_int = -4;
pub const B200MM_ERR_UNSUPPORTED: c_int = -5;
pub const B200MM_ERR_TOLERANCE: c_int = -6;

// b200mm_kernel_id: 1..11 faithful ports of the WGSL shaders, 32.. the B200-native kernels
pub const B200MM_K_GEMM_1: c_int = 1;
pub const B200MM_K_GEMM_1V: c_int = 2;
pub const B200MM_K_GEMM_2: c_int = 3;
pub const B200MM_K_GEMM_3: c_int = 4;
pub const B200MM_K_GEMM_4: c_int = 5;
pub const B200MM_K_GEMM_5: c_int = 6;
pub const B200MM_K_GEMM_WONNX: c_int = 7;
pub const B200MM_K_BRAM: c_int = 8;
pub const B200MM_K_BRAM8X8: c_int = 9;
pub const B200MM_K_GEMM3: c_int = 10;
pub const B200MM_K_QGEMV_1: c_int = 11;
pub const B200MM_K_SGEMM_SIMT: c_int = 32;
pub const B200MM_K_SGEMM_TC3X: c_int = 33;
pub const B200MM_K_GEMV_F32: c_int = 34;
pub const B200MM_K_QGEMV_SINT8: c_int = 35;

// B200MM_F_* flags
pub const B200MM_F_NONE: u32 = 0;
pub const B200MM_F_TC3X_1X: u32 = 0x1;
pub const B200MM_F_PEER_STORE: u32 = 0x2;
pub const B200MM_F_SEQUENTIAL_K: u32 = 0x4;
pub const B200MM_F_AUTOTUNE: u32 = 0x8;
pub const B200MM_F_CONST_B: u32 = 0x10;

extern "C" {
    // ---- library ----
    pub fn b200mm_version() -> *const c_char;
    pub fn b200mm_device_count() -> c_int;
    pub fn b200mm_last_error(ctx: *const b200mm_ctx) -> *const c_char;

    // ---- device bring-up: gpu_handle, src/harness.rs:87-101 ----
    pub fn b200mm_ctx_create(device_ordinal: c_int, out: *mut *mut b200mm_ctx) -> c_int;
    pub fn b200mm_ctx_destroy(ctx: *mut b200mm_ctx) -> c_int;
    pub fn b200mm_ctx_device_info(ctx: *const b200mm_ctx, sm_count: *mut c_int, cc_major: *mut c_int, cc_minor: *mut c_int,
                                  global_mem_bytes: *mut usize, name: *mut c_char, name_len: usize) -> c_int;
    pub fn b200mm_ctx_set_stream(ctx: *mut b200mm_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn b200mm_ctx_stream(ctx: *const b200mm_ctx) -> *mut c_void;
    pub fn b200mm_sync(ctx: *mut b200mm_ctx) -> c_int;
    pub fn b200mm_ctx_launch_count(ctx: *const b200mm_ctx) -> u64;

    // ---- buffers: create_buffer_init src/harness.rs:135,158; to_cpu :289-302 ----
    pub fn b200mm_buffer_create(ctx: *mut b200mm_ctx, bytes: usize, out: *mut *mut b200mm_buffer) -> c_int;
    pub fn b200mm_buffer_create_init(ctx: *mut b200mm_ctx, host: *const c_void, bytes: usize, out: *mut *mut b200mm_buffer) -> c_int;
    pub fn b200mm_buffer_wrap(ctx: *mut b200mm_ctx, device_ptr: *mut c_void, bytes: usize, out: *mut *mut b200mm_buffer) -> c_int;
    pub fn b200mm_buffer_free(ctx: *mut b200mm_ctx, buf: *mut b200mm_buffer) -> c_int;
    pub fn b200mm_buffer_device_ptr(buf: *const b200mm_buffer) -> *mut c_void;
    pub fn b200mm_buffer_bytes(buf: *const b200mm_buffer) -> usize;
    pub fn b200mm_buffer_write(ctx: *mut b200mm_ctx, buf: *mut b200mm_buffer, offset: usize, host: *const c_void, bytes: usize) -> c_int;
    pub fn b200mm_buffer_read(ctx: *mut b200mm_ctx, buf: *const b200mm_buffer, offset: usize, host: *mut c_void, bytes: usize) -> c_int;
    pub fn b200mm_buffer_read_2d(ctx: *mut b200mm_ctx, buf: *const b200mm_buffer, offset: usize, src_pitch: usize, host: *mut c_void,
                                 dst_pitch: usize, width_bytes: usize, rows: usize) -> c_int;
    pub fn b200mm_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn b200mm_host_free(p: *mut c_void) -> c_int;
    pub fn b200mm_buffer_fill_weights(ctx: *mut b200mm_ctx, buf: *mut b200mm_buffer, seed: u64, offset: u64, n: usize) -> c_int;
    pub fn b200mm_buffer_fill_weights_2d(ctx: *mut b200mm_ctx, buf: *mut b200mm_buffer, seed: u64, offset: u64, rows: usize, cols: usize,
                                         src_ld: usize, src_col0: usize) -> c_int;

    // ---- kernels: shader module + pipeline, src/harness.rs:179-191 ----
    pub fn b200mm_kernel_get(ctx: *mut b200mm_ctx, kernel_id: c_int, m: usize, n: usize, k: usize, params: *const b200mm_kernel_params,
                             out: *mut *mut b200mm_kernel) -> c_int;
    pub fn b200mm_kernel_free(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel) -> c_int;
    pub fn b200mm_kernel_name(kernel_id: c_int) -> *const c_char;
    pub fn b200mm_kernel_geometry(kern: *const b200mm_kernel, grid: *mut u32, block: *mut u32) -> c_int;
    pub fn b200mm_kernel_workspace_bytes(kern: *const b200mm_kernel) -> usize;

    // ---- launch: mm, src/harness.rs:250-287 ----
    pub fn b200mm_launch(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel, a: *const b200mm_buffer, b: *const b200mm_buffer,
                         c: *mut b200mm_buffer, grid: *const u32) -> c_int;
    pub fn b200mm_launch_ptr(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel, a: *const c_void, b: *const c_void, c: *mut c_void,
                             grid: *const u32) -> c_int;
    pub fn b200mm_mm_host(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel, host_a: *const c_void, bytes_a: usize, host_b: *const c_void,
                          bytes_b: usize, host_c: *mut c_void, bytes_c: usize, d_a: *mut b200mm_buffer, d_b: *mut b200mm_buffer,
                          d_c: *mut b200mm_buffer) -> c_int;

    // ---- device-free introspection of the SGEMM_TC3X work schedule ----
    pub fn b200mm_tc3x_schedule(m: usize, n: usize, k: usize, bn: c_int, bk: c_int, sms: c_int, pure_stream_k: c_int, out: *mut c_int) -> c_int;
    pub fn b200mm_tc3x_plan(m: usize, n: usize, k: usize, sms: c_int, tune: *const u32, flags: u32, out: *mut c_int) -> c_int;
    pub fn b200mm_tc3x_schedule_cover(m: usize, n: usize, k: usize, bn: c_int, bk: c_int, sms: c_int, pure_stream_k: c_int, cover: *mut u16,
                                      cover_len: usize, max_segments_per_cta: *mut c_int, max_chains_per_cta: *mut c_int) -> c_int;
    pub fn b200mm_tc3x_schedule_replay(m: usize, n: usize, k: usize, bn: c_int, bk: c_int, sms: c_int, pure_stream_k: c_int,
                                       violations: *mut c_int, max_wait_list: *mut c_int) -> c_int;

    // ---- timing / measurement ----
    pub fn b200mm_timer_begin(ctx: *mut b200mm_ctx) -> c_int;
    pub fn b200mm_timer_end(ctx: *mut b200mm_ctx, elapsed_ms: *mut c_float) -> c_int;
    pub fn b200mm_kernel_profile_enable(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel, enable: c_int) -> c_int;
    pub fn b200mm_kernel_profile_read(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel, ms_out: *mut c_float, max_n: c_int, n_out: *mut c_int) -> c_int;
    pub fn b200mm_measure_fma_peak(ctx: *mut b200mm_ctx, packed: c_int, iters: c_int, reps: c_int, tflops_out: *mut c_double) -> c_int;
    pub fn b200mm_flush_l2(ctx: *mut b200mm_ctx) -> c_int;

    // ---- multi-GPU ----
    pub fn b200mm_ipc_export(ctx: *mut b200mm_ctx, buf: *const b200mm_buffer, handle64: *mut c_void) -> c_int;
    pub fn b200mm_ipc_import(ctx: *mut b200mm_ctx, handle64: *const c_void, bytes: usize, out: *mut *mut b200mm_buffer) -> c_int;
    pub fn b200mm_kernel_set_peers(kern: *mut b200mm_kernel, rank: c_int, world: c_int, peer_c: *const *mut c_void, ldc: usize,
                                   col_offset: usize) -> c_int;
    pub fn b200mm_kernel_set_peer_flags(kern: *mut b200mm_kernel, peer_flags: *const *mut c_void, pingpong_stride: usize, deferred: c_int) -> c_int;
    pub fn b200mm_kernel_peer_wait(ctx: *mut b200mm_ctx, kern: *mut b200mm_kernel) -> c_int;
    pub fn b200mm_kernel_peer_epoch(kern: *const b200mm_kernel) -> c_uint;
    pub fn b200mm_peer_barrier(ctx: *mut b200mm_ctx, local_flags: *mut b200mm_buffer, peer_flags: *const *mut c_void, rank: c_int,
                               world: c_int) -> c_int;
    pub fn b200mm_unshard_columns(ctx: *mut b200mm_ctx, gathered: *const c_void, c: *mut c_void, m: usize, n: usize, world: c_int) -> c_int;

    // ---- include/wgpu_mm_c.h: the C++ mirror's test list, codec and Workload arithmetic ----
    pub fn wgpumm_run_test(name: *const c_char, m: usize, n: usize, k: usize, seed: u64, device: c_int, verbose: c_int,
                           out: *mut wgpumm_report) -> c_int;
    pub fn wgpumm_run_test_ex(name: *const c_char, m: usize, n: usize, k: usize, seed: u64, device: c_int, verbose: c_int, grid: *const u32,
                              block: *const u32, quantize_b: c_int, out: *mut wgpumm_report) -> c_int;
    pub fn wgpumm_last_panic() -> *const c_char;
    pub fn wgpumm_entry_workload(name: *const c_char, m: usize, n: usize, k: usize, grid: *mut u32, block: *mut u32, kernel_id: *mut c_int) -> c_int;
    pub fn wgpumm_sint8_quantize(matrix: *const c_float, k: usize, n: usize, out: *mut u32, absmax: *mut c_float) -> c_int;
    pub fn wgpumm_sint8_dequantize(quantized: *const u32, absmax: c_float, k: usize, n: usize, out: *mut c_float) -> c_int;
    pub fn wgpumm_sint8_grouped_words(k: usize, n: usize, group_k: usize) -> usize;
    pub fn wgpumm_sint8_quantize_grouped(matrix: *const c_float, k: usize, n: usize, group_k: usize, packed: *mut u32) -> c_int;
    pub fn wgpumm_sint8_dequantize_grouped(packed: *const u32, k: usize, n: usize, group_k: usize, out: *mut c_float) -> c_int;
    pub fn wgpumm_compute_dim(work_items: usize, dim: c_int, count: *mut u32, size: *mut u32) -> c_int;
    pub fn wgpumm_workload_ceil(num: usize, div: usize) -> usize;
}

/// Every non-zero status becomes a panic carrying the library's message: the reference panics on every error.
pub unsafe fn check(ctx: *const b200mm_ctx, rc: c_int) {
    if rc != B200MM_OK {
        let msg = std::ffi::CStr::from_ptr(b200mm_last_error(ctx)).to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}
