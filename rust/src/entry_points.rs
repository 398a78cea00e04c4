//! Entry points with the upstream signatures (src/gemm.rs:16-150, src/gemv.rs:17-33): they still pick the tile constants
//! and build the Workload; the `String` they return is now the kernel name that `test_harness` maps to a
//! b200mm_kernel_id, instead of rendered WGSL.  UNCOMPILED sketch -- the compiled, tested equivalent is
//! wgpu_mm_b200/csrc/host/entry_points.cc.
use crate::ffi::*;

pub struct KernelSpec {
    pub kernel_id: i32,
    pub name: &'static str,
    pub absmax: f32,
}

pub fn kernel_id_of(shader: &str) -> i32 {
    match shader {
        "gemm_1" => B200MM_K_GEMM_1,
        "gemm_1v" => B200MM_K_GEMM_1V,
        "gemm_2" => B200MM_K_GEMM_2,
        "gemm_3" => B200MM_K_GEMM_3,
        "gemm_4" => B200MM_K_GEMM_4,
        "gemm_5" => B200MM_K_GEMM_5,
        "qgemv_1" => B200MM_K_QGEMV_1,
        "sgemm_simt" => B200MM_K_SGEMM_SIMT,
        "sgemm_tc3x" => B200MM_K_SGEMM_TC3X,
        "gemv_f32" => B200MM_K_GEMV_F32,
        "qgemv_sint8" => B200MM_K_QGEMV_SINT8,
        other => panic!("Template '{}' not found", other), // what tera's render().unwrap() would report
    }
}

// pub fn gemm_5(tera: &mut Tera, context: &mut Context) -> (Workload, String)
//   BM = BN = 32, BK = 16, TM = TN = 4 go into the context exactly as upstream; the Workload is
//   WorkgroupCount(ceil(N, BN), ceil(M, BM), 1) x WorkgroupSize(64, 1, 1); the returned String is "gemm_5".
// pub fn sgemm_tc3x(context: &mut Context) -> (Workload, String)        // new: the B200-native kernel, Workload advisory
