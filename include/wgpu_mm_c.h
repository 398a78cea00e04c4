/*
 * wgpu_mm_c.h -- C entry points of the host-side mirror (include/wgpu_mm.hpp), so that test drivers
 * in any language can run the reference's test list (`cargo test test_gemm_5` ...) and its codec.
 * Reference anchors: src/gemm.rs:158-177 (gemm_test! list), src/gemv.rs:41-49, src/quant.rs:7-43,
 * src/workload.rs:48-68.
 */
#ifndef WGPU_MM_C_H
#define WGPU_MM_C_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define WGPUMM_API __attribute__((visibility("default")))

typedef struct wgpumm_report {
    double max_abs_err, max_rel_err_f64, kernel_ms, wall_ns, gflops, kernel_gflops, kernel_gbps;
    uint64_t seed;
    uint32_t grid[3], block[3]; /* the Workload the entry point produced */
    int rotated;
} wgpumm_report;

/* Runs entry point `name` ("gemm_1".."gemm_5", "gemm_wonnx", "bram", "bram8x8", "gemm3", "sgemm_simt",
 * "sgemm_tc3x", "qgemv_1", "qgemv_sint8", "gemv_f32") through test_harness.  M=N=K=0 keeps the
 * crate's constants.  Returns 0, or a b200mm_status (B200MM_ERR_TOLERANCE for "MAE too high");
 * wgpumm_last_panic() holds the message. */
WGPUMM_API int wgpumm_run_test(const char* name, size_t M, size_t N, size_t K, uint64_t seed, int device, int verbose,
                               wgpumm_report* out);
/* Same, but with the caller's own Workload and quantize_b, i.e. the full signature of test_harness (src/harness.rs:170-175):
 * grid / block = Workload::count / ::size (NULL = what the entry point produced; for the faithful ports they become
 * gridDim / blockDim), quantize_b 0 / 1 (< 0 = the entry point's own).  A quantize_b that does not match the kernel's B
 * operand fails like wgpu's bind-group validation would ("binding 1 type mismatch"). */
WGPUMM_API int wgpumm_run_test_ex(const char* name, size_t M, size_t N, size_t K, uint64_t seed, int device, int verbose,
                                  const uint32_t* grid, const uint32_t* block, int quantize_b, wgpumm_report* out);
WGPUMM_API const char* wgpumm_last_panic(void);

/* Workload produced by an entry point at the given dims, without touching a GPU. */
WGPUMM_API int wgpumm_entry_workload(const char* name, size_t M, size_t N, size_t K, uint32_t grid[3], uint32_t block[3],
                                     int* kernel_id);

/* src/quant.rs:7-28; out has K*N/4 words; returns 0 or B200MM_ERR_INVALID for the assert! failures. */
WGPUMM_API int wgpumm_sint8_quantize(const float* matrix, size_t K, size_t N, uint32_t* out, float* absmax);
/* src/quant.rs:30-43 */
WGPUMM_API int wgpumm_sint8_dequantize(const uint32_t* quantized, float absmax, size_t K, size_t N, float* out);
/* Per-group scales (extension of src/quant.rs:17): `packed` receives K*N/4 weight words followed by
 * ceil(K/group_k)*N f32 scales -- the B buffer of a qgemv_sint8 kernel created with params.group_k = group_k. */
WGPUMM_API size_t wgpumm_sint8_grouped_words(size_t K, size_t N, size_t group_k);
WGPUMM_API int wgpumm_sint8_quantize_grouped(const float* matrix, size_t K, size_t N, size_t group_k, uint32_t* packed);
WGPUMM_API int wgpumm_sint8_dequantize_grouped(const uint32_t* packed, size_t K, size_t N, size_t group_k, float* out);
/* src/workload.rs:48-68; dim 0/1/2 = X/Y/Z; returns B200MM_ERR_LIMITS for "Compute limits exceeded". */
WGPUMM_API int wgpumm_compute_dim(size_t work_items, int dim, uint32_t* count, uint32_t* size);
WGPUMM_API size_t wgpumm_workload_ceil(size_t num, size_t div);

#ifdef __cplusplus
}
#endif
#endif
