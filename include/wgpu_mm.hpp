// wgpu_mm.hpp -- C++ host API mirroring the reference crate's public Rust API, written
// over the C ABI of b200mm.h.  Rust is not installed in the build image, and the reference is compiled
// code, so the host side lives in C++; names, argument meaning and error behaviour follow the crate:
//
//   wgpu_mm::{WorkgroupCount, WorkgroupSize, Workload, WorkloadDim}         src/workload.rs:1-69
//   wgpu_mm::gemm::{insert_matrix_dims, gemm_1 .. gemm_5}                   src/gemm.rs:9-150
//   wgpu_mm::gemv::{ABSMAX, insert_matrix_dims, qgemv_1}                    src/gemv.rs:8-33
//   wgpu_mm::quant::{sint8_quantize, sint8_dequantize}                      src/quant.rs:7-43
//   wgpu_mm::test_harness(workload, shader, dims, quantize_b)               src/harness.rs:170-248
//
// Differences that are deliberate and documented in DESIGN.md:
//   - the `shader: String` (rendered WGSL) becomes a KernelSpec naming a compiled sm_100a kernel plus the
//     constants Tera would have injected; tera::Context becomes wgpu_mm::Context (a string->integer map);
//   - shapes are still fixed per entry point by default (1024^3 / 1x1024x1024, src/gemm.rs:5-7,
//     src/gemv.rs:5-7) but insert_matrix_dims takes an optional override so the BASELINE shapes run;
//   - panics become wgpu_mm::Panic exceptions carrying the reference's message;
//   - test data is seeded (the reference uses an unseeded thread_rng, src/harness.rs:111);
//   - the benchmark loop rotates buffer roles only when M == N == K.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "b200mm.h"

// the library is built with -fvisibility=hidden; everything declared here is part of its ABI
#pragma GCC visibility push(default)
namespace wgpu_mm {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- src/workload.rs ---------------------------------------------------------------------------
struct WorkgroupCount {  // gridDim
    uint32_t x, y, z;
    WorkgroupCount(uint32_t x_, uint32_t y_, uint32_t z_) : x(x_), y(y_), z(z_) {}
};
struct WorkgroupSize {  // blockDim
    uint32_t x, y, z;
    WorkgroupSize(uint32_t x_, uint32_t y_, uint32_t z_) : x(x_), y(y_), z(z_) {}
};
enum class WorkloadDim { X, Y, Z };

class Workload {
   public:
    static constexpr size_t MAX_WORKGROUP_SIZE_X = 256;
    static constexpr size_t MAX_WORKGROUP_SIZE_Y = 256;
    static constexpr size_t MAX_WORKGROUP_SIZE_Z = 64;
    static constexpr size_t MAX_COMPUTE_WORKGROUPS_PER_DIMENSION = 65535;

    Workload(WorkgroupCount count, WorkgroupSize size) : count_(count), size_(size) {}
    const WorkgroupCount& count() const { return count_; }
    const WorkgroupSize& size() const { return size_; }
    static size_t ceil(size_t num, size_t div) { return (num + div - 1) / div; }
    // (workgroup_count, workgroup_size) for one dimension; throws Panic("Compute limits exceeded")
    static std::pair<uint32_t, uint32_t> compute_dim(size_t work_items, WorkloadDim dim);
    std::string debug() const;  // #[derive(Debug)] rendering

   private:
    WorkgroupCount count_;
    WorkgroupSize size_;
};

// ---- what replaces (tera::Context, rendered WGSL String) ---------------------------------------
using Context = std::map<std::string, int64_t>;
using Dims = std::tuple<size_t, size_t, size_t>;  // (M, N, K), src/gemm.rs:9-14

struct KernelSpec {
    int kernel_id = 0;             // b200mm_kernel_id
    std::string name;              // "gemm_5", "sgemm_tc3x", ...
    b200mm_kernel_params params{}; // constants the template would have carried
    std::string describe() const;  // what `println!("shader: {}", shader)` printed
};

namespace gemm {
// src/gemm.rs:9-14; (0,0,0) keeps the crate's constants M = N = K = 1024
Dims insert_matrix_dims(Context& context, Dims override_dims = Dims{0, 0, 0});
// faithful ports wired by the reference (src/gemm.rs:16-150)
std::pair<Workload, KernelSpec> gemm_1(Context& context);
std::pair<Workload, KernelSpec> gemm_1v(Context& context);
std::pair<Workload, KernelSpec> gemm_2(Context& context);
std::pair<Workload, KernelSpec> gemm_3(Context& context);
std::pair<Workload, KernelSpec> gemm_4(Context& context);
std::pair<Workload, KernelSpec> gemm_5(Context& context);
// orphan shaders named by north_star, given entry points (dispatch geometry per SURVEY 2.2)
std::pair<Workload, KernelSpec> gemm_wonnx(Context& context);  // shaders/gemm.wgsl + gemm_macro.wgsl
std::pair<Workload, KernelSpec> bram(Context& context);        // shaders/bram.wgsl
std::pair<Workload, KernelSpec> bram8x8(Context& context);     // shaders/bram8x8.wgsl
std::pair<Workload, KernelSpec> gemm3(Context& context);       // shaders/gemm3.wgsl
// B200-native SGEMM (Workload is advisory: the library derives its own launch configuration)
std::pair<Workload, KernelSpec> sgemm_simt(Context& context);
std::pair<Workload, KernelSpec> sgemm_tc3x(Context& context);
std::pair<Workload, KernelSpec> sgemm_tc3x_1x(Context& context);  // single-pass TF32: fails the gate at large K (panic-path tests)
}  // namespace gemm

namespace gemv {
constexpr float ABSMAX = 2.0f;  // src/gemv.rs:8 (the reference dequantises with this constant)
Dims insert_matrix_dims(Context& context, Dims override_dims = Dims{0, 0, 0});  // default (1,1024,1024)
std::pair<Workload, KernelSpec> qgemv_1(Context& context);      // src/gemv.rs:17-33
std::pair<Workload, KernelSpec> qgemv_sint8(Context& context);  // B200-native streaming kernel
std::pair<Workload, KernelSpec> qgemv_sint8_grouped(Context& context, uint32_t group_k = 128);  // per-group scales
std::pair<Workload, KernelSpec> gemv_f32(Context& context);     // B200-native fp32 GEMV (no reference shader)
}  // namespace gemv

namespace quant {
// src/quant.rs:7-28 -> (packed words, absmax).  Throws Panic on the reference's assert! failures.
std::pair<std::vector<uint32_t>, float> sint8_quantize(const std::vector<float>& matrix, size_t K, size_t N);
// src/quant.rs:30-43
std::vector<float> sint8_dequantize(const std::vector<uint32_t>& quantized, float absmax, size_t K, size_t N);

// Per-group scales (extension of src/quant.rs:17's single global absmax): one absmax per column n and
// per block of group_k consecutive rows.  `packed` is the device format the grouped sint8 GEMV consumes: the K*N/4 weight
// words in the src/quant.rs:20-26 layout, followed by ceil(K/group_k)*N f32 scales (bit-cast into the same u32 vector).
struct GroupedSint8 {
    std::vector<uint32_t> packed;
    size_t K = 0, N = 0, group_k = 0;
    size_t groups() const { return (K + group_k - 1) / group_k; }
    const uint32_t* words() const { return packed.data(); }
    const float* scales() const { return reinterpret_cast<const float*>(packed.data() + K * N / 4); }
};
GroupedSint8 sint8_quantize_grouped(const std::vector<float>& matrix, size_t K, size_t N, size_t group_k);
std::vector<float> sint8_dequantize_grouped(const GroupedSint8& q);
}  // namespace quant

// ---- src/harness.rs ------------------------------------------------------------------------------
struct HarnessReport {
    double max_abs_err = 0;      // "Max Absolute Error" vs mm_ref (src/harness.rs:64-81)
    double max_rel_err_f64 = 0;  // north_star: max |gpu - fp64| / max |fp64|
    double kernel_ms = 0;        // CUDA-event time per launch over the 10 timed launches
    double wall_ns = 0;          // reference-style: submit 10 + read back C (src/harness.rs:225-241)
    double gflops = 0;           // 2*M*N*K*10 / wall (src/harness.rs:245-247)
    double kernel_gflops = 0;    // 2*M*N*K / kernel_ms
    double kernel_gbps = 0;      // GEMV: algorithmic bytes / kernel_ms
    uint64_t seed = 0;
    bool rotated = false;        // buffer roles rotated (only legal when M == N == K)
};

struct HarnessOptions {
    uint64_t seed = 0x5EEDull;
    int device = 0;
    bool verbose = true;      // print what the reference prints
    bool check_f64 = true;    // also report the FP64 relative error
    int warmup = 8, timed = 10;  // src/harness.rs:212-237
    double gate = 1e-3;          // src/harness.rs:82
};

// Verify one launch against mm_ref (throws Panic("MAE too high") above the gate), then 8 warm-up and
// 10 timed launches, reference-style GFLOPS included.
HarnessReport test_harness(const Workload& workload, const KernelSpec& shader, Dims dims, bool quantize_b,
                           const HarnessOptions& opt = HarnessOptions{});

}  // namespace wgpu_mm
#pragma GCC visibility pop
