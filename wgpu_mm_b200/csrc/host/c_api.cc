// C entry points of the host mirror (include/wgpu_mm_c.h): the reference's test list, codec and
// Workload helpers for drivers that cannot include C++ (pytest via ctypes, a Rust FFI shim).
#include <cstring>
#include <functional>
#include <string>

#include "../../../include/wgpu_mm.hpp"
#include "../../../include/wgpu_mm_c.h"

using namespace wgpu_mm;

static thread_local std::string g_panic;

struct Entry {
    const char* name;
    std::function<std::pair<Workload, KernelSpec>(Context&)> fn;
    bool is_gemv;
    bool quantize_b;
};

static const Entry* find_entry(const char* name) {
    static const Entry table[] = {
        {"gemm_1", gemm::gemm_1, false, false},         {"gemm_1v", gemm::gemm_1v, false, false},
        {"gemm_2", gemm::gemm_2, false, false},         {"gemm_3", gemm::gemm_3, false, false},
        {"gemm_4", gemm::gemm_4, false, false},         {"gemm_5", gemm::gemm_5, false, false},
        {"gemm_wonnx", gemm::gemm_wonnx, false, false}, {"bram", gemm::bram, false, false},
        {"bram8x8", gemm::bram8x8, false, false},       {"gemm3", gemm::gemm3, false, false},
        {"sgemm_simt", gemm::sgemm_simt, false, false}, {"sgemm_tc3x", gemm::sgemm_tc3x, false, false},
        {"sgemm_tc3x_1x", gemm::sgemm_tc3x_1x, false, false},
        {"qgemv_1", gemv::qgemv_1, true, true},         {"qgemv_sint8", gemv::qgemv_sint8, true, true},
        {"gemv_f32", gemv::gemv_f32, true, false},
        {"qgemv_sint8_grouped", [](Context& c) { return gemv::qgemv_sint8_grouped(c, 128); }, true, true},
    };
    for (const auto& e : table)
        if (!strcmp(e.name, name)) return &e;
    return nullptr;
}

extern "C" const char* wgpumm_last_panic(void) { return g_panic.c_str(); }

extern "C" int wgpumm_entry_workload(const char* name, size_t M, size_t N, size_t K, uint32_t grid[3], uint32_t block[3],
                                     int* kernel_id) {
    const Entry* e = name ? find_entry(name) : nullptr;
    if (!e) {
        g_panic = std::string("unknown entry point ") + (name ? name : "(null)");
        return B200MM_ERR_INVALID;
    }
    try {
        Context ctx;
        if (e->is_gemv)
            gemv::insert_matrix_dims(ctx, Dims{M, N, K});
        else
            gemm::insert_matrix_dims(ctx, Dims{M, N, K});
        auto ws = e->fn(ctx);
        if (grid) grid[0] = ws.first.count().x, grid[1] = ws.first.count().y, grid[2] = ws.first.count().z;
        if (block) block[0] = ws.first.size().x, block[1] = ws.first.size().y, block[2] = ws.first.size().z;
        if (kernel_id) *kernel_id = ws.second.kernel_id;
    } catch (const std::exception& ex) {
        g_panic = ex.what();
        return B200MM_ERR_INVALID;
    }
    return B200MM_OK;
}

// One `cargo test test_<name>`: src/gemm.rs:158-170 (gemm_test! macro) / src/gemv.rs:41-49.
extern "C" int wgpumm_run_test(const char* name, size_t M, size_t N, size_t K, uint64_t seed, int device, int verbose,
                               wgpumm_report* out) {
    return wgpumm_run_test_ex(name, M, N, K, seed, device, verbose, nullptr, nullptr, -1, out);
}

// test_harness(workload, shader, dims, quantize_b) with the caller's own Workload and quantize_b (src/harness.rs:170-175):
// grid/block NULL = what the entry point produced; quantize_b < 0 = the entry point's own operand type.
extern "C" int wgpumm_run_test_ex(const char* name, size_t M, size_t N, size_t K, uint64_t seed, int device, int verbose,
                                  const uint32_t* grid, const uint32_t* block, int quantize_b, wgpumm_report* out) {
    const Entry* e = name ? find_entry(name) : nullptr;
    if (!e) {
        g_panic = std::string("unknown entry point ") + (name ? name : "(null)");
        return B200MM_ERR_INVALID;
    }
    try {
        Context ctx;
        const Dims dims = e->is_gemv ? gemv::insert_matrix_dims(ctx, Dims{M, N, K}) : gemm::insert_matrix_dims(ctx, Dims{M, N, K});
        auto ws = e->fn(ctx);
        if (grid || block) {
            const WorkgroupCount cnt = grid ? WorkgroupCount{grid[0], grid[1], grid[2]} : ws.first.count();
            const WorkgroupSize sz = block ? WorkgroupSize{block[0], block[1], block[2]} : ws.first.size();
            ws.first = Workload(cnt, sz);
        }
        const bool qb = quantize_b < 0 ? e->quantize_b : quantize_b != 0;
        HarnessOptions opt;
        if (seed) opt.seed = seed;
        opt.device = device;
        opt.verbose = verbose != 0;
        HarnessReport r = test_harness(ws.first, ws.second, dims, qb, opt);
        if (out) {
            out->max_abs_err = r.max_abs_err;
            out->max_rel_err_f64 = r.max_rel_err_f64;
            out->kernel_ms = r.kernel_ms;
            out->wall_ns = r.wall_ns;
            out->gflops = r.gflops;
            out->kernel_gflops = r.kernel_gflops;
            out->kernel_gbps = r.kernel_gbps;
            out->seed = r.seed;
            out->rotated = r.rotated;
            out->grid[0] = ws.first.count().x, out->grid[1] = ws.first.count().y, out->grid[2] = ws.first.count().z;
            out->block[0] = ws.first.size().x, out->block[1] = ws.first.size().y, out->block[2] = ws.first.size().z;
        }
    } catch (const Panic& p) {
        g_panic = p.what();
        if (g_panic == "MAE too high") return B200MM_ERR_TOLERANCE;
        if (g_panic == "Compute limits exceeded") return B200MM_ERR_LIMITS;
        if (g_panic.rfind("No GPU found", 0) == 0) return B200MM_ERR_NO_DEVICE;
        return B200MM_ERR_INVALID;
    } catch (const std::exception& ex) {
        g_panic = ex.what();
        return B200MM_ERR_INVALID;
    }
    return B200MM_OK;
}

extern "C" int wgpumm_sint8_quantize(const float* matrix, size_t K, size_t N, uint32_t* out, float* absmax) {
    try {
        if (!matrix || !out) throw Panic("NULL argument");
        std::vector<float> m(matrix, matrix + K * N);
        auto q = quant::sint8_quantize(m, K, N);
        memcpy(out, q.first.data(), q.first.size() * sizeof(uint32_t));
        if (absmax) *absmax = q.second;
    } catch (const std::exception& ex) {
        g_panic = ex.what();
        return B200MM_ERR_INVALID;
    }
    return B200MM_OK;
}

extern "C" int wgpumm_sint8_dequantize(const uint32_t* quantized, float absmax, size_t K, size_t N, float* out) {
    try {
        if (!quantized || !out) throw Panic("NULL argument");
        std::vector<uint32_t> q(quantized, quantized + K * N / 4);
        auto m = quant::sint8_dequantize(q, absmax, K, N);
        memcpy(out, m.data(), m.size() * sizeof(float));
    } catch (const std::exception& ex) {
        g_panic = ex.what();
        return B200MM_ERR_INVALID;
    }
    return B200MM_OK;
}

extern "C" size_t wgpumm_sint8_grouped_words(size_t K, size_t N, size_t group_k) {
    return group_k ? K * N / 4 + (K + group_k - 1) / group_k * N : 0;
}

extern "C" int wgpumm_sint8_quantize_grouped(const float* matrix, size_t K, size_t N, size_t group_k, uint32_t* packed) {
    try {
        if (!matrix || !packed) throw Panic("NULL argument");
        std::vector<float> m(matrix, matrix + K * N);
        auto q = quant::sint8_quantize_grouped(m, K, N, group_k);
        memcpy(packed, q.packed.data(), q.packed.size() * sizeof(uint32_t));
    } catch (const std::exception& ex) {
        g_panic = ex.what();
        return B200MM_ERR_INVALID;
    }
    return B200MM_OK;
}

extern "C" int wgpumm_sint8_dequantize_grouped(const uint32_t* packed, size_t K, size_t N, size_t group_k, float* out) {
    try {
        if (!packed || !out) throw Panic("NULL argument");
        quant::GroupedSint8 q;
        q.K = K;
        q.N = N;
        q.group_k = group_k;
        q.packed.assign(packed, packed + wgpumm_sint8_grouped_words(K, N, group_k));
        auto m = quant::sint8_dequantize_grouped(q);
        memcpy(out, m.data(), m.size() * sizeof(float));
    } catch (const std::exception& ex) {
        g_panic = ex.what();
        return B200MM_ERR_INVALID;
    }
    return B200MM_OK;
}

extern "C" int wgpumm_compute_dim(size_t work_items, int dim, uint32_t* count, uint32_t* size) {
    try {
        auto r = Workload::compute_dim(work_items, dim == 0 ? WorkloadDim::X : dim == 1 ? WorkloadDim::Y : WorkloadDim::Z);
        if (count) *count = r.first;
        if (size) *size = r.second;
    } catch (const Panic& p) {
        g_panic = p.what();
        return B200MM_ERR_LIMITS;
    }
    return B200MM_OK;
}

extern "C" size_t wgpumm_workload_ceil(size_t num, size_t div) { return Workload::ceil(num, div); }
