// `cargo test <name>` stand-in for the C++ host mirror: runs the reference's test list
// (src/gemm.rs:172-177, src/gemv.rs:41-49, src/quant.rs:48-64) plus the B200-native entry points.
//   wgpu_mm_tests                  -> all tests at the crate's shapes
//   wgpu_mm_tests test_gemm_5      -> one test
//   wgpu_mm_tests test_sgemm_tc3x 4096 4096 4096
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/wgpu_mm.hpp"
#include "../include/wgpu_mm_c.h"

static int test_qdq() {  // src/quant.rs:48-64
    std::vector<float> m = {0.1f, -0.1f, 0.5f, -0.5f, 1.0f, -1.0f, 1.2f, -1.2f, 0.1f, -0.1f, 0.5f, -0.5f, 1.0f, -1.0f, 1.2f, -1.2f};
    auto q = wgpu_mm::quant::sint8_quantize(m, 4, 4);
    const std::vector<uint32_t> want = {3409310987u, 2172622442u, 3409310987u, 2172622442u};
    if (q.first != want) return 1;
    auto d = wgpu_mm::quant::sint8_dequantize(q.first, q.second, 4, 4);
    for (size_t i = 0; i < m.size(); ++i)
        if (!(std::fabs(m[i] - d[i]) < 0.01f)) return 1;
    return 0;
}

static int test_qdq_grouped() {  // per-group scales (extension of src/quant.rs:17): same vectors, one scale per (2-row block, column)
    std::vector<float> m = {0.1f, -0.1f, 0.5f, -0.5f, 1.0f, -1.0f, 1.2f, -1.2f, 0.1f, -0.1f, 0.5f, -0.5f, 1.0f, -1.0f, 1.2f, -1.2f};
    auto q = wgpu_mm::quant::sint8_quantize_grouped(m, 4, 4, 2);
    if (q.groups() != 2 || q.packed.size() != 4 + 2 * 4) return 1;
    const float want_scales[4] = {1.0f, 1.0f, 1.2f, 1.2f};  // column maxima of rows {0,1} and of rows {2,3}
    for (size_t g = 0; g < 2; ++g)
        for (size_t n = 0; n < 4; ++n)
            if (q.scales()[g * 4 + n] != want_scales[n]) return 1;
    // row 0 = {0.1,-0.1,0.5,-0.5} against scales {1,1,1.2,1.2}: round(12.7)=13, -13, round(52.9)=53, -53
    if (q.words()[0] != ((13u & 0xFF) | ((uint32_t)(-13 & 0xFF) << 8) | (53u << 16) | ((uint32_t)(-53 & 0xFF) << 24))) return 1;
    if (q.words()[1] != 0x817F817Fu) return 1;  // row 1 = {1,-1,1.2,-1.2}: every entry is its column's absmax -> +-127
    auto d = wgpu_mm::quant::sint8_dequantize_grouped(q);
    for (size_t i = 0; i < m.size(); ++i)
        if (!(std::fabs(m[i] - d[i]) < 0.006f)) return 1;  // half a step of the coarsest scale: 1.2 / 127 / 2
    return 0;
}

int main(int argc, char** argv) {
    const char* all[] = {"gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5", "gemm_wonnx", "bram", "bram8x8",
                         "gemm3", "sgemm_simt", "sgemm_tc3x", "qgemv_1", "qgemv_sint8", "gemv_f32", "qgemv_sint8_grouped"};
    std::vector<std::string> names;
    size_t M = 0, N = 0, K = 0;
    if (argc >= 2) names.push_back(argv[1]);
    if (argc >= 5) M = strtoull(argv[2], 0, 10), N = strtoull(argv[3], 0, 10), K = strtoull(argv[4], 0, 10);
    if (names.empty()) {
        names.push_back("test_qdq");
        names.push_back("test_qdq_grouped");
        for (auto* n : all) names.push_back(std::string("test_") + n);
    }
    int failed = 0;
    for (auto& t : names) {
        int rc;
        if (t == "test_qdq") {
            rc = test_qdq();
        } else if (t == "test_qdq_grouped") {
            rc = test_qdq_grouped();
        } else {
            std::string entry = t.rfind("test_", 0) == 0 ? t.substr(5) : t;
            wgpumm_report rep{};
            rc = wgpumm_run_test(entry.c_str(), M, N, K, 0, 0, 1, &rep);
            if (rc) printf("panicked: %s\n", wgpumm_last_panic());
        }
        printf("test %s ... %s\n", t.c_str(), rc ? "FAILED" : "ok");
        failed += rc != 0;
    }
    printf("\ntest result: %s. %zu run; %d failed\n", failed ? "FAILED" : "ok", names.size(), failed);
    return failed ? 101 : 0;
}
