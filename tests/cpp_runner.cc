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

int main(int argc, char** argv) {
    const char* all[] = {"gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5", "gemm_wonnx", "bram", "bram8x8",
                         "gemm3", "sgemm_simt", "sgemm_tc3x", "qgemv_1", "qgemv_sint8", "gemv_f32"};
    std::vector<std::string> names;
    size_t M = 0, N = 0, K = 0;
    if (argc >= 2) names.push_back(argv[1]);
    if (argc >= 5) M = strtoull(argv[2], 0, 10), N = strtoull(argv[3], 0, 10), K = strtoull(argv[4], 0, 10);
    if (names.empty()) {
        names.push_back("test_qdq");
        for (auto* n : all) names.push_back(std::string("test_") + n);
    }
    int failed = 0;
    for (auto& t : names) {
        int rc;
        if (t == "test_qdq") {
            rc = test_qdq();
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
