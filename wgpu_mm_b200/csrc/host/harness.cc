// Host harness, mirror of src/harness.rs:17-302: bring up the device, verify one launch against the
// CPU triple loop (gate: max-abs-error <= 1e-3), then 8 warm-up and 10 timed launches and print GFLOPS.
// The GPU work goes through the C ABI of b200mm.h only.  mm_ref below is the harness's checker, exactly
// as in the reference; it is never used to produce a result.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../../include/wgpu_mm.hpp"

namespace wgpu_mm {

namespace {

struct Raii {  // device objects die with the harness like wgpu's Drop (SURVEY 8b "Ownership")
    b200mm_ctx* ctx = nullptr;
    b200mm_kernel* kern = nullptr;
    std::vector<b200mm_buffer*> bufs;
    ~Raii() {
        for (auto* b : bufs) b200mm_buffer_free(ctx, b);
        if (kern) b200mm_kernel_free(ctx, kern);
        if (ctx) b200mm_ctx_destroy(ctx);
    }
};

void expect(b200mm_ctx* ctx, int rc, const char* what) {
    if (rc == B200MM_OK) return;
    std::string msg = std::string(what) + ": " + b200mm_last_error(ctx);
    if (rc == B200MM_ERR_NO_DEVICE) msg = std::string("No GPU found given preference: ") + b200mm_last_error(nullptr);
    if (rc == B200MM_ERR_LIMITS) msg = "Compute limits exceeded";
    throw Panic(msg);
}

// src/harness.rs:103-121 with a seed added; same counter-based stream as the device generator.
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
std::vector<float> generate_weight_data(uint64_t seed, size_t M, size_t N) {
    std::vector<float> data(M * N);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < M * N; ++i) {
        const uint32_t u24 = (uint32_t)(splitmix64(seed * 0xD1342543DE82EF95ull + i) >> 40);
        const float f = (float)u24 * (1.0f / 16777216.0f);
        const float x = f * 20.0f - 10.0f;  // Uniform[-10, 10)
        data[i] = x / 50.0f;                // src/harness.rs:116
    }
    return data;
}

// src/harness.rs:17-28.  Loop order (m,k,n) with a row of accumulators: per output the products are still
// added in k = 0..K-1 order, unfused (compiled with -ffp-contract=off), so it equals the literal triple loop.
void mm_ref(const std::vector<float>& A, const std::vector<float>& B, std::vector<float>& C, Dims dims) {
    const size_t M = std::get<0>(dims), N = std::get<1>(dims), K = std::get<2>(dims);
    const size_t NB = 1024, nblk = (N + NB - 1) / NB;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (size_t m = 0; m < M; ++m)
        for (size_t b = 0; b < nblk; ++b) {
            const size_t n0 = b * NB, n1 = std::min(N, n0 + NB);
            float acc[1024];
            for (size_t n = n0; n < n1; ++n) acc[n - n0] = 0.f;
            for (size_t k = 0; k < K; ++k) {
                const float a = A[m * K + k];
                const float* brow = &B[k * N];
                for (size_t n = n0; n < n1; ++n) {
                    const float p = a * brow[n];
                    acc[n - n0] = acc[n - n0] + p;
                }
            }
            for (size_t n = n0; n < n1; ++n) C[m * N + n] = acc[n - n0];
        }
}

// north_star's second gate: max |gpu - fp64| / max |fp64| (rows sampled when the product is huge)
double rel_err_f64(const std::vector<float>& A, const std::vector<float>& B, const std::vector<float>& gpu, Dims dims) {
    const size_t M = std::get<0>(dims), N = std::get<1>(dims), K = std::get<2>(dims);
    const size_t max_rows = std::max<size_t>(1, (size_t)(4e10 / ((double)N * (double)K)));
    const size_t step = std::max<size_t>(1, M / std::min(M, max_rows));
    double max_err = 0, max_ref = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(max : max_err, max_ref)
    for (size_t m = 0; m < M; m += step) {
        std::vector<double> acc(N, 0.0);
        for (size_t k = 0; k < K; ++k) {
            const double a = A[m * K + k];
            const float* brow = &B[k * N];
            for (size_t n = 0; n < N; ++n) acc[n] += a * (double)brow[n];
        }
        for (size_t n = 0; n < N; ++n) {
            const double e = std::fabs((double)gpu[m * N + n] - acc[n]);
            if (e != e) max_err = INFINITY;
            if (e > max_err) max_err = e;
            if (std::fabs(acc[n]) > max_ref) max_ref = std::fabs(acc[n]);
        }
    }
    return max_ref > 0 ? max_err / max_ref : max_err;
}

void print_head_tail(const char* tag, const std::vector<float>& v) {
    const size_t n = v.size(), h = std::min<size_t>(16, n);
    printf("%s\n[", tag);
    for (size_t i = 0; i < h; ++i) printf("%s%g", i ? ", " : "", v[i]);
    printf("]\n...\n[");
    for (size_t i = n - h; i < n; ++i) printf("%s%g", i > n - h ? ", " : "", v[i]);
    printf("]\n");
}

bool is_quant_kernel(int id) { return id == B200MM_K_QGEMV_1 || id == B200MM_K_QGEMV_SINT8; }

}  // namespace

HarnessReport test_harness(const Workload& workload, const KernelSpec& shader, Dims dims, bool quantize_b,
                           const HarnessOptions& opt) {
    const size_t M = std::get<0>(dims), N = std::get<1>(dims), K = std::get<2>(dims);
    HarnessReport rep;
    rep.seed = opt.seed;
    Raii r;
    // gpu_handle, src/harness.rs:87-101
    expect(nullptr, b200mm_ctx_create(opt.device, &r.ctx), "gpu_handle");
    if (quantize_b != is_quant_kernel(shader.kernel_id))
        throw Panic("binding 1 type mismatch: quantize_b does not match the kernel's B operand");
    // create_shader_module_unchecked + create_compute_pipeline, src/harness.rs:179-191
    b200mm_kernel_params prm = shader.params;
    prm.workgroup_size[0] = workload.size().x;
    prm.workgroup_size[1] = workload.size().y;
    prm.workgroup_size[2] = workload.size().z;
    if (opt.verbose) printf("shader: %s\n", shader.describe().c_str());
    expect(r.ctx, b200mm_kernel_get(r.ctx, shader.kernel_id, M, N, K, &prm, &r.kern), "create_compute_pipeline");
    const uint32_t grid[3] = {workload.count().x, workload.count().y, workload.count().z};

    auto make_buffer = [&](const void* host, size_t bytes) {
        b200mm_buffer* b = nullptr;
        expect(r.ctx, b200mm_buffer_create_init(r.ctx, host, bytes, &b), "create_buffer_init");
        r.bufs.push_back(b);
        return b;
    };
    auto to_cpu = [&](b200mm_buffer* b, size_t n) {  // src/harness.rs:289-302
        std::vector<float> out(n);
        expect(r.ctx, b200mm_buffer_read(r.ctx, b, 0, out.data(), n * sizeof(float)), "Error reading buffer");
        return out;
    };

    // ---- check, src/harness.rs:30-85 ----
    {
        std::vector<float> A_cpu = generate_weight_data(opt.seed + 1, M, K);
        b200mm_buffer* A = make_buffer(A_cpu.data(), A_cpu.size() * 4);
        std::vector<float> B_cpu;
        b200mm_buffer* B;
        if (quantize_b) {
            // rand_quantized_gpu_buffer: the true absmax is discarded (src/harness.rs:134) and both sides
            // dequantise with gemv::ABSMAX (src/harness.rs:44, src/gemv.rs:30) -- kept for drop-in parity
            std::vector<float> w = generate_weight_data(opt.seed + 2, K, N);
            if (prm.group_k) {  // per-group scales travel with the weights, so both sides use the TRUE scales
                auto q = quant::sint8_quantize_grouped(w, K, N, prm.group_k);
                B = make_buffer(q.packed.data(), q.packed.size() * 4);
                B_cpu = quant::sint8_dequantize_grouped(q);
            } else {
                auto q = quant::sint8_quantize(w, K, N);
                B = make_buffer(q.first.data(), q.first.size() * 4);
                B_cpu = quant::sint8_dequantize(q.first, gemv::ABSMAX, K, N);
            }
        } else {
            B_cpu = generate_weight_data(opt.seed + 2, K, N);
            B = make_buffer(B_cpu.data(), B_cpu.size() * 4);
        }
        std::vector<float> C_cpu = generate_weight_data(opt.seed + 3, M, N);  // C starts as noise (src/harness.rs:55)
        b200mm_buffer* C = make_buffer(C_cpu.data(), C_cpu.size() * 4);
        mm_ref(A_cpu, B_cpu, C_cpu, dims);

        expect(r.ctx, b200mm_launch(r.ctx, r.kern, A, B, C, grid), "dispatch_workgroups");
        std::vector<float> gpu_out = to_cpu(C, M * N);

        float mae = 0.0f;
        bool nan = false;
        for (size_t i = 0; i < M * N; ++i) {
            const float diff = std::fabs(gpu_out[i] - C_cpu[i]);
            if (diff != diff) nan = true;
            if (diff > mae) mae = diff;
        }
        rep.max_abs_err = nan ? INFINITY : mae;
        if (opt.check_f64) rep.max_rel_err_f64 = rel_err_f64(A_cpu, B_cpu, gpu_out, dims);
        if (opt.verbose) {
            print_head_tail("GPU", gpu_out);
            print_head_tail("CPU", C_cpu);
            printf("Max Absolute Error: %g\n", rep.max_abs_err);
            if (opt.check_f64) printf("Max Relative Error vs FP64: %g\n", rep.max_rel_err_f64);
        }
        if (!(rep.max_abs_err <= opt.gate)) throw Panic("MAE too high");  // src/harness.rs:82-84
        for (auto* b : r.bufs) b200mm_buffer_free(r.ctx, b);
        r.bufs.clear();
    }

    // ---- benchmark, src/harness.rs:203-247 ----
    std::vector<float> host;
    host = generate_weight_data(opt.seed + 11, M, K);
    b200mm_buffer* A = make_buffer(host.data(), host.size() * 4);
    b200mm_buffer* B;
    if (quantize_b) {
        if (prm.group_k) {
            auto q = quant::sint8_quantize_grouped(generate_weight_data(opt.seed + 12, K, N), K, N, prm.group_k);
            B = make_buffer(q.packed.data(), q.packed.size() * 4);
        } else {
            auto q = quant::sint8_quantize(generate_weight_data(opt.seed + 12, K, N), K, N);
            B = make_buffer(q.first.data(), q.first.size() * 4);
        }
    } else {
        host = generate_weight_data(opt.seed + 12, K, N);
        B = make_buffer(host.data(), host.size() * 4);
    }
    host = generate_weight_data(opt.seed + 13, M, N);
    b200mm_buffer* C = make_buffer(host.data(), host.size() * 4);

    // The reference rotates buffer roles (mm(C,B,A), mm(A,C,B), ...; src/harness.rs:212-237), which is only
    // shape-legal for M == N == K and f32 B; otherwise the same binding is replayed.
    rep.rotated = (M == N && N == K && !quantize_b);
    b200mm_buffer* rot[10][3] = {{A, B, C}, {C, B, A}, {A, C, B}, {B, A, C}, {A, B, C},
                                 {C, B, A}, {A, C, B}, {B, A, C}, {A, B, C}, {B, A, C}};
    auto submit = [&](int n) {
        for (int i = 0; i < n; ++i) {
            b200mm_buffer** t = rot[i % 10];
            if (rep.rotated)
                expect(r.ctx, b200mm_launch(r.ctx, r.kern, t[0], t[1], t[2], grid), "dispatch_workgroups");
            else
                expect(r.ctx, b200mm_launch(r.ctx, r.kern, A, B, C, grid), "dispatch_workgroups");
        }
    };
    submit(opt.warmup);
    (void)to_cpu(C, M * N);

    const auto start = std::chrono::steady_clock::now();
    expect(r.ctx, b200mm_timer_begin(r.ctx), "timer");
    submit(opt.timed);
    float ms = 0.f;
    expect(r.ctx, b200mm_timer_end(r.ctx, &ms), "timer");
    (void)to_cpu(C, M * N);  // the reference's timed region includes the read-back (src/harness.rs:239)
    const auto elapsed = std::chrono::steady_clock::now() - start;

    rep.wall_ns = (double)std::chrono::duration_cast<std::chrono::nanoseconds>(elapsed).count();
    const double flops = (double)M * N * K * 2 * opt.timed;
    rep.gflops = (flops / 1e9) / (rep.wall_ns / 1e9);
    rep.kernel_ms = ms / opt.timed;
    rep.kernel_gflops = ((double)M * N * K * 2 / 1e9) / (rep.kernel_ms / 1e3);
    const double scale_bytes = prm.group_k ? 4.0 * (double)((K + prm.group_k - 1) / prm.group_k) * N : 0.0;
    const double bytes = (quantize_b ? (double)K * N + scale_bytes : (double)K * N * 4) + 4.0 * M * K + 4.0 * M * N;
    rep.kernel_gbps = bytes / 1e9 / (rep.kernel_ms / 1e3);
    if (opt.verbose) {
        printf("%.0f ns\n", rep.wall_ns);
        printf("%g GFLOPS\n", rep.gflops);
        printf("kernel only: %.4f ms/launch, %g GFLOPS, %g GB/s (algorithmic)\n", rep.kernel_ms, rep.kernel_gflops, rep.kernel_gbps);
    }
    return rep;
}

}  // namespace wgpu_mm
