// Kernel entry points, mirror of src/gemm.rs:9-150 and src/gemv.rs:8-33: each picks the tile constants,
// fills the context with what the reference would hand to Tera, and returns (Workload, KernelSpec).
#include "../../../include/wgpu_mm.hpp"

namespace wgpu_mm {

static Dims put_dims(Context& c, Dims d, size_t M0, size_t N0, size_t K0) {
    size_t M = std::get<0>(d), N = std::get<1>(d), K = std::get<2>(d);
    if (M == 0 && N == 0 && K == 0) M = M0, N = N0, K = K0;
    c["M"] = (int64_t)M;
    c["N"] = (int64_t)N;
    c["K"] = (int64_t)K;
    return Dims{M, N, K};
}

static void need_dims(const Context& c, size_t& M, size_t& N, size_t& K) {
    auto m = c.find("M"), n = c.find("N"), k = c.find("K");
    if (m == c.end() || n == c.end() || k == c.end())
        throw Panic("Variable `M` not found in context while rendering");  // what tera's render().unwrap() reports
    M = (size_t)m->second;
    N = (size_t)n->second;
    K = (size_t)k->second;
}

static std::pair<Workload, KernelSpec> finish(Context& c, Workload w, int id) {
    // src/gemm.rs:24-28: the workgroup size goes into the context and from there into the shader text
    c["workgroup_size_x"] = w.size().x;
    c["workgroup_size_y"] = w.size().y;
    c["workgroup_size_z"] = w.size().z;
    KernelSpec s;
    s.kernel_id = id;
    s.name = b200mm_kernel_name(id);
    s.params.workgroup_size[0] = w.size().x;
    s.params.workgroup_size[1] = w.size().y;
    s.params.workgroup_size[2] = w.size().z;
    return {w, s};
}

namespace gemm {

Dims insert_matrix_dims(Context& context, Dims d) { return put_dims(context, d, 1024, 1024, 1024); }  // src/gemm.rs:5-14

std::pair<Workload, KernelSpec> gemm_1(Context& c) {  // src/gemm.rs:16-32
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(M, 16), Workload::ceil(N, 16), 1), WorkgroupSize(16, 16, 1)),
                  B200MM_K_GEMM_1);
}
std::pair<Workload, KernelSpec> gemm_1v(Context& c) {  // src/gemm.rs:34-50
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(M, 16), Workload::ceil(N, 16), 1), WorkgroupSize(16, 16 / 4, 1)),
                  B200MM_K_GEMM_1V);
}
std::pair<Workload, KernelSpec> gemm_2(Context& c) {  // src/gemm.rs:52-67
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(M, 16), Workload::ceil(N, 16), 1), WorkgroupSize(256, 1, 1)),
                  B200MM_K_GEMM_2);
}
std::pair<Workload, KernelSpec> gemm_3(Context& c) {  // src/gemm.rs:69-90
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t BLOCKSIZE = 16;
    c["BLOCKSIZE"] = BLOCKSIZE;
    return finish(c,
                  Workload(WorkgroupCount(Workload::ceil(M, BLOCKSIZE), Workload::ceil(N, BLOCKSIZE), 1),
                           WorkgroupSize(BLOCKSIZE * BLOCKSIZE, 1, 1)),
                  B200MM_K_GEMM_3);
}
std::pair<Workload, KernelSpec> gemm_4(Context& c) {  // src/gemm.rs:92-119
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t BM = 16, BN = 16, BK = 8, TM = 2;
    c["BM"] = BM, c["BN"] = BN, c["BK"] = BK, c["TM"] = TM;
    return finish(c, Workload(WorkgroupCount(Workload::ceil(N, BN), Workload::ceil(M, BM), 1), WorkgroupSize((BM * BN) / TM, 1, 1)),
                  B200MM_K_GEMM_4);
}
std::pair<Workload, KernelSpec> gemm_5(Context& c) {  // src/gemm.rs:121-150
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t BM = 32, BN = 32, BK = 16, TM = 4, TN = 4;
    c["BM"] = BM, c["BN"] = BN, c["BK"] = BK, c["TM"] = TM, c["TN"] = TN;
    return finish(c,
                  Workload(WorkgroupCount(Workload::ceil(N, BN), Workload::ceil(M, BM), 1), WorkgroupSize((BM * BN) / (TM * TN), 1, 1)),
                  B200MM_K_GEMM_5);
}

// ---- orphan shaders: no Rust wires them; geometry inferred from the shader bodies ----
std::pair<Workload, KernelSpec> gemm_wonnx(Context& c) {  // shaders/gemm_macro.wgsl:3-4: exactly M*N/16 invocations along x
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t wg = 256;
    return finish(c, Workload(WorkgroupCount(Workload::ceil(M * N / 16, wg), 1, 1), WorkgroupSize(wg, 1, 1)), B200MM_K_GEMM_WONNX);
}
std::pair<Workload, KernelSpec> bram(Context& c) {  // shaders/bram.wgsl:12-13: gid.x over M/4, gid.y over N/4
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(M / 4, 8), Workload::ceil(N / 4, 8), 1), WorkgroupSize(8, 8, 1)), B200MM_K_BRAM);
}
std::pair<Workload, KernelSpec> bram8x8(Context& c) {  // shaders/bram8x8.wgsl:10: fixed @workgroup_size(4,8,1)
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(M / 4, 4), Workload::ceil(N / 4, 8), 1), WorkgroupSize(4, 8, 1)), B200MM_K_BRAM8X8);
}
std::pair<Workload, KernelSpec> gemm3(Context& c) {  // shaders/gemm3.wgsl:15-16: x over N/8, y over M/4
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(N / 8, 16), Workload::ceil(M / 4, 16), 1), WorkgroupSize(16, 16, 1)), B200MM_K_GEMM3);
}

// ---- B200-native kernels: the Workload documents the library's launch shape and is advisory ----
std::pair<Workload, KernelSpec> sgemm_simt(Context& c) {
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
    c["BM"] = BM, c["BN"] = BN, c["BK"] = BK, c["TM"] = TM, c["TN"] = TN;
    auto r = finish(c, Workload(WorkgroupCount(Workload::ceil(M, BM), Workload::ceil(N, BN), 1), WorkgroupSize((BM * BN) / (TM * TN), 1, 1)),
                    B200MM_K_SGEMM_SIMT);
    return r;
}
std::pair<Workload, KernelSpec> sgemm_tc3x(Context& c) {
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t BM = 128, BN = 256, BK = 32;
    c["BM"] = BM, c["BN"] = BN, c["BK"] = BK;
    const size_t tiles = Workload::ceil(M, BM) * Workload::ceil(N, BN);
    // persistent: one CTA per SM (148 on B200), capped by the tile count
    return finish(c, Workload(WorkgroupCount((uint32_t)(tiles < 148 ? tiles : 148), 1, 1), WorkgroupSize(256, 1, 1)), B200MM_K_SGEMM_TC3X);
}

// single-pass TF32 (B200MM_F_TC3X_1X): misses the reference gate at large K by design -- it exists so that the harness's
// own "MAE too high" panic (src/harness.rs:82-84) can be exercised end to end
std::pair<Workload, KernelSpec> sgemm_tc3x_1x(Context& c) {
    auto r = sgemm_tc3x(c);
    r.second.name = "sgemm_tc3x_1x";
    r.second.params.flags |= B200MM_F_TC3X_1X;
    return r;
}

}  // namespace gemm

namespace gemv {

Dims insert_matrix_dims(Context& context, Dims d) { return put_dims(context, d, 1, 1024, 1024); }  // src/gemv.rs:5-15

std::pair<Workload, KernelSpec> qgemv_1(Context& c) {  // src/gemv.rs:17-33
    size_t M, N, K;
    need_dims(c, M, N, K);
    const size_t workgroup_size_x = 8;
    auto r = finish(c, Workload(WorkgroupCount(Workload::ceil(N, workgroup_size_x * 4), 1, 1), WorkgroupSize(workgroup_size_x, 1, 1)),
                    B200MM_K_QGEMV_1);
    r.second.params.absmax = ABSMAX;  // context.insert("absmax", &ABSMAX), src/gemv.rs:30
    r.second.params.batch = 1;
    return r;
}
std::pair<Workload, KernelSpec> qgemv_sint8(Context& c) {
    size_t M, N, K;
    need_dims(c, M, N, K);
    auto r = finish(c, Workload(WorkgroupCount(Workload::ceil(N, 512), 1, 1), WorkgroupSize(256, 1, 1)), B200MM_K_QGEMV_SINT8);
    r.second.params.absmax = ABSMAX;
    r.second.params.batch = 1;
    return r;
}
std::pair<Workload, KernelSpec> qgemv_sint8_grouped(Context& c, uint32_t group_k) {
    auto r = qgemv_sint8(c);
    r.second.name = "qgemv_sint8_grouped";
    r.second.params.absmax = 0.f;  // every (row block, column) carries its own scale behind the weights
    r.second.params.group_k = group_k;
    c["group_k"] = group_k;
    return r;
}
std::pair<Workload, KernelSpec> gemv_f32(Context& c) {
    size_t M, N, K;
    need_dims(c, M, N, K);
    return finish(c, Workload(WorkgroupCount(Workload::ceil(N, 128), 1, 1), WorkgroupSize(256, 1, 1)), B200MM_K_GEMV_F32);
}

}  // namespace gemv
}  // namespace wgpu_mm
