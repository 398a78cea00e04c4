// libb200mm.so -- implementation of include/b200mm.h: device bring-up, buffers, the kernel registry that
// replaces WGSL-module + pipeline creation, and the launch path that replaces `mm`
// (src/harness.rs:250-287 in the reference).  Everything here runs on a CUDA device; there is no
// host fallback of any kind.
#include "../../include/b200mm.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "kernels/datagen.cuh"
#include "kernels/gemv.cuh"
#include "kernels/sgemm_simt.cuh"
#include "kernels/sgemm_tc3x.cuh"
#include "kernels/wgsl_ports.cuh"
#include "kernels/tc_probe.cuh"
#include "kernels/peak_probe.cuh"

using namespace b200mm;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct b200mm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaDeviceProp prop{};
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    uint64_t launches = 0;
    std::string err;
    // copy streams + events of the pipelined host-buffer path (b200mm_mm_host), created on first use
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> pipe_ev;
};

struct b200mm_buffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = false;
    bool ipc = false;
    unsigned int barrier_epoch = 0;  // b200mm_peer_barrier call count when used as a flags buffer
};

struct b200mm_kernel {
    int id = 0;
    size_t M = 0, N = 0, K = 0;
    b200mm_kernel_params prm{};
    dim3 grid{1, 1, 1}, block{1, 1, 1};
    size_t smem = 0;
    // workspace
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // tc3x
    float *a_lo = nullptr, *b_lo = nullptr;  // lo parts of the operands (the raw operands are consumed as hi)
    float* b_hi = nullptr;                   // padded copy of B, only when N % 32 != 0
    CUtensorMap tmAh{}, tmAl{}, tmBh{}, tmBl{}, tmC{};
    bool tc_tma_store = false;       // pair kernel with the TMA-store epilogue
    const void* tc_c_src = nullptr;  // C the store map currently points at (with the peer set it was built for)
    int tc_bn = 256, tc_bk = 32;
    bool tc_a_prepass = false;       // the whole of A is split by the pre-pass (no row bands in the kernel)
    int tc_split = 0;                // Tc3xCfg::SPLIT: 1 = B_lo, 2 = A_lo and B_lo are computed inside the GEMM (2: no pre-pass at all)
    bool tc_cta2 = false;  // 2-CTA (cta_group::2) instantiation: 256 x 256 tiles on CTA pairs
    const void *tc_a_src = nullptr, *tc_b_src = nullptr;  // operands the hi tensor maps currently point at
    bool tc_b_copy = false;                               // ragged N: B is staged into a padded copy first
    float4* tc_partial = nullptr;
    unsigned int* tc_flags = nullptr;
    unsigned int* tc_band_cnt = nullptr;  // in-kernel A split: per-band completion counters (see Tc3xArgs)
    int tc_bands = 1, tc_prebands = 1;    // row bands of A; bands [0, prebands) are split by the pre-pass
    const void* tc_const_b = nullptr;     // B200MM_F_CONST_B: the B whose lo part currently sits in the workspace
    unsigned int tc_epoch = 0;
    int tc_cpt = 1, tc_full_waves = 0;
    long long tc_sk_units = 0;
    // gemv
    int splits = 1, rows_per_split = 0, panels = 0, gemv_variant = 0;
    bool gemv_blocked = false;  // grouped-scale GEMV with contiguous rows per warp (gemv.cuh BLOCKED)
    bool gemv_cluster = false;
    float* partial = nullptr;
    unsigned int* tickets = nullptr;
    // simt schedule: launch 1 = simt_tiles1 whole tiles, launch 2 = simt_tiles2 tiles x simt.split K-parts
    SimtSched simt{};
    int simt_tiles1 = 0, simt_tiles2 = 0;
    // multi-GPU
    PeerStore peers{};
    unsigned long long* trace_buf = nullptr;  // debug timeline of the GEMV kernels (b200mm_debug_gemv_trace)
    int trace_slots = 0, trace_next = 0;
    unsigned int peer_epoch = 0;   // launches since set_peer_flags (in-kernel cross-rank completion)
    size_t peer_pingpong = 0;      // GEMV: distance (floats) between the two y buffers alternated by epoch parity; 0 = one buffer
    // row-panel kernel object used by the pipelined host-buffer path (owned)
    b200mm_kernel* panel = nullptr;
    int panel_count = 0;
    bool tc_skip_b_split = false;  // B's lo part in the workspace is still valid (same B, pipelined panels)
    // skinny GEMM on the GEMV kernels for row counts that have no instantiation of their own (M = 3, 5, 6, 7, 9 .. 16): the
    // rows are cut into chunks that do (8 / 4 / 2 / 1), one child kernel object per chunk, launched back to back
    struct RowChunk {
        b200mm_kernel* kern;
        size_t row0;
    };
    std::vector<RowChunk> chunks;
    // sgemm_tc3x with N % 4 != 0 or K % 4 != 0: operands are staged into zero-padded copies (16-byte TMA strides) and the
    // padded result is copied back; `inner` is the kernel object of the padded shape (owned)
    b200mm_kernel* inner = nullptr;
    float *pad_a = nullptr, *pad_b = nullptr, *pad_c = nullptr;
    size_t pad_k = 0, pad_n = 0;
    // per-launch profiling of the dominant kernel
    bool profiling = false;
    std::vector<cudaEvent_t> pev;  // pairs
    int pcount = 0;
};

static thread_local std::string g_err;

static int fail(b200mm_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CU_TRY(ctx, expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(ctx, B200MM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                          \
    } while (0)

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// library / ctx
// ------------------------------------------------------------------------------------------------
extern "C" const char* b200mm_version(void) { return "b200mm 0.1.0 (sm_100a)"; }

extern "C" int b200mm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" const char* b200mm_last_error(const b200mm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int b200mm_ctx_create(int device_ordinal, b200mm_ctx** out) {
    if (!out) return fail(nullptr, B200MM_ERR_INVALID, "ctx_create: out is NULL");
    *out = nullptr;
    int n = b200mm_device_count();
    if (n <= 0 || device_ordinal < 0 || device_ordinal >= n)
        return fail(nullptr, B200MM_ERR_NO_DEVICE, "No GPU found given preference (device %d of %d)", device_ordinal, n);
    b200mm_ctx* c = new b200mm_ctx();
    c->device = device_ordinal;
    cudaError_t e = cudaSetDevice(device_ordinal);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&c->prop, device_ordinal);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) {
        int rc = fail(nullptr, B200MM_ERR_CUDA, "Could not create adapter for GPU device: %s", cudaGetErrorString(e));
        delete c;
        return rc;
    }
    *out = c;
    return B200MM_OK;
}

extern "C" int b200mm_ctx_destroy(b200mm_ctx* ctx) {
    if (!ctx) return B200MM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    for (auto e : ctx->pipe_ev) cudaEventDestroy(e);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B200MM_OK;
}

extern "C" int b200mm_ctx_device_info(const b200mm_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                                      size_t* global_mem_bytes, char* name, size_t name_len) {
    if (!ctx) return fail(nullptr, B200MM_ERR_INVALID, "device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (global_mem_bytes) *global_mem_bytes = ctx->prop.totalGlobalMem;
    if (name && name_len) {
        strncpy(name, ctx->prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return B200MM_OK;
}

extern "C" int b200mm_ctx_set_stream(b200mm_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, B200MM_ERR_INVALID, "set_stream: ctx is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->own_stream && ctx->stream) {
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CU_TRY(ctx, cudaStreamDestroy(ctx->stream));
    }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return B200MM_OK;
}

extern "C" void* b200mm_ctx_stream(const b200mm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int b200mm_sync(b200mm_ctx* ctx) {
    if (!ctx) return fail(nullptr, B200MM_ERR_INVALID, "sync: ctx is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200MM_OK;
}

extern "C" uint64_t b200mm_ctx_launch_count(const b200mm_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------
// buffers
// ------------------------------------------------------------------------------------------------
extern "C" int b200mm_buffer_create(b200mm_ctx* ctx, size_t bytes, b200mm_buffer** out) {
    if (!ctx || !out) return fail(ctx, B200MM_ERR_INVALID, "buffer_create: NULL argument");
    *out = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    b200mm_buffer* b = new b200mm_buffer();
    cudaError_t e = cudaMalloc(&b->ptr, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) {
        delete b;
        return fail(ctx, B200MM_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    b->bytes = bytes;
    b->owned = true;
    *out = b;
    return B200MM_OK;
}

extern "C" int b200mm_buffer_create_init(b200mm_ctx* ctx, const void* host, size_t bytes, b200mm_buffer** out) {
    if (!host) return fail(ctx, B200MM_ERR_INVALID, "buffer_create_init: host is NULL");
    int rc = b200mm_buffer_create(ctx, bytes, out);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync((*out)->ptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // contents are visible when the call returns
    if (e != cudaSuccess) {
        b200mm_buffer_free(ctx, *out);
        *out = nullptr;
        return fail(ctx, B200MM_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    }
    return B200MM_OK;
}

extern "C" int b200mm_buffer_wrap(b200mm_ctx* ctx, void* device_ptr, size_t bytes, b200mm_buffer** out) {
    if (!ctx || !out || !device_ptr) return fail(ctx, B200MM_ERR_INVALID, "buffer_wrap: NULL argument");
    b200mm_buffer* b = new b200mm_buffer();
    b->ptr = device_ptr;
    b->bytes = bytes;
    b->owned = false;
    *out = b;
    return B200MM_OK;
}

extern "C" int b200mm_buffer_free(b200mm_ctx* ctx, b200mm_buffer* buf) {
    if (!buf) return B200MM_OK;
    if (ctx) cudaSetDevice(ctx->device);
    if (buf->ipc)
        cudaIpcCloseMemHandle(buf->ptr);
    else if (buf->owned && buf->ptr)
        cudaFree(buf->ptr);
    delete buf;
    return B200MM_OK;
}

extern "C" void* b200mm_buffer_device_ptr(const b200mm_buffer* buf) { return buf ? buf->ptr : nullptr; }
extern "C" size_t b200mm_buffer_bytes(const b200mm_buffer* buf) { return buf ? buf->bytes : 0; }

extern "C" int b200mm_buffer_write(b200mm_ctx* ctx, b200mm_buffer* buf, size_t offset, const void* host, size_t bytes) {
    if (!ctx || !buf || !host) return fail(ctx, B200MM_ERR_INVALID, "buffer_write: NULL argument");
    if (offset + bytes > buf->bytes) return fail(ctx, B200MM_ERR_INVALID, "buffer_write: range exceeds buffer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync((char*)buf->ptr + offset, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B200MM_OK;
}

extern "C" int b200mm_buffer_read(b200mm_ctx* ctx, const b200mm_buffer* buf, size_t offset, void* host, size_t bytes) {
    if (!ctx || !buf || !host) return fail(ctx, B200MM_ERR_INVALID, "buffer_read: NULL argument");
    if (offset + bytes > buf->bytes) return fail(ctx, B200MM_ERR_INVALID, "Error reading buffer: range exceeds buffer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(host, (const char*)buf->ptr + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200MM_OK;
}

extern "C" int b200mm_buffer_read_2d(b200mm_ctx* ctx, const b200mm_buffer* buf, size_t offset, size_t src_pitch, void* host, size_t dst_pitch,
                                     size_t width_bytes, size_t rows) {
    if (!ctx || !buf || !host) return fail(ctx, B200MM_ERR_INVALID, "buffer_read_2d: NULL argument");
    if (width_bytes > src_pitch || width_bytes > dst_pitch) return fail(ctx, B200MM_ERR_INVALID, "buffer_read_2d: width exceeds a pitch");
    if (rows && offset + (rows - 1) * src_pitch + width_bytes > buf->bytes) return fail(ctx, B200MM_ERR_INVALID, "Error reading buffer: range exceeds buffer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpy2DAsync(host, dst_pitch, (const char*)buf->ptr + offset, src_pitch, width_bytes, rows, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200MM_OK;
}

extern "C" int b200mm_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(nullptr, B200MM_ERR_INVALID, "host_alloc: out is NULL");
    CU_TRY(nullptr, cudaMallocHost(out, std::max<size_t>(bytes, 16)));
    return B200MM_OK;
}
extern "C" int b200mm_host_free(void* p) {
    if (p) CU_TRY(nullptr, cudaFreeHost(p));
    return B200MM_OK;
}

extern "C" int b200mm_buffer_fill_weights(b200mm_ctx* ctx, b200mm_buffer* buf, uint64_t seed, uint64_t offset, size_t n) {
    if (!ctx || !buf) return fail(ctx, B200MM_ERR_INVALID, "fill_weights: NULL argument");
    if (n * sizeof(float) > buf->bytes) return fail(ctx, B200MM_ERR_INVALID, "fill_weights: range exceeds buffer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int blocks = (int)std::min<size_t>(ceil_div(n, 256), (size_t)ctx->prop.multiProcessorCount * 16);
    fill_weights_kernel<<<std::max(blocks, 1), 256, 0, ctx->stream>>>((float*)buf->ptr, seed, offset, n);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200MM_OK;
}

extern "C" int b200mm_buffer_fill_weights_2d(b200mm_ctx* ctx, b200mm_buffer* buf, uint64_t seed, uint64_t offset, size_t rows,
                                             size_t cols, size_t src_ld, size_t src_col0) {
    if (!ctx || !buf) return fail(ctx, B200MM_ERR_INVALID, "fill_weights_2d: NULL argument");
    if (rows * cols * sizeof(float) > buf->bytes) return fail(ctx, B200MM_ERR_INVALID, "fill_weights_2d: range exceeds buffer");
    if (src_col0 + cols > src_ld) return fail(ctx, B200MM_ERR_INVALID, "fill_weights_2d: panel exceeds the source row");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int blocks = (int)std::min<size_t>(ceil_div(rows * cols, 256), (size_t)ctx->prop.multiProcessorCount * 16);
    fill_weights_2d_kernel<<<std::max(blocks, 1), 256, 0, ctx->stream>>>((float*)buf->ptr, seed, offset, rows, cols, src_ld, src_col0);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200MM_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// A-like operand: row-major rows x cols f32, box = box_rows x 32 floats (128 B, SWIZZLE_128B), K-major.
static int make_tmap_kmajor(b200mm_ctx* ctx, CUtensorMap* tm, const float* base, size_t rows, size_t cols, int box_rows,
                            int box_cols = 32) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(ctx, B200MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, B200MM_ERR_CUDA, "cuTensorMapEncodeTiled(K-major) failed: %d", (int)r);
    return B200MM_OK;
}

// B operand: row-major K x N f32 consumed MN-major.  Viewed as 3-D (n%32, k, n/32) so that one TMA
// box lands [n/32][k][32] = the canonical SWIZZLE_128B MN-major atoms in shared memory.
static int make_tmap_mnmajor(b200mm_ctx* ctx, CUtensorMap* tm, const float* base, size_t K, size_t N, int box_k,
                             int box_n, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(ctx, B200MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {32, K, ceil_div(N, 32)};
    cuuint64_t strides[2] = {N * sizeof(float), 32 * sizeof(float)};
    cuuint32_t box[3] = {32, (cuuint32_t)box_k, (cuuint32_t)(box_n / 32)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, B200MM_ERR_CUDA, "cuTensorMapEncodeTiled(MN-major) failed: %d", (int)r);
    return B200MM_OK;
}

// C (store side): rows x cols f32 with leading dimension ld, 32 x 32 boxes (128-byte rows), SWIZZLE_128B.
static int make_tmap_c(b200mm_ctx* ctx, CUtensorMap* tm, float* base, size_t rows, size_t cols, size_t ld) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(ctx, B200MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, B200MM_ERR_CUDA, "cuTensorMapEncodeTiled(C) failed: %d", (int)r);
    return B200MM_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel registry
// ------------------------------------------------------------------------------------------------
extern "C" const char* b200mm_kernel_name(int id) {
    switch (id) {
        case B200MM_K_GEMM_1: return "gemm_1";
        case B200MM_K_GEMM_1V: return "gemm_1v";
        case B200MM_K_GEMM_2: return "gemm_2";
        case B200MM_K_GEMM_3: return "gemm_3";
        case B200MM_K_GEMM_4: return "gemm_4";
        case B200MM_K_GEMM_5: return "gemm_5";
        case B200MM_K_GEMM_WONNX: return "gemm_wonnx";
        case B200MM_K_BRAM: return "bram";
        case B200MM_K_BRAM8X8: return "bram8x8";
        case B200MM_K_GEMM3: return "gemm3";
        case B200MM_K_QGEMV_1: return "qgemv_1";
        case B200MM_K_SGEMM_SIMT: return "sgemm_simt";
        case B200MM_K_SGEMM_TC3X: return "sgemm_tc3x";
        case B200MM_K_GEMV_F32: return "gemv_f32";
        case B200MM_K_QGEMV_SINT8: return "qgemv_sint8";
        default: return "unknown";
    }
}

using Tc256 = Tc3xCfg<256, 2, false, 32>;     // 2 stages x 96 KB
using Tc256k16 = Tc3xCfg<256, 4, false, 16>;  // 4 stages x 48 KB: same bytes in flight, finer refill granularity
using Tc256k16x2 = Tc3xCfg<256, 6, false, 16, 256, true>;  // 2-CTA pairs: 256 x 256 tiles, 6 stages x 32 KB per CTA
using Tc256k16x2s = Tc3xCfg<256, 5, false, 16, 256, true, true>;  // same with the TMA-store epilogue: 5 stages + 64 KB of staging (the default pair kernel; tune[2] = 6 selects the one above)
using Tc256k16x2sb = Tc3xCfg<256, 5, false, 16, 256, true, true, 1>;  // ... and B_lo computed in shared memory (SPLIT = 1)
using Tc256k16b = Tc3xCfg<256, 4, false, 16, 256, false, false, 1>;   // 1-CTA kernel with SPLIT = 1
using Tc256k16x2sab = Tc3xCfg<256, 5, false, 16, 256, true, true, 2>;  // A_lo and B_lo computed in shared memory (SPLIT = 2): no pre-pass
using Tc256k16ab = Tc3xCfg<256, 4, false, 16, 256, false, false, 2>;
using Tc256k32x2 = Tc3xCfg<256, 3, false, 32, 256, true>;  // same with BK = 32: 3 stages x 64 KB (tune[2] = 32; measured, not the default)
using Tc128 = Tc3xCfg<128, 3, false, 32>;
using Tc128b = Tc3xCfg<128, 3, false, 32, 256, false, false, 1>;  // 128 x 128 tiles with B_lo computed in shared memory
using Tc256x1 = Tc3xCfg<256, 4, true, 32>;

// Streaming-GEMV instantiations (template arguments: warps, unroll, lanes per row segment, rows of x, grouped scales,
// min CTAs/SM, eager double prefetch).  Variant ids are what params.tune[0] / the autotuner select; measured at cfg3 / cfg4
// with tools/sweep_gemv.py and tools/bench_grouped.py (profiles/r1_qgemv_variant_split_sweep.log).
//   0 (tune 100)  8 warps, unroll 8, 32 lanes/row        4  8 warps, unroll 4, 16 lanes/row (sint8 default until the eager form)
//   1             8 warps, unroll 8, 16 lanes/row        5  4 warps, unroll 4, 32 lanes/row (fp32 default)
//   2             4 warps, unroll 8, 32 lanes/row        6  4 warps, unroll 8, 16 lanes/row
//   3             8 warps, unroll 4, 32 lanes/row        7  4 warps, unroll 4, 16 lanes/row
//   sint8 only: 11 = variant 4 capped at 80 registers (3 CTAs/SM), 12 = at 64 registers (4 CTAs/SM),
//               13 = variant 4 with both register buffers in flight before the PDL wait (128 registers, 2 CTAs/SM; the default)
//   sint8 with per-group scales (8 warps, unroll 4, 16 lanes: window of 128 rows): default = eager, 2 CTAs/SM; 11 = 3 CTAs/SM;
//               12 = 2 CTAs/SM without the eager prefetch; 14 = 4 CTAs/SM (spills)
using GemvFn = void (*)(const float*, const void*, float*, float*, unsigned int*, int, int, int, float, size_t, size_t, size_t, PeerStore, int,
                        int);
struct GemvPick {
    GemvFn fn;
    int warps, lpr;
};
#define GEMV_INST(W, U, L, ...) \
    GemvPick { gemv_stream_kernel<T, W, U, L, ##__VA_ARGS__>, W, L }

// grouped: 0 = no per-group scales, else the quantisation group size (rows).  Groups of 64 / 32 rows need a pipeline window
// (2 x UNROLL x 16 rows) that never straddles a group: UNROLL 2 / 1 instead of 4.
template <class T>
static GemvPick gemv_pick(int variant, int mrows = 1, size_t grouped = 0, bool blocked = false) {
    if constexpr (T::COLS == 16) {
        // (with the blocked row order a thread's 16-row window never straddles a multiple of 32, so the UNROLL 4 kernel serves these too)
        if (grouped && grouped < 128 && !blocked) return grouped >= 64 ? GEMV_INST(8, 2, 16, 1, true, 2, 1) : GEMV_INST(8, 1, 16, 1, true, 2, 1);
        if (grouped) {
            switch (variant) {
                case 11: return GEMV_INST(8, 4, 16, 1, true, 3);
                case 12: return GEMV_INST(8, 4, 16, 1, true, 2);
                case 14: return GEMV_INST(8, 4, 16, 1, true, 4);
                default: return blocked ? GEMV_INST(8, 4, 16, 1, true, 2, 1, true) : GEMV_INST(8, 4, 16, 1, true, 2, 1);
            }
        }
        if (mrows == 2) return GEMV_INST(8, 4, 16, 2);
        if (mrows > 2) return GEMV_INST(8, 4, 16, 4);
        if (variant == 11) return GEMV_INST(8, 4, 16, 1, false, 3);
        if (variant == 12) return GEMV_INST(8, 4, 16, 1, false, 4);
        if (variant == 13) return GEMV_INST(8, 4, 16, 1, false, 1, 1);
        if (variant == 21) return GEMV_INST(8, 4, 8, 1, false, 1, 1);  // 128-column panels: the balanced (ragged-panel) grids, tune[3] = panel count
        if (variant == 25) return GEMV_INST(8, 4, 4, 1, false, 1, 1);  // 64-column panels: small matrices (per-rank panels of an N-sharded run)
    } else {
        // skinny GEMM: only the default geometry of each weight type is instantiated for M = 2, 4, 8
        if (mrows == 2) return GEMV_INST(4, 4, 32, 2);
        if (mrows == 4) return GEMV_INST(4, 4, 32, 4);
        if (mrows > 4) return GEMV_INST(4, 4, 32, 8);
    }
    switch (variant) {
        case 1: return GEMV_INST(8, 8, 16);
        case 2: return GEMV_INST(4, 8, 32);
        case 3: return GEMV_INST(8, 4, 32);
        case 4: return GEMV_INST(8, 4, 16);
        case 5: return GEMV_INST(4, 4, 32);
        case 6: return GEMV_INST(4, 8, 16);
        case 7: return GEMV_INST(4, 4, 16);
        default: return GEMV_INST(8, 8, 32);
    }
}

static int setup_ports(b200mm_ctx* ctx, b200mm_kernel* k) {
    const size_t M = k->M, N = k->N, K = k->K;
    auto ws = k->prm.workgroup_size;
    auto blk = [&](unsigned x, unsigned y, unsigned z) {
        k->block = dim3(ws[0] ? ws[0] : x, ws[1] ? ws[1] : y, ws[2] ? ws[2] : z);
    };
    auto need = [&](bool ok, const char* what) {
        return ok ? 0 : fail(ctx, B200MM_ERR_INVALID, "%s: shape %zux%zux%zu violates %s", b200mm_kernel_name(k->id), M, N, K, what);
    };
    int rc = 0;
    switch (k->id) {
        case B200MM_K_GEMM_1:  // src/gemm.rs:20-26
            blk(16, 16, 1);
            k->grid = dim3(ceil_div(M, k->block.x), ceil_div(N, k->block.y), 1);
            break;
        case B200MM_K_GEMM_1V:  // src/gemm.rs:37-44
            if ((rc = need(N % 4 == 0 && K % 4 == 0, "N%4==0 && K%4==0"))) return rc;
            blk(16, 4, 1);
            k->grid = dim3(ceil_div(M, k->block.x), ceil_div(N / 4, k->block.y), 1);
            break;
        case B200MM_K_GEMM_2:  // src/gemm.rs:55-62
            blk(256, 1, 1);
            if (k->block.x != 256) return fail(ctx, B200MM_ERR_INVALID, "gemm_2 needs workgroup_size_x == 256");
            k->grid = dim3(ceil_div(M, 16), ceil_div(N, 16), 1);
            break;
        case B200MM_K_GEMM_3:  // src/gemm.rs:72-84
            if ((rc = need(M % 16 == 0 && N % 16 == 0 && K % 16 == 0, "M,N,K%16==0 (shader has no guards)"))) return rc;
            blk(256, 1, 1);
            if (k->block.x != 256) return fail(ctx, B200MM_ERR_INVALID, "gemm_3 needs workgroup_size_x == 256");
            k->grid = dim3(M / 16, N / 16, 1);
            break;
        case B200MM_K_GEMM_4:  // src/gemm.rs:95-111
            if ((rc = need(M % 16 == 0 && N % 16 == 0 && K % 8 == 0, "M,N%16==0, K%8==0"))) return rc;
            blk(128, 1, 1);
            if (k->block.x != 128) return fail(ctx, B200MM_ERR_INVALID, "gemm_4 needs workgroup_size_x == 128");
            k->grid = dim3(N / 16, M / 16, 1);
            break;
        case B200MM_K_GEMM_5:  // src/gemm.rs:124-142
            if ((rc = need(M % 32 == 0 && N % 32 == 0 && K % 16 == 0, "M,N%32==0, K%16==0"))) return rc;
            blk(64, 1, 1);
            if (k->block.x != 64) return fail(ctx, B200MM_ERR_INVALID, "gemm_5 needs workgroup_size_x == 64");
            k->grid = dim3(N / 32, M / 32, 1);
            break;
        case B200MM_K_GEMM_WONNX:  // orphan: 1-D dispatch of M*N/16 invocations
            if ((rc = need(M % 4 == 0 && N % 4 == 0 && K % 4 == 0, "M,N,K%4==0"))) return rc;
            blk(256, 1, 1);
            k->grid = dim3(ceil_div(M * N / 16, k->block.x), 1, 1);
            break;
        case B200MM_K_BRAM:  // orphan: gid.x over M/4, gid.y over N/4
            if ((rc = need(M % 4 == 0 && N % 4 == 0 && K % 4 == 0, "M,N,K%4==0"))) return rc;
            blk(8, 8, 1);
            k->grid = dim3(ceil_div(M / 4, k->block.x), ceil_div(N / 4, k->block.y), 1);
            break;
        case B200MM_K_BRAM8X8:  // fixed @workgroup_size(4,8,1), shaders/bram8x8.wgsl:10
            if ((rc = need(M % 4 == 0 && N % 4 == 0 && K % 4 == 0, "M,N,K%4==0"))) return rc;
            k->block = dim3(4, 8, 1);
            k->grid = dim3(ceil_div(M / 4, 4), ceil_div(N / 4, 8), 1);
            break;
        case B200MM_K_GEMM3:  // orphan: gid.x over N/8, gid.y over M/4
            if ((rc = need(M % 4 == 0 && N % 8 == 0 && K % 4 == 0, "M%4==0, N%8==0, K%4==0"))) return rc;
            blk(16, 16, 1);
            k->grid = dim3(ceil_div(N / 8, k->block.x), ceil_div(M / 4, k->block.y), 1);
            break;
        case B200MM_K_QGEMV_1: {  // src/gemv.rs:20-26
            if ((rc = need(M == 1 && N % 4 == 0 && K % 4 == 0, "M==1, N%4==0, K%4==0"))) return rc;
            blk(8, 1, 1);
            const unsigned batch = k->prm.batch ? k->prm.batch : 1;
            k->grid = dim3(ceil_div(N, (size_t)k->block.x * 4), ceil_div(batch, k->block.y), 1);
            break;
        }
        default:
            return fail(ctx, B200MM_ERR_UNSUPPORTED, "unknown port id %d", k->id);
    }
    if ((size_t)k->block.x * k->block.y * k->block.z > 1024 || k->block.x > 1024 || k->block.y > 1024 || k->block.z > 64)
        return fail(ctx, B200MM_ERR_LIMITS, "Compute limits exceeded");
    return B200MM_OK;
}

static int setup_simt(b200mm_ctx* ctx, b200mm_kernel* k) {
    if (k->M > INT32_MAX || k->N > INT32_MAX || k->K > INT32_MAX) return fail(ctx, B200MM_ERR_INVALID, "shape too large");
    k->block = dim3(SimtCfg::THREADS, 1, 1);
    const long long tiles_m = (long long)ceil_div(k->M, SimtCfg::BM), tiles_n = (long long)ceil_div(k->N, SimtCfg::BN);
    const long long tiles = tiles_m * tiles_n;
    if (tiles > INT32_MAX) return fail(ctx, B200MM_ERR_LIMITS, "Compute limits exceeded");
    // 2 resident CTAs per SM (registers); see SimtSched for the two-launch schedule
    const long long slots = (long long)ctx->prop.multiProcessorCount * 2;
    SimtSched& sc = k->simt;
    sc.tiles_m = (int)tiles_m;
    sc.tiles_n = (int)tiles_n;
    sc.group_m = k->prm.tune[1] ? (int)k->prm.tune[1] : 16;
    sc.split = 1;
    const int KT = (int)ceil_div(k->K, SimtCfg::BK);
    const bool seq_k = (k->prm.flags & B200MM_F_SEQUENTIAL_K) != 0;
    // Measured on B200 (tools/bench_simt.py): once every SM has a tile, splitting the tiles of a partial last wave
    // buys nothing (one CTA per SM sustains the same FMA rate as two), so K is split only when there are fewer
    // tiles than resident CTAs (e.g. 1024^3 = 64 tiles -> 4 K-parts each).
    long long rem = tiles < slots ? tiles : 0;
    int split = 1;
    if (!seq_k && rem > 0) split = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(slots / rem, 8), KT / 4));
    if (split <= 1) rem = 0;  // nothing to gain: one launch over all tiles
    k->simt_tiles1 = (int)(tiles - rem);
    k->simt_tiles2 = (int)rem;
    sc.split = split;
    k->grid = dim3((unsigned)std::max<long long>(k->simt_tiles1, rem * split), 1, 1);
    if (rem > 0) {
        const size_t n_cta2 = (size_t)rem * split;
        const size_t part_bytes = n_cta2 * SimtCfg::BM * SimtCfg::BN * sizeof(float);
        const size_t flag_bytes = ceil_div(n_cta2 * sizeof(unsigned int), 256) * 256;
        k->ws_bytes = part_bytes + flag_bytes;
        CU_TRY(ctx, cudaMalloc(&k->ws, k->ws_bytes));
        CU_TRY(ctx, cudaMemsetAsync(k->ws, 0, k->ws_bytes, ctx->stream));
        sc.partial = (float4*)k->ws;
        sc.flags = (unsigned int*)((char*)k->ws + part_bytes);
    }
    return B200MM_OK;
}

// What the default rules pick for a shape (pure: no device, no allocation) -- used by setup_tc3x and exported as b200mm_tc3x_plan so
// that the rules are pinned by CPU tests.
struct Tc3xPlan {
    int bn = 256, bk = 16, split = 0;
    bool cta2 = false, tma_store = false, a_prepass = false;
};
static Tc3xPlan tc3x_plan(size_t M, size_t N, size_t K, int sms, const uint32_t tune[4], bool one_pass) {
    Tc3xPlan p;
    p.bn = (tune[0] == 128) ? 128 : 256;
    // Small problems (tune[0] = 0): 128 x 128 tiles when even those leave no SM without a tile (twice the tiles, so half the k-slices
    // per tile to reach one CTA per SM and half the fix-up traffic) and B is small.  Measured (tools/small_bn.py,
    // profiles/r2_small_bn.log): 1024^3 31.7 -> 24.3 us, 512^3 22.6 -> 16.2 us, 4096 x 512 x 512 26.9 -> 18.5 us; from 2048^3 on, and
    // for skinny M against a big B (128 x 14336 x 4096: 90 vs 167 us), the 256-column tiles win.
    // Against a big B the narrow tiles still pay when there are so few 256-column tiles that each would be cut into >= 8 k-slices
    // (<= SMs / 8 tiles: skinny M with N <= 4096) -- together with B_lo computed in shared memory, see tc_split below:
    // 128 x 4096 x 4096 54.7 -> 39.2 us, 16 x 4096 x 4096 51.3 -> 37.6 us.
    {
        const size_t tiles256 = ceil_div(M, 128) * ceil_div(N, 256);
        if (tune[0] == 0 && !one_pass && tiles256 * 2 <= (size_t)sms && (K * N <= ((size_t)8 << 20) || tiles256 * 8 <= (size_t)sms)) p.bn = 128;
    }
    p.bk = (tune[2] == 32) ? 32 : 16;  // default: BK = 16, 4 stages
    if (one_pass) p.bn = 256;
    if (one_pass || p.bn == 128) p.bk = 32;  // (128 x 128 tiles with BK = 16, 6 stages: measured 8-20 % slower than BK = 32)
    // tune[0] = 512: the 2-CTA kernel (256 x 256 tiles on CTA pairs, cta_group::2); 513: force the 1-CTA kernel.  Default: pairs when
    // there are at least as many 256-row tiles as SM pairs (big GEMMs), single CTAs otherwise (skinny M: a pair would idle one SM).
    {
        const long long tiles2 = (long long)ceil_div(M, 256) * (long long)ceil_div(N, 256);
        // (a tile count that would leave SMs idle runs as stream-K over all of them -- tc3x_make_schedule -- so pairs pay from about
        // 2/3 of a wave on: 1792^3 63.3 -> 60.9 us, 2048^3 79.5 -> 76.1 us, 768 x 4096 x 4096 136 -> 129 us; not for a single row of
        // pair tiles, M <= 256, where the skinny-M rules below decide)
        const bool fits_l2 = 8.0 * ((double)M * (double)K + (double)K * (double)N) <= 100e6;
        const long long min_tiles2 = (fits_l2 || M >= 512) ? 48 : sms / 2;
        // 256-row tiles must not pad M much more than 128-row tiles would (640 rows: 768 vs 640 computed -- measured 126 vs 120 us)
        const bool pad_ok = ceil_div(M, 256) * 256 == ceil_div(M, 128) * 128 || M >= 2048;
        const bool want2 = tune[0] == 512 || (tune[0] == 0 && tiles2 >= min_tiles2 && pad_ok && getenv("B200MM_TC3X_1CTA") == nullptr);
        p.cta2 = want2 && !one_pass && p.bn == 256 && sms % 2 == 0;
        if (p.cta2) p.bk = (tune[2] == 32) ? 32 : 16;
        // pair kernel: TMA-store epilogue (5 stages + double-buffered staging) unless tune[2] = 6 asks for the st.global one (6 stages)
        p.tma_store = p.cta2 && p.bk == 16 && tune[2] != 6;
        // lo tiles computed in shared memory (Tc3xCfg::SPLIT), available in the two default instantiations.  Measured
        // (profiles/r2_split_shapes.log): it pays where the GEMM is bound by reading B -- skinny M: 128 x 14336 x 4096 165 -> 101 us --
        // and costs where the tensor pipe is the bound, because the split's LDS / STS compete with the MMA's operand reads for
        // shared-memory bandwidth (4096^3: 529 -> 554 us with B only, 607 us with A and B; 8192^3: +15 % / +64 %).
        // tune[3]: 0 = that rule (M <= 256 and a B of >= 16 MB: B in the kernel, A -- small -- in the pre-pass; otherwise as 2),
        // 1 = B in the kernel, A in the pre-pass, 2 = B in the pre-pass, A by row bands (pre-pass for the first wave, warp 2 for the
        // rest), 3 = A and B in the pre-pass (round 1), 4 = B in the kernel, A by row bands, 5 = A and B in the kernel (no pre-pass)
        const bool can_split = !one_pass && ((p.bn == 256 && p.bk == 16 && (p.tma_store || !p.cta2)) || p.bn == 128);
        uint32_t t3 = tune[3];
        if (t3 == 0) t3 = (can_split && M <= 256 && K * N >= ((size_t)4 << 20)) ? 1 : 2;
        p.split = !can_split ? 0 : (t3 == 1 || t3 == 4) ? 1 : (t3 == 5 ? (p.bn == 128 ? 1 : 2) : 0);
        p.a_prepass = (t3 == 1 || t3 == 3);
    }
    return p;
}

static int setup_tc3x(b200mm_ctx* ctx, b200mm_kernel* k) {
    const size_t M = k->M, N = k->N, K = k->K;
    if (ctx->prop.major != 10)
        return fail(ctx, B200MM_ERR_UNSUPPORTED, "sgemm_tc3x needs an sm_100 device (got sm_%d%d)", ctx->prop.major, ctx->prop.minor);
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return fail(ctx, B200MM_ERR_INVALID, "shape too large");
    if (N % 4 || K % 4) {
        // TMA needs 16-byte row strides and the epilogue stores float4: run the padded shape (K -> K4, N -> N4, zero fill) on
        // staged copies.  An edge path: three extra HBM passes over the operands, O(MK + KN + MN).
        if (k->prm.flags & B200MM_F_PEER_STORE) return fail(ctx, B200MM_ERR_INVALID, "sgemm_tc3x: peer stores need N%%4==0 and K%%4==0");
        k->pad_k = ceil_div(K, 4) * 4;
        k->pad_n = ceil_div(N, 4) * 4;
        b200mm_kernel_params prm = k->prm;
        int rc = b200mm_kernel_get(ctx, B200MM_K_SGEMM_TC3X, M, k->pad_n, k->pad_k, &prm, &k->inner);
        if (rc) return rc;
        const size_t ab = M * k->pad_k * 4, bb = k->pad_k * k->pad_n * 4, cb = M * k->pad_n * 4;
        const size_t al = ceil_div(ab, 256) * 256, bl = ceil_div(bb, 256) * 256;
        k->ws_bytes = al + bl + cb;
        CU_TRY(ctx, cudaMalloc(&k->ws, k->ws_bytes));
        CU_TRY(ctx, cudaMemsetAsync(k->ws, 0, k->ws_bytes, ctx->stream));  // the pad columns / rows stay zero for ever
        k->pad_a = (float*)k->ws;
        k->pad_b = (float*)((char*)k->ws + al);
        k->pad_c = (float*)((char*)k->ws + al + bl);
        k->grid = k->inner->grid;
        k->block = k->inner->block;
        return B200MM_OK;
    }
    const bool one_pass = (k->prm.flags & B200MM_F_TC3X_1X) != 0;
    {
        const Tc3xPlan p = tc3x_plan(M, N, K, ctx->prop.multiProcessorCount, k->prm.tune, one_pass);
        k->tc_bn = p.bn;
        k->tc_bk = p.bk;
        k->tc_cta2 = p.cta2;
        k->tc_tma_store = p.tma_store;
        k->tc_split = p.split;
        k->tc_a_prepass = p.a_prepass;
    }
    const int tile_m = k->tc_cta2 ? 256 : 128;
    // workspace: lo parts of both operands (the raw operands are consumed as hi); for N % 32 != 0 also a padded
    // copy of B, because the 3-D view (n%32, k, n/32) of a ragged N reads up to 124 B past the last row
    const size_t a_bytes = M * K * sizeof(float), b_bytes = K * N * sizeof(float) + 128;
    const size_t a_al = ceil_div(a_bytes, 1024) * 1024, b_al = ceil_div(b_bytes, 1024) * 1024;
    k->tc_b_copy = (N % 32) != 0;
    // schedule (see Tc3xArgs / tc3x_make_schedule): units = tiles x chains, one contiguous range per CTA
    const int bk = k->tc_bk;
    const int sms = ctx->prop.multiProcessorCount;
    // scheduling units: SMs, or SM pairs for the 2-CTA kernel
    const Tc3xSchedule sched = tc3x_make_schedule(M, N, K, k->tc_bn, bk, k->tc_cta2 ? sms / 2 : sms, k->prm.tune[1] == 1, tile_m);  // tune[1] = 1: pure stream-K (experiments)
    const int grid_x = sched.grid * (k->tc_cta2 ? 2 : 1);
    k->tc_cpt = sched.chains_per_tile;
    k->tc_full_waves = sched.full_waves;
    k->tc_sk_units = sched.sk_units;
    const size_t part_bytes = (size_t)grid_x * 128 * k->tc_bn * sizeof(float);
    // In-kernel A split (Tc3xArgs): the pre-pass covers the row bands the first wave of tiles touches, warp 2 of every CTA
    // does the rest while earlier bands are multiplied (unless tune[3] asks for the whole of A in the pre-pass, see above).
    k->tc_bands = (int)ceil_div(M, (size_t)kTc3xBandRows);
    {
        const long long band_tiles = (long long)(kTc3xBandRows / tile_m) * (long long)ceil_div(N, (size_t)k->tc_bn);
        const long long first_wave = std::min<long long>(sched.grid, sched.tiles);
        long long pre = sched.full_waves == 0 ? k->tc_bands : (first_wave + band_tiles - 1) / band_tiles;
        if (k->tc_a_prepass || one_pass || k->tc_split == 2) pre = k->tc_bands;
        k->tc_prebands = (int)std::min<long long>(std::max<long long>(pre, 1), k->tc_bands);
    }
    const size_t flag_bytes = ceil_div(((size_t)grid_x + (size_t)k->tc_bands) * sizeof(unsigned int), 1024) * 1024;
    // lo parts that are computed in shared memory (tc_split) need no workspace
    const bool need_a_lo = !one_pass && k->tc_split < 2, need_b_lo = !one_pass && k->tc_split < 1;
    k->ws_bytes = (need_a_lo ? a_al : 0) + (need_b_lo ? b_al : 0) + (k->tc_b_copy ? b_al : 0) + part_bytes + flag_bytes;
    CU_TRY(ctx, cudaMalloc(&k->ws, k->ws_bytes));
    char* w = (char*)k->ws;
    if (need_a_lo) {
        k->a_lo = (float*)w;
        w += a_al;
    }
    if (need_b_lo) {
        k->b_lo = (float*)w;
        w += b_al;
    }
    if (k->tc_b_copy) {
        k->b_hi = (float*)w;
        w += b_al;
    }
    k->tc_partial = (float4*)w;
    w += part_bytes;
    k->tc_flags = (unsigned int*)w;
    k->tc_band_cnt = k->tc_flags + grid_x;
    CU_TRY(ctx, cudaMemsetAsync(k->ws, 0, k->ws_bytes, ctx->stream));
    int rc;
    if (need_a_lo && (rc = make_tmap_kmajor(ctx, &k->tmAl, k->a_lo, M, K, 128, bk))) return rc;
    if (need_b_lo && (rc = make_tmap_mnmajor(ctx, &k->tmBl, k->b_lo, K, N, bk, k->tc_cta2 ? k->tc_bn / 2 : k->tc_bn))) return rc;
    // the hi maps point at the caller's A and B and are (re)encoded at launch time
    k->grid = dim3(grid_x, 1, 1);
    k->block = dim3(Tc256::THREADS, 1, 1);
    if (k->tc_tma_store && k->tc_split) {
        k->smem = Tc256k16x2sb::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16x2sb>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16x2sab>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_split && k->tc_bn == 128) {
        k->smem = Tc128b::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc128b>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_split) {
        k->smem = Tc256k16b::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16b>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16ab>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_tma_store) {
        k->smem = Tc256k16x2s::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16x2s>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_cta2 && k->tc_bk == 32) {
        k->smem = Tc256k32x2::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k32x2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_cta2) {
        k->smem = Tc256k16x2::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16x2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (one_pass) {
        k->smem = Tc256x1::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256x1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_bn == 256 && k->tc_bk == 16) {
        k->smem = Tc256k16::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256k16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else if (k->tc_bn == 256) {
        k->smem = Tc256::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    } else {
        k->smem = Tc128::SMEM_BYTES;
        CU_TRY(ctx, cudaFuncSetAttribute(sgemm_tc3x_kernel<Tc128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem));
    }
    return B200MM_OK;
}

static int setup_gemv(b200mm_ctx* ctx, b200mm_kernel* k, bool quant) {
    const size_t M = k->M, N = k->N, K = k->K;
    const int cols = quant ? 16 : 4;
    if (M > 16)
        return fail(ctx, B200MM_ERR_INVALID, "%s takes M <= 16 rows of x (skinny GEMM); use an SGEMM kernel for larger M", b200mm_kernel_name(k->id));
    if ((M != 1 && M != 2 && M != 4 && !(M == 8 && !quant)) || (M > 1 && k->prm.group_k)) {
        // no instantiation for this row count: chunks of 8 / 4 / 2 / 1 rows, each one pass over W (a second pass over a
        // sint8 matrix of the BASELINE size is served from L2)
        if (k->prm.batch > 1) return fail(ctx, B200MM_ERR_INVALID, "%s: batch > 1 needs M in {1, 2, 4%s}", b200mm_kernel_name(k->id), quant ? "" : ", 8");
        if (k->prm.flags & B200MM_F_PEER_STORE) return fail(ctx, B200MM_ERR_INVALID, "%s: peer stores need M == 1", b200mm_kernel_name(k->id));
        size_t row0 = 0;
        for (size_t c : {(size_t)8, (size_t)4, (size_t)2, (size_t)1}) {
            if (c == 8 && quant) continue;
            if (c > 1 && k->prm.group_k) continue;  // the per-group-scale kernel is single-row: one pass per row of x
            while (M - row0 >= c) {
                b200mm_kernel* child = nullptr;
                b200mm_kernel_params prm = k->prm;
                prm.flags &= ~B200MM_F_AUTOTUNE;
                int rc = b200mm_kernel_get(ctx, k->id, c, N, K, &prm, &child);
                if (rc) return rc;
                k->chunks.push_back({child, row0});
                row0 += c;
            }
        }
        k->grid = k->chunks[0].kern->grid;
        k->block = k->chunks[0].kern->block;
        return B200MM_OK;
    }
    const int mrows = (int)M;
    if (N % cols || K % 4)
        return fail(ctx, B200MM_ERR_INVALID, "%s needs N%%%d==0 and K%%4==0", b200mm_kernel_name(k->id), cols);
    if (N > INT32_MAX || K > INT32_MAX) return fail(ctx, B200MM_ERR_INVALID, "shape too large");
    // tune[0]: 0 = default for the weight type (measured on B200, tools/sweep_gemv.py), 100 = variant 0, else the variant id
    // sint8 default = 13: the 8-warp / 256-column geometry of variant 4 with both register buffers in flight (12.75 vs 13.45 us at cfg4)
    k->gemv_variant = k->prm.tune[0] == 0 ? (quant ? 13 : 5) : (k->prm.tune[0] == 100 ? 0 : (int)k->prm.tune[0]);
    if (k->prm.tune[0] == 0 && quant && !k->prm.group_k && M == 1) {
        // small sint8 matrices (e.g. the per-rank panels of an N-sharded run): with 256-column panels x 8 K-splits there are fewer
        // CTAs than SMs and too few bytes in flight -- take the widest panel that still yields a CTA per SM (measured, tools/small_gemv.py:
        // 4096 x 1792: 7.4 -> 4.5 us with 64-column panels; 4096 x 3584: 7.7 -> 5.6 us with 128-column panels)
        const size_t sms = (size_t)ctx->prop.multiProcessorCount;
        if (ceil_div(N, 256) * 8 < sms) k->gemv_variant = ceil_div(N, 128) * 8 >= sms ? 21 : 25;
    }
    if (k->prm.tune[0] == 0 && !quant && M == 1) {
        // same for small fp32 matrices: 64-column panels instead of 128-column ones
        // (tools/small_gemv.py: 4096 x 2048 14.4 -> 7.7 us, 4096 x 4096 15.7 -> 12.0 us)
        if (ceil_div(N, 128) * 8 < (size_t)ctx->prop.multiProcessorCount * 2) k->gemv_variant = N <= 1024 ? 4 : 7;
    }
    const size_t group_k = k->prm.group_k;
    if (group_k) {  // per-(row block, column) scales stored behind the weights
        if (!quant) return fail(ctx, B200MM_ERR_INVALID, "group_k applies to qgemv_sint8 only");
        if (group_k != 32 && group_k != 64 && group_k % 128)
            return fail(ctx, B200MM_ERR_INVALID, "qgemv_sint8: group_k must be 32, 64 or a multiple of 128 (got %zu)", group_k);
    }
    // groups of 32 / 64 rows: plan the geometry with the blocked UNROLL 4 kernel (it is the one that runs when every split is whole)
    const bool try_blocked = group_k && group_k < 128 && mrows == 1 && getenv("B200MM_GEMV_NO_BLOCKED") == nullptr;
    const GemvPick pick = quant ? gemv_pick<GemvS8>(k->gemv_variant, mrows, group_k, try_blocked) : gemv_pick<GemvF32>(k->gemv_variant, mrows);
    const GemvFn fn = pick.fn;
    const int warps = pick.warps, lpr = pick.lpr;
    const int panel = lpr * cols;
    const unsigned batch = k->prm.batch ? k->prm.batch : 1;
    k->panels = (int)ceil_div(N, panel);
    if (k->prm.tune[3] >= 16) {
        // explicit panel count: the column groups are dealt evenly to that many panels (balanced grids, gemv.cuh) -- e.g. one
        // panel x split per CTA slot so that every SM carries the same load.  Each panel must fit the instantiation's width.
        const size_t groups = N / cols, want = k->prm.tune[3];
        if (want > groups || ceil_div(groups, want) > (size_t)lpr)
            return fail(ctx, B200MM_ERR_INVALID, "%s: %zu panels do not fit %zu column groups at %d groups per panel", b200mm_kernel_name(k->id), want, groups, lpr);
        k->panels = (int)want;
    }
    // K-splits: fill exactly ONE wave.  A second, partial wave runs at a fraction of the bandwidth (few CTAs, few
    // loads in flight) and was measured to cost 25 % at cfg3, so the split count is rounded DOWN to what is
    // co-resident: SMs x occupancy of this instantiation (registers, x staging + reduction smem).
    int splits = (int)k->prm.tune[1];
    if (splits <= 0) {
        int occ = 1;
        const size_t smem_guess = ((size_t)K / 4 + (size_t)(warps + 8) * panel) * sizeof(float);
        CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, warps * 32, smem_guess));
        occ = std::max(1, std::min(occ, 2048 / (warps * 32)));
        const size_t slots = (size_t)ctx->prop.multiProcessorCount * occ;
        splits = (int)(slots / ((size_t)k->panels * batch));
        splits = std::max(1, std::min(splits, 64));
        if (quant && mrows == 1) {  // (grouped scales too: equal, whole splits are also what the blocked row order needs -- cfg4 group_k = 128: 5 splits 19.1 us, 4 splits 13.55 us)
            // sint8 at 2 CTAs/SM: a power-of-two count keeps the K-splits equal (no guarded remainder) and leaves CTA slots free
            // for the next programmatic-dependent launch to become resident and prefetch (cfg4: 4 splits 12.75 us, 5 splits 18.2 us)
            int p2 = 1;
            while (p2 * 2 <= splits) p2 *= 2;
            splits = p2;
        }
        splits = (int)std::min<size_t>(splits, std::max<size_t>(1, K / 32));
    }
    // K-splits of a panel are reduced inside a thread-block cluster (<= 8 CTAs, portable size) unless tune[3] == 1
    k->gemv_cluster = (k->prm.tune[3] != 1) || mrows > 1;  // the ticket path exists for M == 1 only; tune[3] >= 16 is a panel count
    if (mrows > 1 && splits > 8) splits = 8;
    if (k->gemv_cluster && splits > 8) {
        if (k->prm.tune[1] > 0)
            k->gemv_cluster = false;  // an explicit split count above 8 keeps the ticket path
        else
            splits = 8;
    }
    // grouped: splits start on a pipeline-window boundary (blocked row order: on a 128-row boundary, so that every warp's block is whole)
    const int rstep = group_k ? (try_blocked ? 128 : (int)std::min<size_t>(group_k, 128)) : warps * (32 / lpr);
    auto smem_for = [&](size_t rows) {
        size_t b = ((size_t)mrows * rows + (size_t)warps * mrows * panel + (size_t)8 * mrows * panel) * sizeof(float);  // x, warp partials, 8 cluster receive slots
        if (group_k) b += ((rows / group_k + 2) * panel + (size_t)warps * 32 * cols) * sizeof(float);  // scales + per-thread totals
        return b;
    };
    auto rows_for = [&](int sp) { return ceil_div(ceil_div(K, (size_t)sp), (size_t)rstep) * rstep; };
    if (k->gemv_cluster && k->prm.tune[1] == 0 && splits > 1) {
        // A cluster lives inside one GPC, so "CTAs <= SMs x occupancy" does not guarantee that all clusters are co-resident
        // (measured at cfg4: 56 clusters of 5 at 2 CTAs/SM spill into a second wave, 18.2 us instead of 12.8 us with 4).
        // Ask the driver how many clusters of each size fit and take the largest split count that stays in one wave.
        const bool dbg = getenv("B200MM_DEBUG_GEMV") != nullptr;
        for (; splits > 1; --splits) {
            const size_t rows = rows_for(splits);
            const int sp = (int)ceil_div(K, rows);
            const size_t smem = smem_for(rows);
            if (smem > 200 * 1024) continue;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(k->panels, sp, batch);
            cfg.blockDim = dim3(warps * 32, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1;
            attr[0].val.clusterDim.y = (unsigned)sp;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int max_clusters = 0;
            CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
            CU_TRY(ctx, cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
            if (dbg) fprintf(stderr, "[b200mm] gemv %zux%zu variant %d: clusters of %d -> %d co-resident (need %zu)\n", K, N, k->gemv_variant, sp, max_clusters, (size_t)k->panels * batch);
            if ((size_t)max_clusters >= (size_t)k->panels * batch) break;
        }
    }
    size_t rps = rows_for(splits);
    splits = (int)ceil_div(K, rps);
    k->splits = splits;
    k->rows_per_split = (int)rps;
    // grouped scales: contiguous rows per warp (gemv.cuh BLOCKED) when every split is whole -- one fold per group and warp instead of
    // one per pipeline window
    k->gemv_blocked = group_k >= 32 && mrows == 1 && K % rps == 0 && rps % 128 == 0 && k->gemv_variant != 11 && k->gemv_variant != 12 &&
                      k->gemv_variant != 14 && getenv("B200MM_GEMV_NO_BLOCKED") == nullptr;
    const GemvFn fn_launch = quant ? gemv_pick<GemvS8>(k->gemv_variant, mrows, group_k, k->gemv_blocked).fn : fn;
    k->grid = dim3(k->panels, splits, batch);
    k->block = dim3(warps * 32, 1, 1);
    k->smem = smem_for(rps);
    if (k->smem > 200 * 1024) return fail(ctx, B200MM_ERR_INVALID, "gemv: K-split too long for shared memory");
    CU_TRY(ctx, cudaFuncSetAttribute(fn_launch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(k->smem, 48 * 1024)));
    if (splits > 1 && !k->gemv_cluster) {
        const size_t pbytes = (size_t)batch * splits * N * sizeof(float);
        const size_t tbytes = (size_t)batch * k->panels * sizeof(unsigned int);
        k->ws_bytes = ceil_div(pbytes, 256) * 256 + tbytes;
        CU_TRY(ctx, cudaMalloc(&k->ws, k->ws_bytes));
        k->partial = (float*)k->ws;
        k->tickets = (unsigned int*)((char*)k->ws + ceil_div(pbytes, 256) * 256);
        CU_TRY(ctx, cudaMemsetAsync(k->ws, 0, k->ws_bytes, ctx->stream));
    }
    return B200MM_OK;
}

// B200MM_F_AUTOTUNE: the best (instantiation, K-split count) of the streaming GEMV depends on how whole clusters pack into
// GPCs and on how much room the NEXT programmatic-dependent launch finds to become resident early -- at cfg4 neighbouring
// split counts differ by 40 % and no occupancy formula predicts the order.  So measure: every candidate is timed with
// back-to-back launches over scratch weight sets that together exceed L2, and the fastest one is kept.
static int gemv_autotune(b200mm_ctx* ctx, b200mm_kernel* k, bool quant) {
    const size_t K = k->K, N = k->N;
    const size_t batch = k->prm.batch ? k->prm.batch : 1;
    const size_t group_k = k->prm.group_k;
    const size_t wbytes = batch * (quant ? K * N + (group_k ? ceil_div(K, group_k) * N * 4 : 0) : K * N * 4);
    const int nsets = (int)std::max<size_t>(2, std::min<size_t>(8, ceil_div((size_t)320 << 20, wbytes)));
    char* W = nullptr;
    float *x = nullptr, *y = nullptr;
    CU_TRY(ctx, cudaMalloc(&W, wbytes * nsets));
    CU_TRY(ctx, cudaMalloc(&x, batch * K * sizeof(float)));
    CU_TRY(ctx, cudaMalloc(&y, batch * N * sizeof(float)));
    CU_TRY(ctx, cudaMemsetAsync(W, 0x11, wbytes * nsets, ctx->stream));  // any byte pattern is a valid int8 / finite f32 weight
    CU_TRY(ctx, cudaMemsetAsync(x, 0, batch * K * sizeof(float), ctx->stream));
    cudaEvent_t e0, e1;
    CU_TRY(ctx, cudaEventCreate(&e0));
    CU_TRY(ctx, cudaEventCreate(&e1));
    // small fp32 matrices (the per-rank panels of an N-sharded run: 32 MiB at 8 GPUs) are short of bytes in flight with the
    // 128-column panels of the large-matrix geometries: the 64-column (16-lane) instantiations double the CTA count
    static const int f32_variants[] = {5, 100, 2, 3, 1, 4, 6, 7};
    static const int s8_variants[] = {4, 11, 12, 13, 21, 25};  // the last two (narrow panels = more CTAs) only for small matrices
    static const int s8g_variants[] = {4, 11, 12, 14};
    const int* variants = quant ? (group_k ? s8g_variants : s8_variants) : f32_variants;
    const int nvar = quant ? ((!group_k && wbytes <= ((size_t)24 << 20)) ? 6 : 4) : (wbytes <= ((size_t)96 << 20) ? 8 : 4);
    const uint64_t launches_before = ctx->launches;
    float best_ms = 1e30f, default_ms = 1e30f;
    uint32_t best_v = 0, best_s = 0;
    // the configuration the built-in rule chose (setup_gemv ran before us): kept unless a candidate beats it by > 3 %
    const uint32_t default_v = k->gemv_variant == 0 ? 100u : (uint32_t)k->gemv_variant, default_s = (uint32_t)k->splits;
    auto same_kernel = [&](uint32_t a, uint32_t b) {  // several variant ids select the same instantiation
        if (a == b) return true;
        return quant && group_k && (a == 4 || a == 13) && (b == 4 || b == 13);
    };
    const bool dbg = getenv("B200MM_DEBUG_GEMV") != nullptr;
    int rc = B200MM_OK;
    for (int vi = 0; vi < nvar && rc == B200MM_OK; ++vi)
        for (uint32_t sp = 1; sp <= 8 && rc == B200MM_OK; ++sp) {
            if (sp > K / 32 && sp > 1) break;
            b200mm_kernel t;
            t.id = k->id;
            t.M = 1;
            t.N = N;
            t.K = K;
            t.prm = k->prm;
            t.prm.flags &= ~(B200MM_F_AUTOTUNE | B200MM_F_PEER_STORE);
            t.prm.tune[0] = (uint32_t)variants[vi];
            t.prm.tune[1] = sp;
            if (setup_gemv(ctx, &t, quant) != B200MM_OK) continue;  // e.g. the split does not fit shared memory
            if ((uint32_t)t.splits != sp) {                         // rounding merged it into a smaller count: already timed
                if (t.ws) cudaFree(t.ws);
                continue;
            }
            const int reps = 64;
            float ms = 1e30f;
            for (int round = 0; round < 4 && rc == B200MM_OK; ++round) {  // round 0 warms up; best of 3 timed rounds
                if (round) cudaEventRecord(e0, ctx->stream);
                for (int i = 0; i < reps && rc == B200MM_OK; ++i) rc = b200mm_launch_ptr(ctx, &t, x, W + (size_t)(i % nsets) * wbytes, y, nullptr);
                if (round) {
                    cudaEventRecord(e1, ctx->stream);
                    if (cudaEventSynchronize(e1) != cudaSuccess) rc = fail(ctx, B200MM_ERR_CUDA, "gemv autotune: %s", cudaGetErrorString(cudaGetLastError()));
                    float m = 0.f;
                    cudaEventElapsedTime(&m, e0, e1);
                    ms = std::min(ms, m / reps);
                }
            }
            if (dbg) fprintf(stderr, "[b200mm] autotune %zux%zu variant %d splits %u: %.2f us\n", K, N, variants[vi], sp, ms * 1e3f);
            if (rc == B200MM_OK && same_kernel((uint32_t)variants[vi], default_v) && sp == default_s) default_ms = ms;
            if (rc == B200MM_OK && ms < best_ms) {
                best_ms = ms;
                best_v = (uint32_t)variants[vi];
                best_s = sp;
            }
            cudaStreamSynchronize(ctx->stream);
            if (t.ws) cudaFree(t.ws);
        }
    // balanced grids (sint8, global scale): SMs x 2 CTA slots dealt as (panels x splits) with ragged 128-column panels
    uint32_t best_p = 0;
    if (quant && !group_k && rc == B200MM_OK) {
        const uint32_t slots = (uint32_t)ctx->prop.multiProcessorCount * 2;
        for (uint32_t sp : {1u, 2u, 4u}) {
            const uint32_t panels = slots / sp;
            if (panels < 16 || panels > N / 16 || ceil_div(N / 16, panels) > 8 || sp > K / 32) continue;
            b200mm_kernel t;
            t.id = k->id;
            t.M = 1;
            t.N = N;
            t.K = K;
            t.prm = k->prm;
            t.prm.flags &= ~(B200MM_F_AUTOTUNE | B200MM_F_PEER_STORE);
            t.prm.tune[0] = 21;
            t.prm.tune[1] = sp;
            t.prm.tune[3] = panels;
            if (setup_gemv(ctx, &t, quant) != B200MM_OK || (uint32_t)t.splits != sp) {
                if (t.ws) cudaFree(t.ws);
                continue;
            }
            float ms = 1e30f;
            for (int round = 0; round < 4 && rc == B200MM_OK; ++round) {
                if (round) cudaEventRecord(e0, ctx->stream);
                for (int i = 0; i < 64 && rc == B200MM_OK; ++i) rc = b200mm_launch_ptr(ctx, &t, x, W + (size_t)(i % nsets) * wbytes, y, nullptr);
                if (round) {
                    cudaEventRecord(e1, ctx->stream);
                    if (cudaEventSynchronize(e1) != cudaSuccess) rc = fail(ctx, B200MM_ERR_CUDA, "gemv autotune: %s", cudaGetErrorString(cudaGetLastError()));
                    float m = 0.f;
                    cudaEventElapsedTime(&m, e0, e1);
                    ms = std::min(ms, m / 64);
                }
            }
            if (dbg) fprintf(stderr, "[b200mm] autotune %zux%zu balanced: variant 21, %u panels x %u splits: %.2f us\n", K, N, panels, sp, ms * 1e3f);
            if (rc == B200MM_OK && ms < best_ms) {
                best_ms = ms;
                best_v = 21;
                best_s = sp;
                best_p = panels;
            }
            cudaStreamSynchronize(ctx->stream);
            if (t.ws) cudaFree(t.ws);
        }
    }
    cudaStreamSynchronize(ctx->stream);
    ctx->launches = launches_before;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(W);
    cudaFree(x);
    cudaFree(y);
    if (rc != B200MM_OK) return rc;
    if (best_s == 0) return fail(ctx, B200MM_ERR_INVALID, "gemv autotune: no candidate geometry fits");
    if (default_ms < 1e29f && best_ms > 0.97f * default_ms) {  // within measurement noise of the rule's choice: keep the rule
        best_v = default_v;
        best_s = default_s;
        best_ms = default_ms;
        best_p = 0;
    }
    k->prm.tune[0] = best_v;
    k->prm.tune[1] = best_s;
    if (best_p) k->prm.tune[3] = best_p;
    if (dbg) fprintf(stderr, "[b200mm] autotune %zux%zu -> variant %u, %u splits, %.2f us\n", K, N, best_v, best_s, best_ms * 1e3f);
    return B200MM_OK;
}

extern "C" int b200mm_kernel_get(b200mm_ctx* ctx, int kernel_id, size_t M, size_t N, size_t K,
                                 const b200mm_kernel_params* params, b200mm_kernel** out) {
    if (!ctx || !out) return fail(ctx, B200MM_ERR_INVALID, "kernel_get: NULL argument");
    *out = nullptr;
    if (M == 0 || N == 0 || K == 0) return fail(ctx, B200MM_ERR_INVALID, "kernel_get: empty shape %zux%zux%zu", M, N, K);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    b200mm_kernel* k = new b200mm_kernel();
    k->id = kernel_id;
    k->M = M;
    k->N = N;
    k->K = K;
    if (params) k->prm = *params;
    int rc;
    if (kernel_id >= B200MM_K_GEMM_1 && kernel_id <= B200MM_K_QGEMV_1)
        rc = setup_ports(ctx, k);
    else if (kernel_id == B200MM_K_SGEMM_SIMT)
        rc = setup_simt(ctx, k);
    else if (kernel_id == B200MM_K_SGEMM_TC3X)
        rc = setup_tc3x(ctx, k);
    else if (kernel_id == B200MM_K_GEMV_F32 || kernel_id == B200MM_K_QGEMV_SINT8) {
        const bool quant = kernel_id == B200MM_K_QGEMV_SINT8;
        rc = setup_gemv(ctx, k, quant);  // validates the shape; also the result when no tuning is asked for
        if (rc == B200MM_OK && (k->prm.flags & B200MM_F_AUTOTUNE) && M == 1 && k->prm.tune[0] == 0 && k->prm.tune[1] == 0) {
            if (k->ws) cudaFree(k->ws);
            k->ws = nullptr;
            k->ws_bytes = 0;
            k->partial = nullptr;
            k->tickets = nullptr;
            rc = gemv_autotune(ctx, k, quant);
            if (rc == B200MM_OK) rc = setup_gemv(ctx, k, quant);
        }
    }
    else
        rc = fail(ctx, B200MM_ERR_UNSUPPORTED, "unknown kernel id %d", kernel_id);
    if (rc) {
        if (k->inner) b200mm_kernel_free(ctx, k->inner);
        for (auto& c : k->chunks) b200mm_kernel_free(ctx, c.kern);
        if (k->ws) cudaFree(k->ws);
        delete k;
        return rc;
    }
    *out = k;
    return B200MM_OK;
}

extern "C" int b200mm_kernel_free(b200mm_ctx* ctx, b200mm_kernel* kern) {
    if (!kern) return B200MM_OK;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (kern->panel) b200mm_kernel_free(ctx, kern->panel);
    if (kern->inner) b200mm_kernel_free(ctx, kern->inner);
    for (auto& c : kern->chunks) b200mm_kernel_free(ctx, c.kern);
    if (kern->ws) cudaFree(kern->ws);
    for (auto e : kern->pev) cudaEventDestroy(e);
    delete kern;
    return B200MM_OK;
}

static constexpr int kProfRing = 256;

extern "C" int b200mm_kernel_profile_enable(b200mm_ctx* ctx, b200mm_kernel* k, int enable) {
    if (!ctx || !k) return fail(ctx, B200MM_ERR_INVALID, "profile_enable: NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (enable && k->pev.empty()) {
        k->pev.resize(2 * kProfRing);
        for (auto& e : k->pev) CU_TRY(ctx, cudaEventCreate(&e));
    }
    k->profiling = enable != 0;
    k->pcount = 0;
    return B200MM_OK;
}

extern "C" int b200mm_kernel_profile_read(b200mm_ctx* ctx, b200mm_kernel* k, float* ms_out, int max_n, int* n_out) {
    if (!ctx || !k || !ms_out || !n_out) return fail(ctx, B200MM_ERR_INVALID, "profile_read: NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const int n = std::min(std::min(k->pcount, kProfRing), max_n);
    const int first = k->pcount - n;
    for (int i = 0; i < n; ++i) {
        const int slot = (first + i) % kProfRing;
        CU_TRY(ctx, cudaEventElapsedTime(&ms_out[i], k->pev[2 * slot], k->pev[2 * slot + 1]));
    }
    *n_out = n;
    k->pcount = 0;
    return B200MM_OK;
}

extern "C" int b200mm_kernel_geometry(const b200mm_kernel* k, uint32_t grid[3], uint32_t block[3]) {
    if (!k) return fail(nullptr, B200MM_ERR_INVALID, "kernel_geometry: kern is NULL");
    if (grid) grid[0] = k->grid.x, grid[1] = k->grid.y, grid[2] = k->grid.z;
    if (block) block[0] = k->block.x, block[1] = k->block.y, block[2] = k->block.z;
    return B200MM_OK;
}

extern "C" size_t b200mm_kernel_workspace_bytes(const b200mm_kernel* k) { return k ? k->ws_bytes : 0; }

extern "C" int b200mm_kernel_set_peers(b200mm_kernel* k, int rank, int world, void* const* peer_c, size_t ldc,
                                       size_t col_offset) {
    if (!k) return fail(nullptr, B200MM_ERR_INVALID, "set_peers: kern is NULL");
    if (k->id != B200MM_K_SGEMM_SIMT && k->id != B200MM_K_SGEMM_TC3X && k->id != B200MM_K_GEMV_F32 && k->id != B200MM_K_QGEMV_SINT8)
        return fail(nullptr, B200MM_ERR_INVALID, "set_peers: only the B200-native SGEMM / GEMV kernels store to peers");
    if (world < 0 || world > 8 || rank < 0 || (world && rank >= world)) return fail(nullptr, B200MM_ERR_INVALID, "set_peers: bad rank/world");
    if (world && (ldc % 4 || col_offset % 4)) return fail(nullptr, B200MM_ERR_INVALID, "set_peers: ldc and col_offset must be multiples of 4");
    k->peers = PeerStore{};
    k->tc_c_src = nullptr;  // the TMA store map addresses the local C of the peer set: rebuild it
    k->peers.world = world;
    k->peers.rank = rank;
    k->peers.ldc = ldc;
    k->peers.col0 = col_offset;
    for (int i = 0; i < world; ++i) {
        if (!peer_c[i]) return fail(nullptr, B200MM_ERR_INVALID, "set_peers: peer %d is NULL", i);
        k->peers.c[i] = (float*)peer_c[i];
    }
    return B200MM_OK;
}

extern "C" int b200mm_kernel_set_peer_flags(b200mm_kernel* k, void* const* peer_flags, size_t pingpong_stride, int deferred) {
    if (!k) return fail(nullptr, B200MM_ERR_INVALID, "set_peer_flags: kern is NULL");
    if (k->id != B200MM_K_GEMV_F32 && k->id != B200MM_K_QGEMV_SINT8)
        return fail(nullptr, B200MM_ERR_INVALID, "set_peer_flags: only the streaming GEMV kernels complete across ranks in-kernel");
    if (k->peers.world < 2) return fail(nullptr, B200MM_ERR_INVALID, "set_peer_flags: call b200mm_kernel_set_peers first (world >= 2)");
    for (int i = 0; i < 8; ++i) k->peers.flags[i] = nullptr;
    k->peers.deferred = 0;
    k->peer_epoch = 0;
    k->peer_pingpong = 0;
    if (!peer_flags) return B200MM_OK;  // back to "caller synchronises the ranks" (b200mm_peer_barrier / a collective)
    for (int i = 0; i < k->peers.world; ++i) {
        if (!peer_flags[i]) return fail(nullptr, B200MM_ERR_INVALID, "set_peer_flags: peer %d is NULL", i);
        k->peers.flags[i] = (unsigned int*)peer_flags[i];
    }
    k->peer_pingpong = pingpong_stride;
    k->peers.deferred = deferred ? 1u : 0u;
    return B200MM_OK;
}

extern "C" int b200mm_kernel_peer_wait(b200mm_ctx* ctx, b200mm_kernel* k) {
    if (!ctx || !k) return fail(ctx, B200MM_ERR_INVALID, "peer_wait: NULL argument");
    if (k->peers.world < 2 || !k->peers.flags[0]) return fail(ctx, B200MM_ERR_INVALID, "peer_wait: no peer flags set on this kernel");
    if (k->peer_epoch == 0) return B200MM_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(k->peers, k->peer_epoch);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200MM_OK;
}

extern "C" unsigned int b200mm_kernel_peer_epoch(const b200mm_kernel* k) { return k ? k->peer_epoch : 0; }

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
template <class Cfg>
static cudaError_t launch_tc3x(b200mm_kernel* k, cudaStream_t s, const float* A, float* C) {
    Tc3xArgs a{};
    a.A = A;
    a.A_lo = k->a_lo;
    a.band_cnt = k->tc_band_cnt;
    a.prebands = k->tc_prebands;
    a.C = C;
    a.M = (int)k->M;
    a.N = (int)k->N;
    a.K = (int)k->K;
    a.ldc = (int)k->N;
    a.tiles_m = (int)ceil_div(k->M, Cfg::TILE_M);
    a.tiles_n = (int)ceil_div(k->N, Cfg::BN);
    a.chains_per_tile = k->tc_cpt;
    a.full_waves = k->tc_full_waves;
    a.sk_units = k->tc_sk_units;
    a.partial = k->tc_partial;
    a.flags = k->tc_flags;
    a.epoch = ++k->tc_epoch;
    static_assert(Cfg::CHAIN * Cfg::BK == 256, "setup_tc3x assumes chains of 256 k");
    a.peers = k->peers;
    // Cooperative launch: the in-kernel A split makes every CTA's producer wait on the splitter warps of ALL CTAs, so the
    // whole grid (<= one CTA per SM) must be co-resident; the driver refuses the launch otherwise instead of letting it hang.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = k->grid;
    cfg.blockDim = k->block;
    cfg.dynamicSmemBytes = k->smem;
    cfg.stream = s;
    static const bool pdl_off = getenv("B200MM_TC3X_NO_PDL") != nullptr;  // experiment knob
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = (k->tc_prebands < k->tc_bands) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!attr[0].val.cooperative && !pdl_off) {
        // not cooperative (no in-kernel split): launch programmatically dependent on the split pass instead, so that the GEMM's
        // prologue overlaps its tail (the kernel waits in griddepcontrol.wait before it touches any operand)
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
    }
    if constexpr (Cfg::CTA2) {  // CTA pairs: a cluster of two CTAs is placed on one TPC
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    return cudaLaunchKernelEx(&cfg, sgemm_tc3x_kernel<Cfg>, k->tmAh, k->tmAl, k->tmBh, k->tmBl, k->tmC, a);
}

extern "C" int b200mm_launch_ptr(b200mm_ctx* ctx, b200mm_kernel* k, const void* A, const void* B, void* C,
                                 const uint32_t grid_in[3]) {
    if (!ctx || !k || !A || !B || !C) return fail(ctx, B200MM_ERR_INVALID, "launch: NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const unsigned M = (unsigned)k->M, N = (unsigned)k->N, K = (unsigned)k->K;
    dim3 grid = k->grid;
    if (k->id < 32 && grid_in) {
        grid = dim3(grid_in[0], grid_in[1], grid_in[2]);
        if (grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535)
            return fail(ctx, B200MM_ERR_LIMITS, "Compute limits exceeded");
    }
    const float* Af = (const float*)A;
    const float* Bf = (const float*)B;
    float* Cf = (float*)C;
    if (!k->chunks.empty()) {  // skinny GEMM in row chunks (see b200mm_kernel::chunks)
        for (auto& c : k->chunks) {
            int rc = b200mm_launch_ptr(ctx, c.kern, Af + c.row0 * k->K, B, Cf + c.row0 * k->N, nullptr);
            if (rc) return rc;
        }
        return B200MM_OK;
    }
    if (k->inner) {  // ragged N / K through zero-padded staging copies (see b200mm_kernel::inner)
        const size_t Kp = k->pad_k, Np = k->pad_n;
        CU_TRY(ctx, cudaMemcpy2DAsync(k->pad_a, Kp * 4, A, k->K * 4, k->K * 4, k->M, cudaMemcpyDeviceToDevice, s));
        CU_TRY(ctx, cudaMemcpy2DAsync(k->pad_b, Np * 4, B, k->N * 4, k->N * 4, k->K, cudaMemcpyDeviceToDevice, s));
        int rc = b200mm_launch_ptr(ctx, k->inner, k->pad_a, k->pad_b, k->pad_c, nullptr);
        if (rc) return rc;
        CU_TRY(ctx, cudaMemcpy2DAsync(C, k->N * 4, k->pad_c, Np * 4, k->N * 4, k->M, cudaMemcpyDeviceToDevice, s));
        return B200MM_OK;
    }
    const int pslot = k->pcount % kProfRing;
    auto prof_begin = [&]() {
        if (k->profiling) cudaEventRecord(k->pev[2 * pslot], s);
    };
    auto prof_end = [&]() {
        if (k->profiling) {
            cudaEventRecord(k->pev[2 * pslot + 1], s);
            k->pcount++;
        }
    };
    if (k->id != B200MM_K_SGEMM_TC3X) prof_begin();
    switch (k->id) {
        case B200MM_K_GEMM_1: wgsl::gemm_1<<<grid, k->block, 0, s>>>(Af, Bf, Cf, M, N, K); break;
        case B200MM_K_GEMM_1V: wgsl::gemm_1v<<<grid, k->block, 0, s>>>((const float4*)A, (const float4*)B, (float4*)C, M, N, K); break;
        case B200MM_K_GEMM_2: wgsl::gemm_2<<<grid, k->block, 0, s>>>(Af, Bf, Cf, M, N, K); break;
        case B200MM_K_GEMM_3: wgsl::gemm_3<<<grid, k->block, 0, s>>>(Af, Bf, Cf, M, N, K); break;
        case B200MM_K_GEMM_4: wgsl::gemm_4<<<grid, k->block, 0, s>>>(Af, Bf, Cf, M, N, K); break;
        case B200MM_K_GEMM_5: wgsl::gemm_5<<<grid, k->block, 0, s>>>(Af, Bf, Cf, M, N, K); break;
        case B200MM_K_GEMM_WONNX: wgsl::gemm_wonnx<<<grid, k->block, 0, s>>>((const float4*)A, (const float4*)B, (float4*)C, M, N, K); break;
        case B200MM_K_BRAM:
        case B200MM_K_BRAM8X8: wgsl::bram<<<grid, k->block, 0, s>>>((const float4*)A, (const float4*)B, (float4*)C, M, N, K); break;
        case B200MM_K_GEMM3: wgsl::gemm3<<<grid, k->block, 0, s>>>((const float4*)A, (const float4*)B, (float4*)C, M, N, K); break;
        case B200MM_K_QGEMV_1:
            wgsl::qgemv_1<<<grid, k->block, 0, s>>>((const float4*)A, (const uint32_t*)B, (float4*)C, N, K, k->prm.absmax, k->prm.batch ? k->prm.batch : 1u);
            break;
        case B200MM_K_SGEMM_SIMT: {
            const bool aligned = (M % SimtCfg::BM == 0) && (N % SimtCfg::BN == 0) && (K % SimtCfg::BK == 0);
            k->simt.epoch++;
            if (!aligned && k->peers.world) return fail(ctx, B200MM_ERR_INVALID, "peer stores need tile-aligned shapes");
            for (int pass = 0; pass < 2; ++pass) {
                SimtSched sc = k->simt;
                const int ntiles = pass == 0 ? k->simt_tiles1 : k->simt_tiles2;
                if (ntiles == 0) continue;
                sc.tile_offset = pass == 0 ? 0 : k->simt_tiles1;
                sc.split = pass == 0 ? 1 : k->simt.split;
                const dim3 g((unsigned)ntiles * sc.split, 1, 1);
                if (aligned)
                    sgemm_simt_kernel<false><<<g, k->block, 0, s>>>(Af, Bf, Cf, (int)M, (int)N, (int)K, (int)N, k->peers, sc);
                else
                    sgemm_simt_kernel<true><<<g, k->block, 0, s>>>(Af, Bf, Cf, (int)M, (int)N, (int)K, (int)N, k->peers, sc);
                if (pass == 1 && k->simt_tiles1) ctx->launches++;
            }
            break;
        }
        case B200MM_K_SGEMM_TC3X: {
            const bool one_pass = (k->prm.flags & B200MM_F_TC3X_1X) != 0;
            const int sms = ctx->prop.multiProcessorCount;
            const size_t a4 = k->M * k->K / 4, b4 = k->K * k->N / 4;
            if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return fail(ctx, B200MM_ERR_INVALID, "sgemm_tc3x needs 16-byte aligned A, B, C");
            // hi operands = the caller's buffers (the tensor core truncates them to tf32); re-encode the maps when they move
            const void* b_src = B;
            if (k->tc_b_copy) {
                CU_TRY(ctx, cudaMemcpyAsync(k->b_hi, B, b4 * 16, cudaMemcpyDeviceToDevice, s));
                b_src = k->b_hi;
            }
            int rc;
            if (k->tc_a_src != A) {
                if ((rc = make_tmap_kmajor(ctx, &k->tmAh, (const float*)A, k->M, k->K, 128, k->tc_bk))) return rc;
                k->tc_a_src = A;
            }
            if (k->tc_b_src != b_src) {
                if ((rc = make_tmap_mnmajor(ctx, &k->tmBh, (const float*)b_src, k->K, k->N, k->tc_bk, k->tc_cta2 ? k->tc_bn / 2 : k->tc_bn))) return rc;
                k->tc_b_src = b_src;
            }
            cudaError_t le;
            if (one_pass) {
                k->tmAl = k->tmAh;
                k->tmBl = k->tmBh;
                prof_begin();
                le = launch_tc3x<Tc256x1>(k, s, Af, Cf);
            } else {
                // pre-pass: B (unless its lo part is still valid) and the first row bands of A; the rest of A is split in-kernel
                const size_t a4_pre = std::min<size_t>(a4, (size_t)k->tc_prebands * kTc3xBandRows * k->K / 4);
                bool skip_b = k->tc_skip_b_split || k->tc_split >= 1;
                if (k->prm.flags & B200MM_F_CONST_B) {
                    skip_b = skip_b || (k->tc_const_b == B);
                    k->tc_const_b = B;
                }
                if (k->tc_split < 2) {
                    split_lo_kernel<<<sms * 8, 256, 0, s>>>((const float4*)A, (float4*)k->a_lo, a4_pre, (const float4*)B, (float4*)k->b_lo,
                                                            skip_b ? 0 : b4);
                    ctx->launches += 1;
                }
                prof_begin();
                if (k->tc_tma_store) {
                    // store map: this rank's panel of the (local) C -- 32 x 32 boxes, SWIZZLE_128B; rebuilt when C moves
                    float* cbase = k->peers.world ? k->peers.c[k->peers.rank] + k->peers.col0 : Cf;
                    const size_t ldc = k->peers.world ? k->peers.ldc : k->N;
                    if (k->tc_c_src != cbase) {
                        if ((rc = make_tmap_c(ctx, &k->tmC, cbase, k->M, k->N, ldc))) return rc;
                        k->tc_c_src = cbase;
                    }
                    le = k->tc_split == 2   ? launch_tc3x<Tc256k16x2sab>(k, s, Af, Cf)
                         : k->tc_split == 1 ? launch_tc3x<Tc256k16x2sb>(k, s, Af, Cf)
                                            : launch_tc3x<Tc256k16x2s>(k, s, Af, Cf);
                } else if (k->tc_split && k->tc_bn == 128)
                    le = launch_tc3x<Tc128b>(k, s, Af, Cf);
                else if (k->tc_split == 2)
                    le = launch_tc3x<Tc256k16ab>(k, s, Af, Cf);
                else if (k->tc_split == 1)
                    le = launch_tc3x<Tc256k16b>(k, s, Af, Cf);
                else if (k->tc_cta2 && k->tc_bk == 32)
                    le = launch_tc3x<Tc256k32x2>(k, s, Af, Cf);
                else if (k->tc_cta2)
                    le = launch_tc3x<Tc256k16x2>(k, s, Af, Cf);
                else if (k->tc_bn == 256 && k->tc_bk == 16)
                    le = launch_tc3x<Tc256k16>(k, s, Af, Cf);
                else if (k->tc_bn == 256)
                    le = launch_tc3x<Tc256>(k, s, Af, Cf);
                else
                    le = launch_tc3x<Tc128>(k, s, Af, Cf);
            }
            if (le != cudaSuccess) return fail(ctx, B200MM_ERR_CUDA, "sgemm_tc3x launch failed: %s", cudaGetErrorString(le));
            break;
        }
        case B200MM_K_GEMV_F32:
        case B200MM_K_QGEMV_SINT8: {
            const bool quant = k->id == B200MM_K_QGEMV_SINT8;
            const GemvFn fn = (quant ? gemv_pick<GemvS8>(k->gemv_variant, (int)k->M, k->prm.group_k, k->gemv_blocked) : gemv_pick<GemvF32>(k->gemv_variant, (int)k->M)).fn;
            const size_t group_k = k->prm.group_k;
            const float scale = quant ? (group_k ? 1.0f : k->prm.absmax) / 127.0f : 1.0f;
            const size_t wstride = quant ? (size_t)K * N + (group_k ? ceil_div(K, group_k) * N * 4 : 0) : (size_t)K * N * 4;
            if (k->peers.world && (k->prm.batch > 1 || k->M > 1)) return fail(ctx, B200MM_ERR_INVALID, "peer stores are not supported for batched / multi-row GEMV");
            if (((uintptr_t)A | (uintptr_t)B) & 15) return fail(ctx, B200MM_ERR_INVALID, "gemv: x and W must be 16-byte aligned (128-bit loads)");
            // launched with programmatic stream serialization (PDL): see the griddepcontrol comments in gemv.cuh
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = k->grid;
            cfg.blockDim = k->block;
            cfg.dynamicSmemBytes = k->smem;
            cfg.stream = s;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = (k->prm.tune[2] == 1) ? 0 : 1;  // tune[2] = 1 disables PDL
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            const bool cluster = k->gemv_cluster && k->splits > 1;
            if (cluster) {
                attr[1].id = cudaLaunchAttributeClusterDimension;
                attr[1].val.clusterDim.x = 1;
                attr[1].val.clusterDim.y = (unsigned)k->splits;
                attr[1].val.clusterDim.z = 1;
                cfg.numAttrs = 2;
            }
            PeerStore ps = k->peers;
            if (ps.world > 1 && ps.flags[0]) {
                ps.epoch = ++k->peer_epoch;
                ps.signal_ctas = (unsigned)k->panels;  // one storing CTA per column panel (batch == 1 with peers)
                ps.col0 += (size_t)(ps.epoch & 1u) * k->peer_pingpong;
            }
            int cluster_arg = cluster ? 1 : 0;
            unsigned int* tickets = k->tickets;
            if (k->trace_buf && (cluster || k->splits == 1)) {  // the ticket path needs `tickets` itself
                cluster_arg |= (k->trace_next % k->trace_slots + 1) << 8;
                k->trace_next++;
                tickets = (unsigned int*)k->trace_buf;
            }
            CU_TRY(ctx, cudaLaunchKernelEx(&cfg, fn, Af, (const void*)B, Cf, k->partial, tickets, (int)K, (int)N, k->rows_per_split, scale,
                                           (size_t)k->M * K, wstride, (size_t)k->M * N, ps, cluster_arg, (int)group_k));
            break;
        }
        default:
            return fail(ctx, B200MM_ERR_UNSUPPORTED, "unknown kernel id %d", k->id);
    }
    prof_end();
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200MM_OK;
}

static size_t b_bytes_needed(const b200mm_kernel* k) {
    const size_t batch = k->prm.batch ? k->prm.batch : 1;
    if (k->id == B200MM_K_QGEMV_SINT8 && k->prm.group_k) return batch * (k->K * k->N + ceil_div(k->K, (size_t)k->prm.group_k) * k->N * 4);
    if (k->id == B200MM_K_QGEMV_1 || k->id == B200MM_K_QGEMV_SINT8) return batch * k->K * k->N;
    if (k->id == B200MM_K_GEMV_F32) return batch * k->K * k->N * 4;
    return k->K * k->N * 4;
}

extern "C" int b200mm_launch(b200mm_ctx* ctx, b200mm_kernel* k, const b200mm_buffer* A, const b200mm_buffer* B,
                             b200mm_buffer* C, const uint32_t grid[3]) {
    if (!ctx || !k || !A || !B || !C) return fail(ctx, B200MM_ERR_INVALID, "launch: NULL argument");
    const size_t batch = (k->id == B200MM_K_QGEMV_1 || k->id == B200MM_K_QGEMV_SINT8 || k->id == B200MM_K_GEMV_F32)
                             ? (k->prm.batch ? k->prm.batch : 1)
                             : 1;
    // the reference binds whole buffers unchecked (src/harness.rs:179-184); we refuse undersized ones instead
    if (A->bytes < batch * k->M * k->K * 4) return fail(ctx, B200MM_ERR_INVALID, "launch: A is %zu bytes, kernel needs %zu", A->bytes, batch * k->M * k->K * 4);
    if (B->bytes < b_bytes_needed(k)) return fail(ctx, B200MM_ERR_INVALID, "launch: B is %zu bytes, kernel needs %zu", B->bytes, b_bytes_needed(k));
    if (C->bytes < batch * k->M * k->N * 4) return fail(ctx, B200MM_ERR_INVALID, "launch: C is %zu bytes, kernel needs %zu", C->bytes, batch * k->M * k->N * 4);
    return b200mm_launch_ptr(ctx, k, A->ptr, B->ptr, C->ptr, grid);
}

// End-to-end call with host buffers.  For the SGEMM kernels the transfers are pipelined over row panels of A / C on
// two copy streams: B and the A panels go up back to back, panel i is multiplied as soon as it has landed, and C panel
// i goes down while panel i+1 is being multiplied -- the call approaches the H2D floor instead of paying
// H2D + compute + D2H in series.  (The reference pays them in series too: create_buffer_init, dispatch, to_cpu,
// src/harness.rs:40-62.)
extern "C" int b200mm_mm_host(b200mm_ctx* ctx, b200mm_kernel* kern, const void* hostA, size_t bytesA, const void* hostB,
                              size_t bytesB, void* hostC, size_t bytesC, b200mm_buffer* dA, b200mm_buffer* dB,
                              b200mm_buffer* dC) {
    if (!ctx || !kern || !hostA || !hostB || !hostC || !dA || !dB || !dC) return fail(ctx, B200MM_ERR_INVALID, "mm_host: NULL argument");
    int rc;
    const bool sgemm = kern->id == B200MM_K_SGEMM_TC3X || kern->id == B200MM_K_SGEMM_SIMT;
    const size_t M = kern->M, N = kern->N, K = kern->K;
    int P = 0;
    if (sgemm && kern->peers.world == 0 && bytesA == M * K * 4 && bytesB == K * N * 4 && bytesC == M * N * 4) {
        // up to 16 panels of >= 256 rows (measured at 4096^3: 1 panel 4.19 ms, 4: 3.22, 8: 3.15, 16: 2.97; the H2D of A and B
        // alone is ~2.4 ms).  B200MM_HOST_PANELS overrides the cap for experiments.
        const char* env = getenv("B200MM_HOST_PANELS");
        const int want = env ? atoi(env) : 16;
        for (int cand : {16, 8, 4, 2})
            if (cand <= want && M % ((size_t)cand * 128) == 0 && M / cand >= 256) {
                P = cand;
                break;
            }
    }
    if (P == 0) {
        if ((rc = b200mm_buffer_write(ctx, dA, 0, hostA, bytesA))) return rc;
        if ((rc = b200mm_buffer_write(ctx, dB, 0, hostB, bytesB))) return rc;
        if ((rc = b200mm_launch(ctx, kern, dA, dB, dC, nullptr))) return rc;
        return b200mm_buffer_read(ctx, dC, 0, hostC, bytesC);
    }
    if (dA->bytes < bytesA || dB->bytes < bytesB || dC->bytes < bytesC) return fail(ctx, B200MM_ERR_INVALID, "mm_host: staging buffer too small");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->s_h2d) {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    }
    static const bool host_trace = getenv("B200MM_HOST_TRACE") != nullptr;  // debug: print the timeline of the pipelined call
    timespec ts_entry{}, ts_enq{}, ts_done{};
    if (host_trace) clock_gettime(CLOCK_MONOTONIC, &ts_entry);
    while ((int)ctx->pipe_ev.size() < 3 * 16 + 2) {
        cudaEvent_t e;
        CU_TRY(ctx, cudaEventCreateWithFlags(&e, host_trace ? cudaEventDefault : cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    if (!kern->panel || kern->panel_count != P) {
        if (kern->panel) b200mm_kernel_free(ctx, kern->panel);
        kern->panel = nullptr;
        b200mm_kernel_params prm = kern->prm;
        // the panels share one B: its lo part is computed once (panel 0) and reused, so the pre-pass form is the cheaper one here
        if (kern->id == B200MM_K_SGEMM_TC3X && prm.tune[3] == 0) prm.tune[3] = 2;
        if ((rc = b200mm_kernel_get(ctx, kern->id, M / P, N, K, &prm, &kern->panel))) return rc;
        kern->panel_count = P;
    }
    b200mm_kernel* pk = kern->panel;
    const size_t Mp = M / P;
    // Experiment (off by default): B200MM_HOST_ZEROCOPY_C=1 lets the epilogue store C straight into the caller's pinned host buffer
    // over PCIe instead of staging it in dC and copying panel by panel.  Needs hostC to be device-accessible (b200mm_host_alloc).
    bool zero_copy_c = false;
    if (const char* z = getenv("B200MM_HOST_ZEROCOPY_C")) {
        if (atoi(z) == 1) {
            cudaPointerAttributes pa{};
            zero_copy_c = cudaPointerGetAttributes(&pa, hostC) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr &&
                          (((uintptr_t)pa.devicePointer) & 15) == 0;
            cudaGetLastError();
            if (zero_copy_c) hostC = pa.devicePointer;  // same address under UVA
        }
    }
    cudaEvent_t ev_start = ctx->pipe_ev[0], ev_b = ctx->pipe_ev[1];
    // the copy streams must not run ahead of work already queued on the compute stream (it may still use dA/dB/dC)
    CU_TRY(ctx, cudaEventRecord(ev_start, ctx->stream));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev_start, 0));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_d2h, ev_start, 0));
    CU_TRY(ctx, cudaMemcpyAsync(dB->ptr, hostB, bytesB, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU_TRY(ctx, cudaEventRecord(ev_b, ctx->s_h2d));
    for (int i = 0; i < P; ++i) {
        CU_TRY(ctx, cudaMemcpyAsync((char*)dA->ptr + (size_t)i * Mp * K * 4, (const char*)hostA + (size_t)i * Mp * K * 4, Mp * K * 4,
                                    cudaMemcpyHostToDevice, ctx->s_h2d));
        CU_TRY(ctx, cudaEventRecord(ctx->pipe_ev[2 + i], ctx->s_h2d));
    }
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ev_b, 0));
    for (int i = 0; i < P; ++i) {
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 + i], 0));
        pk->tc_skip_b_split = (i > 0);  // B_lo from panel 0 is still in the panel kernel's workspace
        rc = b200mm_launch_ptr(ctx, pk, (const char*)dA->ptr + (size_t)i * Mp * K * 4, dB->ptr,
                               (zero_copy_c ? (char*)hostC : (char*)dC->ptr) + (size_t)i * Mp * N * 4, nullptr);
        pk->tc_skip_b_split = false;
        if (rc) return rc;
        if (zero_copy_c) continue;  // C is already on its way to host memory; the final stream synchronize makes it visible
        CU_TRY(ctx, cudaEventRecord(ctx->pipe_ev[2 + 16 + i], ctx->stream));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->pipe_ev[2 + 16 + i], 0));
        CU_TRY(ctx, cudaMemcpyAsync((char*)hostC + (size_t)i * Mp * N * 4, (const char*)dC->ptr + (size_t)i * Mp * N * 4, Mp * N * 4,
                                    cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (host_trace) CU_TRY(ctx, cudaEventRecord(ctx->pipe_ev[2 + 32 + i], ctx->s_d2h));
    }
    if (host_trace) clock_gettime(CLOCK_MONOTONIC, &ts_enq);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->s_d2h));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (host_trace) clock_gettime(CLOCK_MONOTONIC, &ts_done);
    if (host_trace && !zero_copy_c) {
        auto ms_between = [](const timespec& a, const timespec& b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6; };
        fprintf(stderr, "[b200mm] mm_host host side: everything enqueued after %.3f ms, returned after %.3f ms\n", ms_between(ts_entry, ts_enq),
                ms_between(ts_entry, ts_done));
        auto at = [&](cudaEvent_t e) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev_start, e);
            return ms;
        };
        fprintf(stderr, "[b200mm] mm_host %zux%zux%zu, %d panels: B up %.3f ms |", M, N, K, P, at(ev_b));
        for (int i = 0; i < P; ++i)
            fprintf(stderr, " p%d: A up %.3f gemm done %.3f C down %.3f |", i, at(ctx->pipe_ev[2 + i]), at(ctx->pipe_ev[2 + 16 + i]), at(ctx->pipe_ev[2 + 32 + i]));
        fprintf(stderr, "\n");
    }
    return B200MM_OK;
}

// ------------------------------------------------------------------------------------------------
// timing / cache control
// ------------------------------------------------------------------------------------------------
extern "C" int b200mm_timer_begin(b200mm_ctx* ctx) {
    if (!ctx) return fail(nullptr, B200MM_ERR_INVALID, "timer_begin: ctx is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return B200MM_OK;
}

extern "C" int b200mm_timer_end(b200mm_ctx* ctx, float* elapsed_ms) {
    if (!ctx || !elapsed_ms) return fail(ctx, B200MM_ERR_INVALID, "timer_end: NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    CU_TRY(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
    return B200MM_OK;
}

extern "C" int b200mm_flush_l2(b200mm_ctx* ctx) {
    if (!ctx) return fail(nullptr, B200MM_ERR_INVALID, "flush_l2: ctx is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->flush_buf) {
        ctx->flush_bytes = std::max<size_t>((size_t)ctx->prop.l2CacheSize * 2, (size_t)256 << 20);
        CU_TRY(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    flush_kernel<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>((float4*)ctx->flush_buf, ctx->flush_bytes / 16, 0.f);
    CU_TRY(ctx, cudaGetLastError());
    return B200MM_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU helpers
// ------------------------------------------------------------------------------------------------
extern "C" int b200mm_ipc_export(b200mm_ctx* ctx, const b200mm_buffer* buf, void* handle64) {
    if (!ctx || !buf || !handle64) return fail(ctx, B200MM_ERR_INVALID, "ipc_export: NULL argument");
    if (!buf->owned) return fail(ctx, B200MM_ERR_INVALID, "ipc_export: only library-allocated buffers can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU_TRY(ctx, cudaIpcGetMemHandle(&h, buf->ptr));
    memcpy(handle64, &h, 64);
    return B200MM_OK;
}

extern "C" int b200mm_ipc_import(b200mm_ctx* ctx, const void* handle64, size_t bytes, b200mm_buffer** out) {
    if (!ctx || !handle64 || !out) return fail(ctx, B200MM_ERR_INVALID, "ipc_import: NULL argument");
    *out = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    b200mm_buffer* b = new b200mm_buffer();
    b->ptr = p;
    b->bytes = bytes;
    b->ipc = true;
    *out = b;
    return B200MM_OK;
}

extern "C" int b200mm_peer_barrier(b200mm_ctx* ctx, b200mm_buffer* local_flags, void* const* peer_flags, int rank, int world) {
    if (!ctx || !local_flags || !peer_flags) return fail(ctx, B200MM_ERR_INVALID, "peer_barrier: NULL argument");
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(ctx, B200MM_ERR_INVALID, "peer_barrier: bad rank/world");
    if (local_flags->bytes < (size_t)world * 4) return fail(ctx, B200MM_ERR_INVALID, "peer_barrier: flags buffer too small");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    PeerFlags pf{};
    for (int i = 0; i < world; ++i) {
        if (!peer_flags[i]) return fail(ctx, B200MM_ERR_INVALID, "peer_barrier: peer %d is NULL", i);
        pf.f[i] = (unsigned int*)peer_flags[i];
    }
    const unsigned int epoch = ++local_flags->barrier_epoch;
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(pf, (unsigned int*)local_flags->ptr, rank, world, epoch);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200MM_OK;
}

extern "C" int b200mm_unshard_columns(b200mm_ctx* ctx, const void* gathered, void* C, size_t M, size_t N, int world) {
    if (!ctx || !gathered || !C) return fail(ctx, B200MM_ERR_INVALID, "unshard_columns: NULL argument");
    if (world <= 0 || N % ((size_t)4 * world)) return fail(ctx, B200MM_ERR_INVALID, "unshard_columns: N must be a multiple of 4*world");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    unshard_columns_kernel<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>((const float4*)gathered, (float4*)C, M, N / 4, world);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200MM_OK;
}

// ------------------------------------------------------------------------------------------------
// debug: tcgen05 bring-up probe (tools/probe_tc.py); not part of the public header
// ------------------------------------------------------------------------------------------------
// ---- device-free introspection of the tc3x schedule (tests/test_host.py) ---------------------------------------------------
// bn == 512 selects the schedule of the 2-CTA kernel: 256 x 256 tiles over sms / 2 CTA pairs (out[0] = number of pairs)
static Tc3xSchedule tc3x_sched_for(size_t M, size_t N, size_t K, int bn, int bk, int sms, bool pure) {
    return bn == 512 ? tc3x_make_schedule(M, N, K, 256, bk, sms / 2, pure, 256) : tc3x_make_schedule(M, N, K, bn, bk, sms, pure);
}
extern "C" int b200mm_tc3x_plan(size_t M, size_t N, size_t K, int sms, const uint32_t tune[4], uint32_t flags, int out[9]) {
    if (!out || !M || !N || !K || sms <= 0) return B200MM_ERR_INVALID;
    static const uint32_t zero[4] = {0, 0, 0, 0};
    if (!tune) tune = zero;
    if (N % 4 || K % 4) {  // the padded path plans for the padded shape
        N = ceil_div(N, 4) * 4;
        K = ceil_div(K, 4) * 4;
    }
    const Tc3xPlan p = tc3x_plan(M, N, K, sms, tune, (flags & B200MM_F_TC3X_1X) != 0);
    const Tc3xSchedule sc = tc3x_make_schedule(M, N, K, p.bn, p.bk, p.cta2 ? sms / 2 : sms, tune[1] == 1, p.cta2 ? 256 : 128);
    out[0] = p.bn;
    out[1] = p.bk;
    out[2] = p.cta2 ? 1 : 0;
    out[3] = p.tma_store ? 1 : 0;
    out[4] = p.split;
    out[5] = p.a_prepass ? 1 : 0;
    out[6] = sc.grid * (p.cta2 ? 2 : 1);
    out[7] = sc.full_waves;
    out[8] = sc.k_split;
    return B200MM_OK;
}

extern "C" int b200mm_tc3x_schedule(size_t M, size_t N, size_t K, int bn, int bk, int sms, int pure_stream_k, int out[6]) {
    if (!out || !M || !N || !K || (bn != 128 && bn != 256 && bn != 512) || (bk != 16 && bk != 32) || sms <= 0) return B200MM_ERR_INVALID;
    const Tc3xSchedule sc = tc3x_sched_for(M, N, K, bn, bk, sms, pure_stream_k != 0);
    out[0] = sc.grid;
    out[1] = sc.full_waves;
    out[2] = sc.chains_per_tile;
    out[3] = sc.k_split;
    out[4] = (int)sc.tiles;
    out[5] = (int)sc.sk_units;
    return B200MM_OK;
}

extern "C" int b200mm_tc3x_schedule_cover(size_t M, size_t N, size_t K, int bn, int bk, int sms, int pure_stream_k, uint16_t* cover,
                                          size_t cover_len, int* max_segments_per_cta, int* max_chains_per_cta) {
    if (!cover || !M || !N || !K || (bn != 128 && bn != 256 && bn != 512) || (bk != 16 && bk != 32) || sms <= 0) return B200MM_ERR_INVALID;
    const Tc3xSchedule sc = tc3x_sched_for(M, N, K, bn, bk, sms, pure_stream_k != 0);
    if (cover_len < (size_t)sc.tiles * sc.chains_per_tile) return B200MM_ERR_INVALID;
    int max_seg = 0, max_ch = 0;
    for (int b = 0; b < sc.grid; ++b) {
        SegIter it(sc.chains_per_tile, sc.full_waves, sc.sk_units, b, sc.grid);  // the very iterator the device roles run
        int tile, c0, c1, skt, seg = 0, ch = 0;
        while (it.next(tile, c0, c1, skt)) {
            if (tile < 0 || tile >= sc.tiles || c0 < 0 || c1 > sc.chains_per_tile || c0 >= c1) return B200MM_ERR_INVALID;
            for (int c = c0; c < c1; ++c) cover[(size_t)tile * sc.chains_per_tile + c]++;
            ++seg;
            ch += c1 - c0;
        }
        max_seg = std::max(max_seg, seg);
        max_ch = std::max(max_ch, ch);
    }
    if (max_segments_per_cta) *max_segments_per_cta = max_seg;
    if (max_chains_per_cta) *max_chains_per_cta = max_ch;
    return B200MM_OK;
}

// Replays the stream-K fix-up protocol on the host with the kernel's own iterator and contributor rule: every finisher
// must wait only on LOWER-numbered CTAs that really park a part of the same tile, and the parts must cover the tile
// exactly.  Returns B200MM_OK and *violations == 0 when the protocol is consistent (no wait on a CTA that never publishes).
extern "C" int b200mm_tc3x_schedule_replay(size_t M, size_t N, size_t K, int bn, int bk, int sms, int pure_stream_k, int* violations,
                                           int* max_wait_list) {
    if (!violations || !M || !N || !K || (bn != 128 && bn != 256 && bn != 512) || (bk != 16 && bk != 32) || sms <= 0) return B200MM_ERR_INVALID;
    const Tc3xSchedule sc = tc3x_sched_for(M, N, K, bn, bk, sms, pure_stream_k != 0);
    struct Parked { int tile, c0, c1; };
    std::vector<Parked> parked(sc.grid, Parked{-1, 0, 0});
    std::vector<int> publishes(sc.grid, 0);
    for (int b = 0; b < sc.grid; ++b) {
        SegIter it(sc.chains_per_tile, sc.full_waves, sc.sk_units, b, sc.grid);
        int tile, c0, c1, skt;
        while (it.next(tile, c0, c1, skt))
            if (c1 != sc.chains_per_tile) {
                parked[b] = Parked{tile, c0, c1};
                publishes[b]++;
            }
    }
    int bad = 0, max_wait = 0;
    for (int b = 0; b < sc.grid; ++b) {
        if (publishes[b] > 1) ++bad;  // one workspace slot / flag per CTA
        SegIter it(sc.chains_per_tile, sc.full_waves, sc.sk_units, b, sc.grid);
        int tile, c0, c1, skt;
        // order inside a CTA: the parked segment must be the FIRST phase-2 segment and a finisher segment the LAST one,
        // otherwise the CTAs serialise (each finisher waiting for a part its predecessor parks at the end of its range)
        int nseg2 = 0, park_at = -1, finish_at = -1;
        {
            SegIter it2(sc.chains_per_tile, sc.full_waves, sc.sk_units, b, sc.grid);
            while (it2.next(tile, c0, c1, skt)) {
                if (skt < 0) continue;
                if (c1 != sc.chains_per_tile) park_at = nseg2;
                else if (c0 != 0) finish_at = nseg2;
                ++nseg2;
            }
            if (park_at > 0) ++bad;
            if (finish_at >= 0 && finish_at != nseg2 - 1) ++bad;
        }
        while (it.next(tile, c0, c1, skt)) {
            if (c1 != sc.chains_per_tile || c0 == 0) continue;
            // finisher: the exact loop of the epilogue warps
            int next_chain = 0, waits = 0;
            const int j0 = tc3x_first_contributor(skt, sc.chains_per_tile, sc.sk_units, b, sc.grid);
            for (int j = j0; j < b; ++j) {
                if (tc3x_cta_is_empty(j, sc.sk_units, sc.grid)) continue;
                ++waits;
                if (j >= b || publishes[j] != 1 || parked[j].tile != tile || parked[j].c0 != next_chain) {
                    ++bad;
                    continue;
                }
                next_chain = parked[j].c1;
            }
            if (next_chain != c0) ++bad;  // the parts and the finisher's own chains must tile [0, cpt)
            max_wait = std::max(max_wait, waits);
        }
    }
    *violations = bad;
    if (max_wait_list) *max_wait_list = max_wait;
    return B200MM_OK;
}

// Measurement tool (bench.py "measured_peaks"): FP32 FMA-pipe throughput of the device in TFLOP/s, packed (FFMA2) or scalar
// FFMA, best of `reps` timed launches of a register-only kernel (2 x 1024 threads per SM, 8 independent chains per thread).
extern "C" int b200mm_measure_fma_peak(b200mm_ctx* ctx, int packed, int iters, int reps, double* tflops_out) {
    if (!ctx || !tflops_out || iters <= 0 || reps <= 0) return fail(ctx, B200MM_ERR_INVALID, "measure_fma_peak: bad argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    float* sink = nullptr;
    CU_TRY(ctx, cudaMalloc(&sink, 256));
    const int blocks = ctx->prop.multiProcessorCount * 2;
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {  // first launch warms up
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        if (packed)
            fma_peak_kernel<true><<<blocks, 1024, 0, ctx->stream>>>(sink, iters, 0.999f, 1e-3f);
        else
            fma_peak_kernel<false><<<blocks, 1024, 0, ctx->stream>>>(sink, iters, 0.999f, 1e-3f);
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU_TRY(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (r) best = std::min(best, ms);
    }
    cudaFree(sink);
    const double flop = (double)blocks * 1024.0 * iters * 4.0 * 8.0 * 2.0 /*x and y*/ * 2.0 /*fma = 2 flop*/;
    *tflops_out = flop / (best * 1e-3) / 1e12;
    return B200MM_OK;
}

// Measurement tool (tools/peer_latency.py): flag ping-pong between two ranks, see peer_pingpong_kernel.  local_flags /
// peer_flags: one u32 each (IPC-mapped), peer_payload: >= payload floats on the peer.  Returns ns per round trip.
extern "C" B200MM_API int b200mm_debug_peer_pingpong(b200mm_ctx* ctx, void* local_flag, void* peer_flag, void* peer_payload, int payload,
                                                     int rank, int iters, int mode, unsigned int base, double* ns_per_round) {
    if (!ctx || !local_flag || !peer_flag || !ns_per_round || iters <= 0) return fail(ctx, B200MM_ERR_INVALID, "peer_pingpong: bad argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned long long* out = nullptr;
    CU_TRY(ctx, cudaMalloc(&out, 8));
    peer_pingpong_kernel<<<1, 256, 0, ctx->stream>>>((unsigned int*)local_flag, (unsigned int*)peer_flag, (float*)peer_payload, payload, rank,
                                                     iters, mode, base, out);
    CU_TRY(ctx, cudaGetLastError());
    unsigned long long ns = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&ns, out, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(out);
    *ns_per_round = (double)ns / iters;
    return B200MM_OK;
}

// Debug (tools/trace_gemv.py): per-CTA globaltimer stamps of the next `slots` launches of a GEMV kernel object into
// `buf` ([slots][CTAs][8] u64, device memory); slots == 0 switches it off.
extern "C" B200MM_API int b200mm_debug_gemv_trace(b200mm_kernel* k, void* buf, int slots) {
    if (!k) return B200MM_ERR_INVALID;
    k->trace_buf = slots > 0 ? (unsigned long long*)buf : nullptr;
    k->trace_slots = slots;
    k->trace_next = 0;
    return B200MM_OK;
}

extern "C" B200MM_API int b200mm_debug_tc_probe(b200mm_ctx* ctx, const void* A, const void* B, size_t M, size_t N, size_t K,
                                                const uint32_t* u32args /*11*/, void* dumpA, void* dumpB, void* dumpD) {
    if (!ctx || !A || !B) return fail(ctx, B200MM_ERR_INVALID, "probe: NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CUtensorMap tmA, tmB;
    int rc;
    if ((rc = make_tmap_kmajor(ctx, &tmA, (const float*)A, M, K, 128))) return rc;
    if ((rc = make_tmap_mnmajor(ctx, &tmB, (const float*)B, K, N, 32, 256, (CUtensorMapSwizzle)u32args[10]))) return rc;
    ProbeArgs pa{};
    pa.idesc = u32args[0];
    pa.a_lbo = u32args[1]; pa.a_sbo = u32args[2]; pa.a_kstep = u32args[3];
    pa.b_lbo = u32args[4]; pa.b_sbo = u32args[5]; pa.b_kstep = u32args[6];
    pa.layout = u32args[7]; pa.nk = u32args[8]; pa.mode = u32args[9];
    pa.dumpA = (float*)dumpA; pa.dumpB = (float*)dumpB; pa.dumpD = (float*)dumpD;
    const int smem = 128 * 32 * 4 + 32 * 256 * 4 + 1024 + 256;
    CU_TRY(ctx, cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_probe_kernel<<<1, 256, smem, ctx->stream>>>(tmA, tmB, pa);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200MM_OK;
}
