// CUDA re-expressions of the reference's twelve WGSL kernels.
//
// These are NOT the fast path.  They exist so that (i) the reference's own test list
// (test_gemm_1 .. test_gemm_5, test_qgemv_1; src/gemm.rs:172-177, src/gemv.rs:41-49) runs unchanged
// through the new harness with the Workload's WorkgroupCount / WorkgroupSize used as gridDim /
// blockDim, and (ii) "agreement with the WGSL output" is testable per shader: every kernel keeps the
// shader's work split per invocation and its per-output accumulation order.  Where the shader writes
// `acc += a * b` the port uses separately rounded multiply and add (__fmul_rn/__fadd_rn) and where it
// calls fma() the port calls fmaf, so each port is bit-identical to its CPU restatement in
// oracle/oracle.c (WGSL itself leaves contraction to the back-end, so that is a choice, not a fact
// about the reference; see DESIGN.md "Parity").
//
// WGSL builtins map as: global_invocation_id = blockIdx*blockDim+threadIdx, local_invocation_id =
// threadIdx, workgroup_id = blockIdx.  M, N, K are runtime arguments here (Tera literals there).
#pragma once
#include "common.cuh"

namespace b200mm {
namespace wgsl {

__device__ __forceinline__ float madd(float a, float b, float c) { return __fadd_rn(c, __fmul_rn(a, b)); }  // c + a*b, unfused

// shaders/gemm/gemm_1.wgsl:12-28
__global__ void gemm_1(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, unsigned M,
                       unsigned N, unsigned K) {
    const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;  // row: the FAST thread index (uncoalesced on purpose)
    const unsigned y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < M && y < N) {
        float tmp = 0.f;
        for (unsigned i = 0; i < K; ++i) tmp = madd(A[(size_t)x * K + i], B[(size_t)i * N + y], tmp);
        C[(size_t)x * N + y] = tmp;
    }
}

// shaders/gemm/gemm_1v.wgsl:12-32 (buffers viewed as vec4)
__global__ void gemm_1v(const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, unsigned M,
                        unsigned N, unsigned K) {
    const unsigned cRow = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cCol = blockIdx.y * blockDim.y + threadIdx.y;
    if (cRow < M && cCol < N / 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned k = 0; k < K / 4; ++k) {
            const float4 a = A[(size_t)cRow * K / 4 + k];
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (unsigned c = 0; c < 4; ++c) {
                const float4 b = B[(size_t)k * N + cCol + (size_t)c * N / 4];
                t.x = madd(av[c], b.x, t.x);
                t.y = madd(av[c], b.y, t.y);
                t.z = madd(av[c], b.z, t.z);
                t.w = madd(av[c], b.w, t.w);
            }
        }
        C[(size_t)cRow * N / 4 + cCol] = t;
    }
}

// shaders/gemm/gemm_2.wgsl:12-30
__global__ void gemm_2(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, unsigned M,
                       unsigned N, unsigned K) {
    const unsigned cRow = blockIdx.x * 16u + threadIdx.x / 16u;
    const unsigned cCol = blockIdx.y * 16u + threadIdx.x % 16u;
    if (cRow < M && cCol < N) {
        float tmp = 0.f;
        for (unsigned i = 0; i < K; ++i) tmp = madd(A[(size_t)cRow * K + i], B[(size_t)i * N + cCol], tmp);
        C[(size_t)cRow * N + cCol] = tmp;
    }
}

// shaders/gemm/gemm_3.wgsl:12-49 (BLOCKSIZE = 16, src/gemm.rs:72)
__global__ void gemm_3(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, unsigned M,
                       unsigned N, unsigned K) {
    constexpr unsigned BS = 16;
    __shared__ float As[BS * BS], Bs[BS * BS];
    const unsigned cRow = blockIdx.x, cCol = blockIdx.y;
    const unsigned tc = threadIdx.x % BS, tr = threadIdx.x / BS;
    size_t a = (size_t)cRow * BS * K, b = (size_t)cCol * BS;
    const size_t c = (size_t)cRow * BS * N + (size_t)cCol * BS;
    float tmp = 0.f;
    for (unsigned bk = 0; bk < K; bk += BS) {
        As[tr * BS + tc] = A[a + (size_t)tr * K + tc];
        Bs[tr * BS + tc] = B[b + (size_t)tr * N + tc];
        __syncthreads();
        a += BS;
        b += (size_t)BS * N;
        for (unsigned d = 0; d < BS; ++d) tmp = fmaf(As[tr * BS + d], Bs[d * BS + tc], tmp);
        __syncthreads();
    }
    C[c + (size_t)tr * N + tc] = tmp;
}

// shaders/gemm/gemm_4.wgsl:15-63 (BM=BN=16, BK=8, TM=2; src/gemm.rs:95-98).  grid.x <-> N, grid.y <-> M.
__global__ void gemm_4(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, unsigned M,
                       unsigned N, unsigned K) {
    constexpr unsigned BM = 16, BN = 16, BK = 8, TM = 2;
    __shared__ float As[BM * BK], Bs[BK * BN];
    const unsigned cRow = blockIdx.y, cCol = blockIdx.x;
    const unsigned tc = threadIdx.x % BN, tr = threadIdx.x / BN;
    size_t a = (size_t)cRow * BM * K, b = (size_t)cCol * BN;
    const size_t c = (size_t)cRow * BM * N + (size_t)cCol * BN;
    const unsigned icA = threadIdx.x % BK, irA = threadIdx.x / BK, icB = threadIdx.x % BN, irB = threadIdx.x / BN;
    float res[TM] = {0.f, 0.f};
    for (unsigned bk = 0; bk < K; bk += BK) {
        As[irA * BK + icA] = A[a + (size_t)irA * K + icA];
        Bs[irB * BN + icB] = B[b + (size_t)irB * N + icB];
        __syncthreads();
        a += BK;
        b += (size_t)BK * N;
        for (unsigned d = 0; d < BK; ++d) {
            const float tb = Bs[d * BN + tc];
#pragma unroll
            for (unsigned r = 0; r < TM; ++r) res[r] = fmaf(As[(tr * TM + r) * BK + d], tb, res[r]);
        }
        __syncthreads();
    }
#pragma unroll
    for (unsigned r = 0; r < TM; ++r) C[c + (size_t)(tr * TM + r) * N + tc] = res[r];
}

// shaders/gemm/gemm_5.wgsl:15-86 (BM=BN=32, BK=16, TM=TN=4; src/gemm.rs:124-128)
__global__ void gemm_5(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, unsigned M,
                       unsigned N, unsigned K) {
    constexpr unsigned BM = 32, BN = 32, BK = 16, TM = 4, TN = 4, NT = BM * BN / (TM * TN);
    __shared__ float As[BM * BK], Bs[BK * BN];
    const unsigned cRow = blockIdx.y, cCol = blockIdx.x;
    const unsigned tc = threadIdx.x % (BN / TN), tr = threadIdx.x / (BN / TN);
    size_t a = (size_t)cRow * BM * K, b = (size_t)cCol * BN;
    const size_t c = (size_t)cRow * BM * N + (size_t)cCol * BN;
    const unsigned icA = threadIdx.x % BK, irA = threadIdx.x / BK, icB = threadIdx.x % BN, irB = threadIdx.x / BN;
    constexpr unsigned strideA = NT / BK, strideB = NT / BN;
    float res[TM * TN], regM[TM], regN[TN];
#pragma unroll
    for (unsigned i = 0; i < TM * TN; ++i) res[i] = 0.f;
    for (unsigned bk = 0; bk < K; bk += BK) {
        for (unsigned lo = 0; lo < BM; lo += strideA) As[(irA + lo) * BK + icA] = A[a + (size_t)(irA + lo) * K + icA];
        for (unsigned lo = 0; lo < BK; lo += strideB) Bs[(irB + lo) * BN + icB] = B[b + (size_t)(irB + lo) * N + icB];
        __syncthreads();
        a += BK;
        b += (size_t)BK * N;
        for (unsigned d = 0; d < BK; ++d) {
#pragma unroll
            for (unsigned i = 0; i < TM; ++i) regM[i] = As[(tr * TM + i) * BK + d];
#pragma unroll
            for (unsigned i = 0; i < TN; ++i) regN[i] = Bs[d * BN + tc * TN + i];
#pragma unroll
            for (unsigned rm = 0; rm < TM; ++rm)
#pragma unroll
                for (unsigned rn = 0; rn < TN; ++rn) res[rm * TN + rn] = fmaf(regM[rm], regN[rn], res[rm * TN + rn]);
        }
        __syncthreads();
    }
#pragma unroll
    for (unsigned rm = 0; rm < TM; ++rm)
#pragma unroll
        for (unsigned rn = 0; rn < TN; ++rn) C[c + (size_t)(tr * TM + rm) * N + tc * TN + rn] = res[rm * TN + rn];
}

// shaders/gemm.wgsl:11-14 + shaders/gemm_macro.wgsl:2-53 (WONNX).  1-D grid, M*N/16 invocations; a 4x4
// block per invocation; product = mat_right * mat_left summed over 4 k first, then added to result.
__global__ void gemm_wonnx(const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, unsigned M,
                           unsigned N, unsigned K) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)M * N / 16) return;  // WGSL relies on the dispatch being exact; guard for ragged grids
    const unsigned y = gid % (N / 4), x = gid / (N / 4);
    const size_t index = (size_t)x * N + y;
    float4 result[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) result[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned k = 0; k < K / 4; ++k) {
        float4 L[4], R[4];
#pragma unroll
        for (unsigned i = 0; i < 4; ++i) {
            L[i] = A[(size_t)x * K + k + (size_t)i * K / 4];
            R[i] = B[(size_t)k * N + y + (size_t)i * N / 4];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float l[4] = {L[j].x, L[j].y, L[j].z, L[j].w};
            float4 pr = make_float4(__fmul_rn(R[0].x, l[0]), __fmul_rn(R[0].y, l[0]), __fmul_rn(R[0].z, l[0]),
                                    __fmul_rn(R[0].w, l[0]));
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                pr.x = madd(R[i].x, l[i], pr.x);
                pr.y = madd(R[i].y, l[i], pr.y);
                pr.z = madd(R[i].z, l[i], pr.z);
                pr.w = madd(R[i].w, l[i], pr.w);
            }
            result[j].x = __fadd_rn(result[j].x, pr.x);
            result[j].y = __fadd_rn(result[j].y, pr.y);
            result[j].z = __fadd_rn(result[j].z, pr.z);
            result[j].w = __fadd_rn(result[j].w, pr.w);
        }
    }
#pragma unroll
    for (unsigned j = 0; j < 4; ++j) C[index + (size_t)j * N / 4] = result[j];
}

// shaders/bram.wgsl:11-50 and shaders/bram8x8.wgsl:10-50 (same body; the literal 256u is K/4 = N/4 at 1024^3)
__global__ void bram(const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, unsigned M,
                     unsigned N, unsigned K) {
    const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = blockIdx.y * blockDim.y + threadIdx.y;
    if (m >= M / 4 || n >= N / 4) return;
    const unsigned k4 = K / 4, n4 = N / 4;
    float4 r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned k = 0; k < k4; ++k) {
        float4 a[4], b[4];
#pragma unroll
        for (unsigned i = 0; i < 4; ++i) {
            a[i] = A[(size_t)(m * 4 + i) * k4 + k];
            b[i] = B[(size_t)(k * 4 + i) * n4 + n];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float s = c == 0 ? a[i].x : c == 1 ? a[i].y : c == 2 ? a[i].z : a[i].w;
                r[i].x = madd(s, b[c].x, r[i].x);
                r[i].y = madd(s, b[c].y, r[i].y);
                r[i].z = madd(s, b[c].z, r[i].z);
                r[i].w = madd(s, b[c].w, r[i].w);
            }
    }
#pragma unroll
    for (unsigned i = 0; i < 4; ++i) C[(size_t)(m * 4 + i) * n4 + n] = r[i];
}

// shaders/gemm3.wgsl:11-92 (webgpu-blas): x = gid.x over N/8, y = gid.y over M/4; 4 rows x 2 vec4 per invocation
__global__ void gemm3(const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, unsigned M,
                      unsigned N, unsigned K) {
    const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= N / 8 || y >= M / 4) return;
    const unsigned KD4 = K / 4, ND4 = N / 4;
    float4 r[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) r[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned k = 0; k < KD4; ++k) {
        float4 ar[4];
#pragma unroll
        for (unsigned i = 0; i < 4; ++i) ar[i] = A[(size_t)(y * 4 + i) * KD4 + k];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 br = B[(size_t)(k * 4 + c) * ND4 + x * 2 + h];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float s = c == 0 ? ar[i].x : c == 1 ? ar[i].y : c == 2 ? ar[i].z : ar[i].w;
                    r[h][i].x = madd(s, br.x, r[h][i].x);
                    r[h][i].y = madd(s, br.y, r[h][i].y);
                    r[h][i].z = madd(s, br.z, r[h][i].z);
                    r[h][i].w = madd(s, br.w, r[h][i].w);
                }
            }
    }
#pragma unroll
    for (unsigned h = 0; h < 2; ++h)
#pragma unroll
        for (unsigned i = 0; i < 4; ++i) C[x * 2 + h + (size_t)(y * 4 + i) * ND4] = r[h][i];
}

__device__ __forceinline__ float4 unpack4x8snorm_scaled(uint32_t w, float absmax) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = (int)(signed char)((w >> (8 * i)) & 0xFFu);
        v[i] = __fmul_rn(fmaxf(__fdiv_rn((float)q, 127.0f), -1.0f), absmax);
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

// shaders/gemv/qgemv_1.wgsl:10-39; gid.y is the batch (offsets :12-14)
// The shader has no bounds checks (wgpu's robust buffer access absorbs stray invocations); here invocations beyond the
// N/4 outputs or the batch count return, so a workgroup_size_y that does not divide the batch cannot touch memory past the buffers.
__global__ void qgemv_1(const float4* __restrict__ A, const uint32_t* __restrict__ B, float4* __restrict__ C, unsigned N,
                        unsigned K, float absmax, unsigned batch) {
    const unsigned gx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= N / 4 || gy >= batch) return;
    const size_t left_offset = (size_t)gy * (K / 4), right_offset = (size_t)gy * ((size_t)K * N / 4),
                 output_offset = (size_t)gy * (N / 4);
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    for (unsigned k = 0; k < K / 4; ++k) {
        const float4 left = A[left_offset + k];
        const size_t index_right = right_offset + gx + (size_t)k * N;
        float4 rt[4];
#pragma unroll
        for (unsigned i = 0; i < 4; ++i) rt[i] = unpack4x8snorm_scaled(B[index_right + (size_t)i * (N / 4)], absmax);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            auto comp = [&](const float4& v) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; };
            float d = __fmul_rn(left.x, comp(rt[0]));
            d = madd(left.y, comp(rt[1]), d);
            d = madd(left.z, comp(rt[2]), d);
            d = madd(left.w, comp(rt[3]), d);
            res[c] = __fadd_rn(res[c], d);
        }
    }
    C[output_offset + gx] = make_float4(res[0], res[1], res[2], res[3]);
}

}  // namespace wgsl
}  // namespace b200mm
