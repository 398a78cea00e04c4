// Device-side synthetic data, bit-identical to oracle_generate_weight_data_at (oracle/oracle.c):
// U[-10,10)/50 as in the reference's generate_weight_data (src/harness.rs:103-121), but seeded and
// counter-based so that 16384^2 operands can be produced in place on each GPU.
#pragma once
#include "common.cuh"

namespace b200mm {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ float weight_value(uint64_t seed, uint64_t i) {
    uint32_t u24 = (uint32_t)(splitmix64(seed * 0xD1342543DE82EF95ull + i) >> 40);
    float f = __fmul_rn((float)u24, 1.0f / 16777216.0f);
    float x = __fadd_rn(__fmul_rn(f, 20.0f), -10.0f);  // no FMA contraction: must match the CPU oracle bit for bit
    return __fdiv_rn(x, 50.0f);
}

__global__ void fill_weights_kernel(float* __restrict__ out, uint64_t seed, uint64_t offset, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = weight_value(seed, offset + i);
}

__global__ void fill_weights_2d_kernel(float* __restrict__ out, uint64_t seed, uint64_t offset, size_t rows, size_t cols,
                                       size_t src_ld, size_t src_col0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x, n = rows * cols;
    for (; i < n; i += stride) {
        const size_t r = i / cols, c = i % cols;
        out[i] = weight_value(seed, offset + r * src_ld + src_col0 + c);
    }
}

// [world][M][N/world] -> row-major M x N (after an all-gather of column panels).
__global__ void unshard_columns_kernel(const float4* __restrict__ gathered, float4* __restrict__ C, size_t M,
                                       size_t n4, int world) {
    const size_t p4 = n4 / world;  // float4 per panel row
    const size_t total = M * n4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        size_t m = i / n4, c = i % n4;
        size_t r = c / p4, cc = c % p4;
        C[i] = gathered[(r * M + m) * p4 + cc];
    }
}

// Cross-GPU barrier over peer-mapped flag words (b200mm_peer_barrier).  Thread d publishes this rank's arrival to
// rank d (system-scope release: everything this GPU wrote before, including stores to peers, is visible first), then
// waits for rank d's arrival in the local flags.
struct PeerFlags {
    unsigned int* f[8];
};
__global__ void peer_barrier_kernel(PeerFlags peers, unsigned int* local, int rank, int world, unsigned int epoch) {
    const int d = threadIdx.x;
    if (d >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.f[d] + rank), "r"(epoch) : "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    unsigned int seen;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local + d) : "memory");
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) __trap();  // 10 s: a peer never arrived -- fail loudly instead of hanging the GPU
    } while ((int)(seen - epoch) < 0);
}

// > L2-sized write used by b200mm_flush_l2.
__global__ void flush_kernel(float4* __restrict__ p, size_t n4, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) p[i] = make_float4(v, v, v, v);
}

}  // namespace b200mm
