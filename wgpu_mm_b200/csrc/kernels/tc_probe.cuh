// Bring-up probe for the tcgen05 path (debug only, not on the hot path): one CTA loads one k-block with
// the production tensor maps, dumps the shared-memory tiles, issues BK/8 MMAs with descriptor fields
// taken from runtime arguments, and dumps the TMEM accumulator.  Lets one GPU run test many descriptor
// hypotheses.
#pragma once
#include "sgemm_tc3x.cuh"

namespace b200mm {

struct ProbeArgs {
    uint32_t idesc;
    uint32_t a_lbo, a_sbo, a_kstep;  // bytes
    uint32_t b_lbo, b_sbo, b_kstep;  // bytes
    uint32_t layout;                 // descriptor layout types: low nibble A, high nibble B (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
    uint32_t nk;                     // number of K=8 MMAs to issue (1..4)
    uint32_t mode;                   // bit0: explicit disable-output-lane mask form; bit1: kind::f16; bit2: tcgen05.st pattern instead of MMA;
                                     // bit3: all 32 lanes of the warp reach the MMA code (elect.sync picks the issuer)
    float* dumpA;                    // 128*32 floats (raw smem image)
    float* dumpB;                    // 32*256 floats (raw smem image)
    float* dumpD;                    // 128*256 floats
};

__device__ __forceinline__ uint64_t probe_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout & 7u) << 61);
}

__global__ void __launch_bounds__(256, 1)
tc_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ProbeArgs p) {
    constexpr uint32_t A_BYTES = 128 * 32 * 4, B_BYTES = 32 * 256 * 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + A_BYTES, bar0 = base + A_BYTES + B_BYTES, bar1 = bar0 + 8, slot = bar0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar0, 1);
        ptx::mbar_init(bar1, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(slot, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(bar0, A_BYTES + B_BYTES);
        ptx::tma_load_2d(sA, &tmA, bar0, 0, 0);
        ptx::tma_load_3d(sB, &tmB, bar0, 0, 0, 0);
    }
    ptx::mbar_wait(bar0, 0);
    const float* gA = reinterpret_cast<const float*>(gen);
    const float* gB = reinterpret_cast<const float*>(gen + A_BYTES);
    for (int i = threadIdx.x; i < 128 * 32; i += 256) p.dumpA[i] = gA[i];
    for (int i = threadIdx.x; i < 32 * 256; i += 256) p.dumpB[i] = gB[i];
    __syncthreads();
    if (p.mode & 4u) {
        // TMEM write/read round trip without the tensor core: lane r, column c <- r*1000 + c
        if (warp >= 4) {
            const int q = warp & 3;
            for (int c = 0; c < 256; ++c) {
                const uint32_t val = __float_as_uint((float)((q * 32 + lane) * 1000 + c));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem_base + ((uint32_t)(q * 32) << 16) + c), "r"(val) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        ptx::tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 32) ptx::mbar_arrive(bar1);
    } else if ((p.mode & 8u) ? (warp == 1) : (threadIdx.x == 32)) {
        ptx::tc_fence_after();
        uint32_t elected = 1;
        if (p.mode & 8u) {
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        }
        if (elected) {
            for (uint32_t j = 0; j < p.nk; ++j) {
                const uint64_t ad = probe_desc(sA + j * p.a_kstep, p.a_lbo, p.a_sbo, p.layout & 15u);
                const uint64_t bd = (p.mode & 16u) ? probe_desc(sA + j * p.b_kstep, p.b_lbo, p.b_sbo, p.layout >> 4)   // B := A tile (K-major)
                                                   : probe_desc(sB + j * p.b_kstep, p.b_lbo, p.b_sbo, p.layout >> 4);
                const uint32_t acc = j > 0 ? 1u : 0u;
                if (p.mode & 2u) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base),
                                 "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc) : "memory");
                } else if (p.mode & 1u) {
                    uint32_t z = 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(
                                     tmem_base),
                                 "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
                } else {
                    ptx::mma_tf32_ss(tmem_base, ad, bd, p.idesc, acc);
                }
            }
            ptx::mma_commit(bar1);
        }
        if (p.mode & 8u) __syncwarp();
    }
    ptx::mbar_wait(bar1, 0);
    ptx::tc_fence_after();
    if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        for (int c = 0; c < 8; ++c) {
            float v[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) p.dumpD[row * 256 + c * 32 + j] = v[j];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 256);
    }
    if (threadIdx.x == 0) p.dumpD[128 * 256] = __uint_as_float(tmem_base);
}

}  // namespace b200mm
