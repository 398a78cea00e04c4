// Measurement tool, not a hot-path kernel: the FP32 FMA-pipe ceiling of this device, measured instead of derived
// (BASELINE.md section 2 asks for "a pure-FMA microbenchmark"; the SIMT SGEMM's roofline denominator).
// Every thread runs ILP independent dependency chains of packed (FFMA2, two IEEE fmas per instruction) or scalar FFMA
// for `iters` rounds; nothing touches memory until the final store that keeps the chains alive.
#pragma once
#include "common.cuh"

namespace b200mm {

template <bool PACKED>
__global__ void __launch_bounds__(1024, 2) fma_peak_kernel(float* __restrict__ sink, int iters, float a, float b) {
    constexpr int ILP = 8;
    float x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = (float)threadIdx.x * 1e-6f + (float)i;
        y[i] = (float)blockIdx.x * 1e-6f - (float)i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if constexpr (PACKED) {
                    asm volatile(
                        "{\n\t.reg .b64 rc, ra, rb;\n\tmov.b64 rc, {%0,%1};\n\tmov.b64 ra, {%2,%2};\n\tmov.b64 rb, {%3,%3};\n\t"
                        "fma.rn.f32x2 rc, rc, ra, rb;\n\tmov.b64 {%0,%1}, rc;\n\t}"
                        : "+f"(x[i]), "+f"(y[i])
                        : "f"(a), "f"(b));
                } else {
                    x[i] = fmaf(x[i], a, b);
                    y[i] = fmaf(y[i], a, b);
                }
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    if (s == 123456.789f) sink[0] = s;  // never true for the inputs used; keeps the chains live
}

// Measurement tool: NVLink flag ping-pong between two ranks inside one kernel launch per rank (tools/peer_latency.py).
// Rank 0 stores round i into the peer's flag word and spins on its own word until the peer has answered; rank 1 mirrors.
// mode 0: st.release.sys / ld.acquire.sys     mode 1: st.relaxed.sys / ld.relaxed.sys (no ordering)
// mode 2: __threadfence_system() + st.relaxed.sys, `payload` floats stored to the peer before every flag (data + flag)
// out[0] = nanoseconds (globaltimer) for all rounds on this rank.
__global__ void peer_pingpong_kernel(unsigned int* local, unsigned int* remote, float* remote_payload, int payload, int rank,
                                     int iters, int mode, unsigned int base, unsigned long long* out) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 1; i <= iters; ++i) {
        const unsigned int v = base + (unsigned)i;
        auto send = [&]() {
            for (int j = threadIdx.x; j < payload; j += blockDim.x) remote_payload[j] = (float)i;
            if (mode == 2) {
                __threadfence_system();
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                if (mode == 0)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(v) : "memory");
                else
                    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(v) : "memory");
            }
        };
        auto recv = [&]() {
            if (threadIdx.x == 0) {
                unsigned int seen;
                unsigned long long ta, tb;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ta));
                do {
                    if (mode == 0)
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local) : "memory");
                    else
                        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local) : "memory");
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tb));
                    if (tb - ta > 5000000000ull) __trap();
                } while ((int)(seen - v) < 0);
            }
            __syncthreads();
        };
        if (rank == 0) {
            send();
            recv();
        } else {
            recv();
            send();
        }
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

}  // namespace b200mm
