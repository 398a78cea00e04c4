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

}  // namespace b200mm
