// Host codec, mirror of src/quant.rs:7-43: one global absmax, symmetric int8, 4 values packed
// little-endian into a u32 along the contiguous (N) axis.  This is product code (the reference's
// quantiser also runs on the host); it is independent of oracle/.
#include <algorithm>
#include <cmath>
#include <limits>

#include "../../../include/wgpu_mm.hpp"

namespace wgpu_mm {
namespace quant {

static inline int32_t f32_as_i32(float v) {  // Rust `as i32`: saturating, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
    if (v <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
    return (int32_t)v;
}

std::pair<std::vector<uint32_t>, float> sint8_quantize(const std::vector<float>& matrix, size_t K, size_t N) {
    if (matrix.size() != K * N) throw Panic("assertion failed: matrix.len() == K * N");  // src/quant.rs:12
    if (matrix.size() % 4 != 0) throw Panic("assertion failed: matrix.len() % 4 == 0");  // src/quant.rs:13
    const size_t block_size = 4;
    std::vector<uint32_t> quantized(K * N / block_size, 0u);
    float absmax = 0.f;
    for (float x : matrix) {
        const float a = std::fabs(x);
        if (a > absmax) absmax = a;
    }
    const float sf = 127.f;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < K * N; i += block_size) {
        uint32_t packed = 0;
        for (size_t j = 0; j < 4; ++j) {
            const float q = std::round(matrix[i + j] / absmax * sf);  // half away from zero, like f32::round
            packed |= ((uint32_t)f32_as_i32(q) & 0xFFu) << (8 * j);
        }
        quantized[i / block_size] = packed;
    }
    return {std::move(quantized), absmax};
}

std::vector<float> sint8_dequantize(const std::vector<uint32_t>& quantized, float absmax, size_t K, size_t N) {
    const size_t block_size = 4;
    if (quantized.size() * block_size < K * N) throw Panic("index out of bounds: quantized matrix too short");
    std::vector<float> matrix(K * N, 0.f);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < K * N; i += block_size) {
        const int32_t p = (int32_t)quantized[i / block_size];
        matrix[i + 0] = (float)((int32_t)((uint32_t)p << 24) >> 24) / 127.0f * absmax;
        matrix[i + 1] = (float)((int32_t)((uint32_t)p << 16) >> 24) / 127.0f * absmax;
        matrix[i + 2] = (float)((int32_t)((uint32_t)p << 8) >> 24) / 127.0f * absmax;
        matrix[i + 3] = (float)(p >> 24) / 127.0f * absmax;
    }
    return matrix;
}

GroupedSint8 sint8_quantize_grouped(const std::vector<float>& matrix, size_t K, size_t N, size_t group_k) {
    if (matrix.size() != K * N) throw Panic("assertion failed: matrix.len() == K * N");
    if (N % 4 != 0) throw Panic("assertion failed: N % 4 == 0");
    if (group_k == 0) throw Panic("assertion failed: group_k > 0");
    GroupedSint8 out;
    out.K = K;
    out.N = N;
    out.group_k = group_k;
    const size_t groups = out.groups(), nwords = K * N / 4;
    out.packed.assign(nwords + groups * N, 0u);
    uint32_t* words = out.packed.data();
    float* scales = reinterpret_cast<float*>(out.packed.data() + nwords);
#pragma omp parallel for schedule(static)
    for (size_t g = 0; g < groups; ++g) {
        const size_t k0 = g * group_k, k1 = std::min(K, k0 + group_k);
        float* sc = scales + g * N;
        for (size_t k = k0; k < k1; ++k) {  // row-wise sweep: contiguous reads
            const float* row = matrix.data() + k * N;
            for (size_t n = 0; n < N; ++n) sc[n] = std::max(sc[n], std::fabs(row[n]));
        }
        for (size_t k = k0; k < k1; ++k) {
            const float* row = matrix.data() + k * N;
            for (size_t n = 0; n < N; n += 4) {
                uint32_t packed = 0;
                for (size_t j = 0; j < 4; ++j) {
                    const float q = std::round(row[n + j] / sc[n + j] * 127.f);  // same expression as src/quant.rs:21-24
                    packed |= ((uint32_t)f32_as_i32(q) & 0xFFu) << (8 * j);
                }
                words[(k * N + n) / 4] = packed;
            }
        }
    }
    return out;
}

std::vector<float> sint8_dequantize_grouped(const GroupedSint8& q) {
    const size_t K = q.K, N = q.N;
    if (q.group_k == 0 || q.packed.size() < K * N / 4 + q.groups() * N) throw Panic("index out of bounds: grouped matrix too short");
    std::vector<float> matrix(K * N, 0.f);
    const float* scales = q.scales();
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < K; ++k) {
        const float* sc = scales + (k / q.group_k) * N;
        for (size_t n = 0; n < N; n += 4) {
            const int32_t p = (int32_t)q.packed[(k * N + n) / 4];
            matrix[k * N + n + 0] = (float)((int32_t)((uint32_t)p << 24) >> 24) / 127.0f * sc[n + 0];
            matrix[k * N + n + 1] = (float)((int32_t)((uint32_t)p << 16) >> 24) / 127.0f * sc[n + 1];
            matrix[k * N + n + 2] = (float)((int32_t)((uint32_t)p << 8) >> 24) / 127.0f * sc[n + 2];
            matrix[k * N + n + 3] = (float)(p >> 24) / 127.0f * sc[n + 3];
        }
    }
    return matrix;
}

}  // namespace quant
}  // namespace wgpu_mm
