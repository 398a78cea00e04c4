// Host codec, mirror of src/quant.rs:7-43: one global absmax, symmetric int8, 4 values packed
// little-endian into a u32 along the contiguous (N) axis.  This is product code (the reference's
// quantiser also runs on the host); it is independent of oracle/.
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

}  // namespace quant
}  // namespace wgpu_mm
