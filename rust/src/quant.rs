//! `sint8_quantize` / `sint8_dequantize` (upstream src/quant.rs:7-43) plus the per-group-scale extension.  The arithmetic
//! runs in libb200mm.so (csrc/host/quant.cc: global absmax, `x / absmax * 127`, round half away from zero, four int8 per
//! u32 little-endian along N) so that the Rust, C++ and Python front ends produce the same words -- `test_qdq` pins them.
use crate::ffi;

/// The float types callers quantise from (upstream is generic over `num_traits::Float`; the harness only ever uses f32).
pub trait QuantFloat: Copy + std::fmt::Debug {
    fn to_f32(self) -> f32;
    fn from_f32(v: f32) -> Self;
}
impl QuantFloat for f32 {
    fn to_f32(self) -> f32 {
        self
    }
    fn from_f32(v: f32) -> Self {
        v
    }
}
impl QuantFloat for f64 {
    fn to_f32(self) -> f32 {
        self as f32
    }
    fn from_f32(v: f32) -> Self {
        v as f64
    }
}

/// -> (K*N/4 packed words, absmax).  Panics where upstream's `assert!`s fire (length != K*N, length % 4 != 0).
pub fn sint8_quantize<F: QuantFloat>(matrix: &[F], K: usize, N: usize) -> (Vec<u32>, F) {
    assert!(matrix.len() == K * N);
    assert!(matrix.len() % 4 == 0);
    let as_f32: Vec<f32> = matrix.iter().map(|v| v.to_f32()).collect();
    let mut words = vec![0u32; K * N / 4];
    let mut absmax = 0f32;
    let rc = unsafe { ffi::wgpumm_sint8_quantize(as_f32.as_ptr(), K, N, words.as_mut_ptr(), &mut absmax) };
    assert!(rc == 0, "sint8_quantize failed");
    (words, F::from_f32(absmax))
}

pub fn sint8_dequantize(quantized_matrix: &[u32], absmax: f32, K: usize, N: usize) -> Vec<f32> {
    assert!(quantized_matrix.len() * 4 == K * N);
    let mut out = vec![0f32; K * N];
    let rc = unsafe { ffi::wgpumm_sint8_dequantize(quantized_matrix.as_ptr(), absmax, K, N, out.as_mut_ptr()) };
    assert!(rc == 0, "sint8_dequantize failed");
    out
}

/// Per-(row block of `group_k`, column) scales: K*N/4 weight words followed by ceil(K/group_k)*N f32 scales -- the B operand
/// of a `qgemv_sint8` kernel created with `group_k`.
pub fn sint8_quantize_grouped(matrix: &[f32], K: usize, N: usize, group_k: usize) -> Vec<u32> {
    assert!(matrix.len() == K * N);
    let mut packed = vec![0u32; unsafe { ffi::wgpumm_sint8_grouped_words(K, N, group_k) }];
    let rc = unsafe { ffi::wgpumm_sint8_quantize_grouped(matrix.as_ptr(), K, N, group_k, packed.as_mut_ptr()) };
    assert!(rc == 0, "sint8_quantize_grouped failed");
    packed
}

pub fn sint8_dequantize_grouped(packed: &[u32], K: usize, N: usize, group_k: usize) -> Vec<f32> {
    let mut out = vec![0f32; K * N];
    let rc = unsafe { ffi::wgpumm_sint8_dequantize_grouped(packed.as_ptr(), K, N, group_k, out.as_mut_ptr()) };
    assert!(rc == 0, "sint8_dequantize_grouped failed");
    out
}

#[cfg(test)]
mod tests {
    use super::*;

    /// the reference's only golden vector (src/quant.rs:48-64), also tests/golden/test_qdq.json
    #[test]
    pub fn test_qdq() {
        let row = [0.1f32, -0.1, 0.5, -0.5, 1.0, -1.0, 1.2, -1.2];
        let matrix: Vec<f32> = row.iter().chain(row.iter()).copied().collect();
        let (words, absmax) = sint8_quantize(&matrix, 4, 4);
        assert_eq!(words, vec![0xCB35_F50Bu32, 0x817F_966A, 0xCB35_F50B, 0x817F_966A]);
        assert_eq!(absmax, 1.2f32);
        let back = sint8_dequantize(&words, absmax, 4, 4);
        for (x, y) in matrix.iter().zip(back.iter()) {
            assert!((x - y).abs() < 0.01);
        }
    }

    #[test]
    pub fn test_qdq_grouped_one_group_equals_global() {
        let row = [0.1f32, -0.1, 0.5, -0.5, 1.0, -1.0, 1.2, -1.2];
        let matrix: Vec<f32> = row.iter().chain(row.iter()).copied().collect();
        let packed = sint8_quantize_grouped(&matrix, 4, 4, 4);
        assert_eq!(packed.len(), 4 + 4);
        let back = sint8_dequantize_grouped(&packed, 4, 4, 4);
        for (x, y) in matrix.iter().zip(back.iter()) {
            assert!((x - y).abs() < 0.01);
        }
    }

    #[test]
    #[should_panic]
    pub fn test_quantize_rejects_wrong_length() {
        sint8_quantize(&[0.5f32; 6], 2, 4);
    }
}
