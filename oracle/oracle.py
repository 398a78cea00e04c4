"""ctypes/numpy front-end of oracle/liboracle.so (built from oracle.c by oracle/Makefile).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c for the pinning status of each function
and the reference file:line it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_sz = C.c_size_t


def _declare(l):
    l.oracle_num_threads.restype = C.c_int
    l.oracle_set_num_threads.argtypes = [C.c_int]
    l.oracle_generate_weight_data.argtypes = [C.c_uint64, _f32p, _sz]
    l.oracle_generate_weight_data_at.argtypes = [C.c_uint64, C.c_uint64, _f32p, _sz]
    for name in ("oracle_mm_ref_literal", "oracle_mm_ref", "oracle_wgsl_gemm_1", "oracle_wgsl_gemm_1v",
                 "oracle_wgsl_gemm_2", "oracle_wgsl_gemm_3", "oracle_wgsl_gemm_4", "oracle_wgsl_gemm_5",
                 "oracle_wgsl_gemm_wonnx", "oracle_wgsl_bram", "oracle_wgsl_gemm3"):
        getattr(l, name).argtypes = [_f32p, _f32p, _f32p, _sz, _sz, _sz]
        getattr(l, name).restype = None
    l.oracle_mm_f64.argtypes = [_f32p, _f32p, _f64p, _sz, _sz, _sz]
    l.oracle_mm_f64_rows.argtypes = [_f32p, _f32p, _f64p, _i64p, _sz, _sz, _sz]
    l.oracle_max_abs_err.argtypes = [_f32p, _f32p, _sz]
    l.oracle_max_abs_err.restype = C.c_float
    l.oracle_err_vs_f64.argtypes = [_f32p, _f64p, _sz, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    l.oracle_sint8_quantize.argtypes = [_f32p, _sz, _sz, _u32p]
    l.oracle_sint8_quantize.restype = C.c_float
    l.oracle_sint8_dequantize.argtypes = [_u32p, C.c_float, _sz, _sz, _f32p]
    l.oracle_sint8_quantize_grouped.argtypes = [_f32p, _sz, _sz, _sz, _u32p, _f32p]
    l.oracle_sint8_dequantize_grouped.argtypes = [_u32p, _f32p, _sz, _sz, _sz, _f32p]
    l.oracle_qgemv_grouped_ref.argtypes = [_f32p, _u32p, _f32p, _f32p, _sz, _sz, _sz, _sz]
    l.oracle_qgemv_grouped_f64.argtypes = [_f32p, _u32p, _f32p, _f64p, _sz, _sz, _sz, _sz]
    l.oracle_wgsl_qgemv_1.argtypes = [_f32p, _u32p, _f32p, _sz, _sz, _sz, C.c_float]
    l.oracle_qgemv_ref.argtypes = [_f32p, _u32p, _f32p, _sz, _sz, _sz, C.c_float]
    l.oracle_qgemv_f64.argtypes = [_f32p, _u32p, _f64p, _sz, _sz, _sz, C.c_float]
    l.oracle_compute_dim.argtypes = [_sz, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    l.oracle_compute_dim.restype = C.c_int


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def generate_weight_data(seed: int, rows: int, cols: int, offset: int = 0) -> np.ndarray:
    """src/harness.rs:103-121, seeded: U[-10,10)/50 as f32, row-major rows x cols."""
    out = np.empty((rows, cols), dtype=np.float32)
    lib().oracle_generate_weight_data_at(seed, offset, out.reshape(-1), out.size)
    return out


def _mm(fn, A, B):
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape
    K2, N = B.shape
    assert K == K2
    # pre-fill with noise like src/harness.rs:55 -- an accumulate-into-C bug must show
    Cm = np.full((M, N), 123.25, dtype=np.float32)
    fn(A, B, Cm, M, N, K)
    return Cm


def mm_ref(A, B):
    """src/harness.rs:17-28 (vectorised loop order, bit-identical to the literal loop)."""
    return _mm(lib().oracle_mm_ref, A, B)


def mm_ref_literal(A, B):
    return _mm(lib().oracle_mm_ref_literal, A, B)


def mm_f64(A, B):
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape
    _, N = B.shape
    Cm = np.empty((M, N), dtype=np.float64)
    lib().oracle_mm_f64(A, B, Cm, M, N, K)
    return Cm


def mm_f64_rows(A, B, rows):
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    _, K = A.shape
    _, N = B.shape
    Cm = np.empty((len(rows), N), dtype=np.float64)
    lib().oracle_mm_f64_rows(A, B, Cm, rows, len(rows), N, K)
    return Cm


WGSL_GEMM = ("gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5", "gemm_wonnx", "bram", "gemm3")


def wgsl_gemm(name: str, A, B):
    """Per-shader accumulation-order restatement; name in WGSL_GEMM ('bram' also stands for bram8x8)."""
    fn = {"bram8x8": "oracle_wgsl_bram", "gemm3": "oracle_wgsl_gemm3"}.get(name, "oracle_wgsl_" + name)
    return _mm(getattr(lib(), fn), A, B)


def max_abs_err(gpu, cpu) -> float:
    """src/harness.rs:64-70 ('mae' is a max)."""
    g = np.ascontiguousarray(gpu, dtype=np.float32).reshape(-1)
    c = np.ascontiguousarray(cpu, dtype=np.float32).reshape(-1)
    assert g.size == c.size
    return float(lib().oracle_max_abs_err(g, c, g.size))


def err_vs_f64(gpu, ref64):
    """(max |gpu-ref|, max |ref|): north_star's relative error is their ratio."""
    g = np.ascontiguousarray(gpu, dtype=np.float32).reshape(-1)
    r = np.ascontiguousarray(ref64, dtype=np.float64).reshape(-1)
    e, m = C.c_double(), C.c_double()
    lib().oracle_err_vs_f64(g, r, g.size, C.byref(e), C.byref(m))
    return e.value, m.value


def sint8_quantize(matrix, K: int, N: int):
    """src/quant.rs:7-28 -> (uint32 words of K*N/4, absmax)."""
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(-1)
    assert m.size == K * N and m.size % 4 == 0
    out = np.empty(K * N // 4, dtype=np.uint32)
    absmax = lib().oracle_sint8_quantize(m, K, N, out)
    return out, float(absmax)


def sint8_dequantize(words, absmax: float, K: int, N: int) -> np.ndarray:
    """src/quant.rs:30-43."""
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    out = np.empty(K * N, dtype=np.float32)
    lib().oracle_sint8_dequantize(w, absmax, K, N, out)
    return out.reshape(K, N)


def wgsl_qgemv_1(A, Bq, N: int, K: int, absmax: float, batch: int = 1) -> np.ndarray:
    """shaders/gemv/qgemv_1.wgsl:10-39."""
    a = np.ascontiguousarray(A, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(Bq, dtype=np.uint32).reshape(-1)
    out = np.full((batch, N), 123.25, dtype=np.float32)
    lib().oracle_wgsl_qgemv_1(a, b, out, batch, N, K, absmax)
    return out


def qgemv_ref(A, Bq, M: int, N: int, K: int, absmax: float) -> np.ndarray:
    """src/harness.rs:42-48,58: mm_ref(A, sint8_dequantize(Bq, ABSMAX))."""
    a = np.ascontiguousarray(A, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(Bq, dtype=np.uint32).reshape(-1)
    out = np.full((M, N), 123.25, dtype=np.float32)
    lib().oracle_qgemv_ref(a, b, out, M, N, K, absmax)
    return out


def qgemv_f64(A, Bq, M: int, N: int, K: int, absmax: float) -> np.ndarray:
    a = np.ascontiguousarray(A, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(Bq, dtype=np.uint32).reshape(-1)
    out = np.empty((M, N), dtype=np.float64)
    lib().oracle_qgemv_f64(a, b, out, M, N, K, absmax)
    return out


def sint8_quantize_grouped(matrix, K: int, N: int, group_k: int):
    """Per-(group of group_k rows, column) absmax variant of src/quant.rs:7-28 (extension).
    -> (uint32 words of K*N/4, float32 scales of ceil(K/group_k) x N)."""
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(-1)
    assert m.size == K * N and N % 4 == 0 and group_k > 0
    out = np.empty(K * N // 4, dtype=np.uint32)
    scales = np.empty((-(-K // group_k), N), dtype=np.float32)
    lib().oracle_sint8_quantize_grouped(m, K, N, group_k, out, scales.reshape(-1))
    return out, scales


def sint8_dequantize_grouped(words, scales, K: int, N: int, group_k: int) -> np.ndarray:
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    s = np.ascontiguousarray(scales, dtype=np.float32).reshape(-1)
    out = np.empty(K * N, dtype=np.float32)
    lib().oracle_sint8_dequantize_grouped(w, s, K, N, group_k, out)
    return out.reshape(K, N)


def qgemv_grouped_ref(A, Bq, scales, M: int, N: int, K: int, group_k: int) -> np.ndarray:
    a = np.ascontiguousarray(A, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(Bq, dtype=np.uint32).reshape(-1)
    s = np.ascontiguousarray(scales, dtype=np.float32).reshape(-1)
    out = np.full((M, N), 123.25, dtype=np.float32)
    lib().oracle_qgemv_grouped_ref(a, b, s, out, M, N, K, group_k)
    return out


def qgemv_grouped_f64(A, Bq, scales, M: int, N: int, K: int, group_k: int) -> np.ndarray:
    a = np.ascontiguousarray(A, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(Bq, dtype=np.uint32).reshape(-1)
    s = np.ascontiguousarray(scales, dtype=np.float32).reshape(-1)
    out = np.empty((M, N), dtype=np.float64)
    lib().oracle_qgemv_grouped_f64(a, b, s, out, M, N, K, group_k)
    return out


def compute_dim(work_items: int, dim: str):
    """src/workload.rs:48-68; raises RuntimeError where the reference panics."""
    c, s = C.c_uint32(), C.c_uint32()
    rc = lib().oracle_compute_dim(work_items, {"X": 0, "Y": 1, "Z": 2}[dim], C.byref(c), C.byref(s))
    if rc != 0:
        raise RuntimeError("Compute limits exceeded")
    return c.value, s.value
