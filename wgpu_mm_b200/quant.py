"""Binding of the host codec (src/quant.rs:7-43); the arithmetic runs in libb200mm.so (host/quant.cc)."""
import ctypes as C

import numpy as np

from ._lib import B200mmError, lib


def sint8_quantize(matrix, K: int, N: int):
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(-1)
    if m.size != K * N:
        raise B200mmError(-1, "assertion failed: matrix.len() == K * N")
    if m.size % 4 != 0:
        raise B200mmError(-1, "assertion failed: matrix.len() % 4 == 0")
    out = np.empty(K * N // 4, dtype=np.uint32)
    absmax = C.c_float()
    rc = lib().wgpumm_sint8_quantize(m.ctypes.data_as(C.c_void_p), K, N, out.ctypes.data_as(C.c_void_p), C.byref(absmax))
    if rc != 0:
        raise B200mmError(rc, lib().wgpumm_last_panic().decode())
    return out, float(absmax.value)


def sint8_dequantize(words, absmax: float, K: int, N: int) -> np.ndarray:
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    if w.size * 4 < K * N:
        raise B200mmError(-1, "index out of bounds: quantized matrix too short")
    out = np.empty(K * N, dtype=np.float32)
    rc = lib().wgpumm_sint8_dequantize(w.ctypes.data_as(C.c_void_p), absmax, K, N, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise B200mmError(rc, lib().wgpumm_last_panic().decode())
    return out.reshape(K, N)


def sint8_quantize_grouped(matrix, K: int, N: int, group_k: int) -> np.ndarray:
    """Per-group scales (extension of src/quant.rs:17).  Returns the packed uint32 array the grouped
    qgemv_sint8 kernel takes as B: K*N/4 weight words, then ceil(K/group_k)*N float32 scales (bit-cast)."""
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(-1)
    if m.size != K * N:
        raise B200mmError(-1, "assertion failed: matrix.len() == K * N")
    if N % 4 != 0 or group_k <= 0:
        raise B200mmError(-1, "assertion failed: N % 4 == 0 && group_k > 0")
    out = np.empty(lib().wgpumm_sint8_grouped_words(K, N, group_k), dtype=np.uint32)
    rc = lib().wgpumm_sint8_quantize_grouped(m.ctypes.data_as(C.c_void_p), K, N, group_k, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise B200mmError(rc, lib().wgpumm_last_panic().decode())
    return out


def split_grouped(packed, K: int, N: int, group_k: int):
    """(weight words, scales[groups, N]) views of a packed grouped buffer."""
    p = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1)
    nw = K * N // 4
    return p[:nw], p[nw:].view(np.float32).reshape(-(-K // group_k), N)


def sint8_dequantize_grouped(packed, K: int, N: int, group_k: int) -> np.ndarray:
    p = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1)
    if group_k <= 0 or p.size < lib().wgpumm_sint8_grouped_words(K, N, group_k):
        raise B200mmError(-1, "index out of bounds: grouped matrix too short")
    out = np.empty(K * N, dtype=np.float32)
    rc = lib().wgpumm_sint8_dequantize_grouped(p.ctypes.data_as(C.c_void_p), K, N, group_k, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise B200mmError(rc, lib().wgpumm_last_panic().decode())
    return out.reshape(K, N)
