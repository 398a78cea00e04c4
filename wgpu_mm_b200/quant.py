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
