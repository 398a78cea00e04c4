"""Entry points of src/gemv.rs (plus the B200-native streaming GEMV kernels)."""
from .gemm import _entry

M, N, K = 1, 1024, 1024  # src/gemv.rs:5-7
ABSMAX = 2.0  # src/gemv.rs:8


def insert_matrix_dims(context: dict, dims=None):
    m, n, k = dims or (M, N, K)
    context.update(M=m, N=n, K=k)
    return (m, n, k)


qgemv_1, qgemv_sint8, gemv_f32 = (_entry(n) for n in ("qgemv_1", "qgemv_sint8", "gemv_f32"))
qgemv_sint8_grouped = _entry("qgemv_sint8_grouped")  # per-group scales, group_k = 128
