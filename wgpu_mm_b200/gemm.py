"""Entry points of src/gemm.rs (plus the orphan shaders and the B200-native SGEMM kernels)."""
from .workload import entry_workload

M = N = K = 1024  # src/gemm.rs:5-7


def insert_matrix_dims(context: dict, dims=None):
    m, n, k = dims or (M, N, K)
    context.update(M=m, N=n, K=k)
    return (m, n, k)


def _entry(name):
    def fn(context: dict):
        wl, kid = entry_workload(name, context["M"], context["N"], context["K"])
        context.update(workgroup_size_x=wl.size.x, workgroup_size_y=wl.size.y, workgroup_size_z=wl.size.z)
        return wl, name
    fn.__name__ = name
    return fn


gemm_1, gemm_1v, gemm_2, gemm_3, gemm_4, gemm_5 = (_entry(n) for n in ("gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5"))
gemm_wonnx, bram, bram8x8, gemm3 = (_entry(n) for n in ("gemm_wonnx", "bram", "bram8x8", "gemm3"))
sgemm_simt, sgemm_tc3x, sgemm_tc3x_1x = (_entry(n) for n in ("sgemm_simt", "sgemm_tc3x", "sgemm_tc3x_1x"))
