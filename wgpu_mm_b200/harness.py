"""test_harness (src/harness.rs:170-248) -- runs in libb200mm.so (host/harness.cc); this is the binding."""
import ctypes as C
from dataclasses import dataclass

from ._lib import B200mmError, ReportC, lib


@dataclass
class Report:
    max_abs_err: float
    max_rel_err_f64: float
    kernel_ms: float
    wall_ns: float
    gflops: float
    kernel_gflops: float
    kernel_gbps: float
    seed: int
    grid: tuple
    block: tuple
    rotated: bool


def test_harness(workload, shader: str, dims, quantize_b: bool, seed: int = 0, device: int = 0, verbose: bool = False) -> Report:
    """`shader` is the entry-point name returned by gemm.* / gemv.* (the reference passes rendered WGSL).

    `workload` (a wgpu_mm_b200.workload.Workload, normally the one the entry point returned) is honoured exactly as in
    the reference: its count is the dispatch grid and its size the workgroup size of the faithful ports
    (src/harness.rs:197,213); None = what the entry point produces.  `quantize_b` selects the B operand
    (src/harness.rs:201-206); a value that does not match the kernel fails ("binding 1 type mismatch").
    Raises B200mmError where the reference panics ("MAE too high", "No GPU found ...").
    """
    M, N, K = dims
    rep = ReportC()
    grid = block = None
    if workload is not None:
        grid = (C.c_uint32 * 3)(workload.count.x, workload.count.y, workload.count.z)
        block = (C.c_uint32 * 3)(workload.size.x, workload.size.y, workload.size.z)
    rc = lib().wgpumm_run_test_ex(shader.encode(), M, N, K, seed, device, 1 if verbose else 0, grid, block,
                                  1 if quantize_b else 0, C.byref(rep))
    if rc != 0:
        raise B200mmError(rc, lib().wgpumm_last_panic().decode())
    return Report(rep.max_abs_err, rep.max_rel_err_f64, rep.kernel_ms, rep.wall_ns, rep.gflops, rep.kernel_gflops,
                  rep.kernel_gbps, rep.seed, tuple(rep.grid), tuple(rep.block), bool(rep.rotated))


test_harness.__test__ = False  # not a pytest test
