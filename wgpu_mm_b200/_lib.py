"""ctypes loader for lib/libb200mm.so and the declarations of include/b200mm.h + include/wgpu_mm_c.h."""
from __future__ import annotations

import ctypes as C
import enum
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "lib", "libb200mm.so")


class B200mmError(RuntimeError):
    """Non-zero b200mm_status; the reference panics in the same places."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[b200mm {code}] {message}")
        self.code = code
        self.message = message


class KernelId(enum.IntEnum):
    GEMM_1 = 1
    GEMM_1V = 2
    GEMM_2 = 3
    GEMM_3 = 4
    GEMM_4 = 5
    GEMM_5 = 6
    GEMM_WONNX = 7
    BRAM = 8
    BRAM8X8 = 9
    GEMM3 = 10
    QGEMV_1 = 11
    SGEMM_SIMT = 32
    SGEMM_TC3X = 33
    GEMV_F32 = 34
    QGEMV_SINT8 = 35


class Flags(enum.IntFlag):
    NONE = 0
    TC3X_1X = 0x1
    CONST_B = 0x10
    PEER_STORE = 0x2
    SEQUENTIAL_K = 0x4
    AUTOTUNE = 0x8


ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_LIMITS, ERR_UNSUPPORTED, ERR_TOLERANCE = -1, -2, -3, -4, -5, -6


class KernelParamsC(C.Structure):
    _fields_ = [
        ("workgroup_size", C.c_uint32 * 3),
        ("absmax", C.c_float),
        ("batch", C.c_uint32),
        ("flags", C.c_uint32),
        ("tune", C.c_uint32 * 4),
        ("group_k", C.c_uint32),
    ]


class ReportC(C.Structure):
    _fields_ = [
        ("max_abs_err", C.c_double), ("max_rel_err_f64", C.c_double), ("kernel_ms", C.c_double),
        ("wall_ns", C.c_double), ("gflops", C.c_double), ("kernel_gflops", C.c_double),
        ("kernel_gbps", C.c_double), ("seed", C.c_uint64), ("grid", C.c_uint32 * 3),
        ("block", C.c_uint32 * 3), ("rotated", C.c_int),
    ]


def lib_path() -> str:
    return _SO


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libb200mm.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", _CSRC] + (["-B"] if force else [])
    res = subprocess.run(args, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libb200mm.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return _SO


_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Fails loudly if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C "
                f"{_CSRC}).  wgpu_mm_b200 has no CPU fallback.")
        _lib = C.CDLL(_SO, mode=C.RTLD_GLOBAL)
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    vp, sz, u32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)
    l.b200mm_version.restype = C.c_char_p
    l.b200mm_device_count.restype = C.c_int
    l.b200mm_last_error.restype = C.c_char_p
    l.b200mm_last_error.argtypes = [vp]
    l.b200mm_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    l.b200mm_ctx_destroy.argtypes = [vp]
    l.b200mm_ctx_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(sz), C.c_char_p, sz]
    l.b200mm_ctx_set_stream.argtypes = [vp, vp]
    l.b200mm_ctx_stream.restype = vp
    l.b200mm_ctx_stream.argtypes = [vp]
    l.b200mm_sync.argtypes = [vp]
    l.b200mm_ctx_launch_count.restype = C.c_uint64
    l.b200mm_ctx_launch_count.argtypes = [vp]
    l.b200mm_buffer_create.argtypes = [vp, sz, C.POINTER(vp)]
    l.b200mm_buffer_create_init.argtypes = [vp, vp, sz, C.POINTER(vp)]
    l.b200mm_buffer_wrap.argtypes = [vp, vp, sz, C.POINTER(vp)]
    l.b200mm_buffer_free.argtypes = [vp, vp]
    l.b200mm_buffer_device_ptr.restype = vp
    l.b200mm_buffer_device_ptr.argtypes = [vp]
    l.b200mm_buffer_bytes.restype = sz
    l.b200mm_buffer_bytes.argtypes = [vp]
    l.b200mm_buffer_write.argtypes = [vp, vp, sz, vp, sz]
    l.b200mm_buffer_read.argtypes = [vp, vp, sz, vp, sz]
    l.b200mm_host_alloc.argtypes = [sz, C.POINTER(vp)]
    l.b200mm_host_free.argtypes = [vp]
    l.b200mm_buffer_fill_weights.argtypes = [vp, vp, C.c_uint64, C.c_uint64, sz]
    l.b200mm_buffer_fill_weights_2d.argtypes = [vp, vp, C.c_uint64, C.c_uint64, sz, sz, sz, sz]
    l.b200mm_kernel_get.argtypes = [vp, C.c_int, sz, sz, sz, C.POINTER(KernelParamsC), C.POINTER(vp)]
    l.b200mm_kernel_free.argtypes = [vp, vp]
    l.b200mm_kernel_name.restype = C.c_char_p
    l.b200mm_kernel_name.argtypes = [C.c_int]
    l.b200mm_kernel_geometry.argtypes = [vp, u32p, u32p]
    l.b200mm_kernel_workspace_bytes.restype = sz
    l.b200mm_kernel_workspace_bytes.argtypes = [vp]
    l.b200mm_launch.argtypes = [vp, vp, vp, vp, vp, u32p]
    l.b200mm_launch_ptr.argtypes = [vp, vp, vp, vp, vp, u32p]
    l.b200mm_mm_host.argtypes = [vp, vp, vp, sz, vp, sz, vp, sz, vp, vp, vp]
    l.b200mm_timer_begin.argtypes = [vp]
    l.b200mm_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    l.b200mm_flush_l2.argtypes = [vp]
    l.b200mm_buffer_read_2d.argtypes = [vp, vp, sz, sz, vp, sz, sz, sz]
    l.b200mm_measure_fma_peak.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    l.b200mm_kernel_profile_enable.argtypes = [vp, vp, C.c_int]
    l.b200mm_kernel_profile_read.argtypes = [vp, vp, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    l.b200mm_ipc_export.argtypes = [vp, vp, vp]
    l.b200mm_ipc_import.argtypes = [vp, vp, sz, C.POINTER(vp)]
    l.b200mm_kernel_set_peers.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), sz, sz]
    l.b200mm_kernel_set_peer_flags.argtypes = [vp, C.POINTER(vp), sz, C.c_int]
    l.b200mm_kernel_peer_wait.argtypes = [vp, vp]
    l.b200mm_kernel_peer_epoch.argtypes = [vp]
    l.b200mm_kernel_peer_epoch.restype = C.c_uint
    l.b200mm_peer_barrier.argtypes = [vp, vp, C.POINTER(vp), C.c_int, C.c_int]
    l.b200mm_unshard_columns.argtypes = [vp, vp, vp, sz, sz, C.c_int]
    l.b200mm_tc3x_schedule.argtypes = [sz, sz, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    l.b200mm_tc3x_plan.argtypes = [sz, sz, sz, C.c_int, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_int)]
    l.b200mm_tc3x_schedule_cover.argtypes = [sz, sz, sz, C.c_int, C.c_int, C.c_int, C.c_int, vp, sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    l.b200mm_tc3x_schedule_replay.argtypes = [sz, sz, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    # wgpu_mm_c.h
    l.wgpumm_run_test.argtypes = [C.c_char_p, sz, sz, sz, C.c_uint64, C.c_int, C.c_int, C.POINTER(ReportC)]
    l.wgpumm_run_test_ex.argtypes = [C.c_char_p, sz, sz, sz, C.c_uint64, C.c_int, C.c_int, u32p, u32p, C.c_int, C.POINTER(ReportC)]
    l.wgpumm_last_panic.restype = C.c_char_p
    l.wgpumm_entry_workload.argtypes = [C.c_char_p, sz, sz, sz, u32p, u32p, C.POINTER(C.c_int)]
    l.wgpumm_sint8_quantize.argtypes = [vp, sz, sz, vp, C.POINTER(C.c_float)]
    l.wgpumm_sint8_dequantize.argtypes = [vp, C.c_float, sz, sz, vp]
    l.wgpumm_sint8_grouped_words.restype = sz
    l.wgpumm_sint8_grouped_words.argtypes = [sz, sz, sz]
    l.wgpumm_sint8_quantize_grouped.argtypes = [vp, sz, sz, sz, vp]
    l.wgpumm_sint8_dequantize_grouped.argtypes = [vp, sz, sz, sz, vp]
    l.wgpumm_compute_dim.argtypes = [sz, C.c_int, u32p, u32p]
    l.wgpumm_workload_ceil.restype = sz
    l.wgpumm_workload_ceil.argtypes = [sz, sz]


def device_count() -> int:
    return int(lib().b200mm_device_count())


def check(rc: int, ctx=None) -> None:
    if rc != 0:
        msg = lib().b200mm_last_error(ctx)
        raise B200mmError(rc, msg.decode() if msg else "unknown error")
