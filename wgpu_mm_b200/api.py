"""Object wrappers over the C ABI handles (ctx / buffer / kernel).  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KernelId, KernelParamsC, check, lib


@dataclass
class KernelParams:
    workgroup_size: Sequence[int] = (0, 0, 0)
    absmax: float = 0.0
    batch: int = 0
    flags: int = 0
    tune: Sequence[int] = field(default_factory=lambda: (0, 0, 0, 0))
    group_k: int = 0

    def to_c(self) -> KernelParamsC:
        p = KernelParamsC()
        for i in range(3):
            p.workgroup_size[i] = int(self.workgroup_size[i])
        p.absmax = float(self.absmax)
        p.batch = int(self.batch)
        p.flags = int(self.flags)
        t = list(self.tune) + [0, 0, 0, 0]
        for i in range(4):
            p.tune[i] = int(t[i])
        p.group_k = int(self.group_k)
        return p


def _ptr(a: np.ndarray) -> C.c_void_p:
    """Address of a numpy array's data.  `a.ctypes.data_as(...)` builds a helper object on every call (tens of microseconds --
    measurable in the end-to-end timings of the microsecond-scale GEMV calls and even of mm_host); the array interface is a dict lookup."""
    return C.c_void_p(a.__array_interface__["data"][0])


class Context:
    """gpu_handle (src/harness.rs:87-101): one device, one in-order stream."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().b200mm_ctx_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().b200mm_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: the module globals the loader needs may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().b200mm_sync(self._h), self._h)

    def set_stream(self, cuda_stream: Optional[int]):
        check(lib().b200mm_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)), self._h)

    @property
    def stream(self) -> int:
        return lib().b200mm_ctx_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return int(lib().b200mm_ctx_launch_count(self._h))

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        name = C.create_string_buffer(256)
        check(lib().b200mm_ctx_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem), name, 256), self._h)
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "global_mem": mem.value, "name": name.value.decode()}

    # ---- buffers ----
    def buffer(self, nbytes: int) -> "Buffer":
        h = C.c_void_p()
        check(lib().b200mm_buffer_create(self._h, nbytes, C.byref(h)), self._h)
        return Buffer(self, h, nbytes)

    def buffer_from(self, array: np.ndarray) -> "Buffer":
        """create_buffer_init (src/harness.rs:135,158)."""
        a = np.ascontiguousarray(array)
        h = C.c_void_p()
        check(lib().b200mm_buffer_create_init(self._h, _ptr(a), a.nbytes, C.byref(h)), self._h)
        return Buffer(self, h, a.nbytes)

    def wrap(self, device_ptr: int, nbytes: int) -> "Buffer":
        h = C.c_void_p()
        check(lib().b200mm_buffer_wrap(self._h, C.c_void_p(device_ptr), nbytes, C.byref(h)), self._h)
        return Buffer(self, h, nbytes)

    def ipc_import(self, handle: bytes, nbytes: int) -> "Buffer":
        h = C.c_void_p()
        hb = C.create_string_buffer(handle, 64)
        check(lib().b200mm_ipc_import(self._h, hb, nbytes, C.byref(h)), self._h)
        return Buffer(self, h, nbytes)

    # ---- kernels ----
    def kernel(self, kernel_id: int, M: int, N: int, K: int, params: Optional[KernelParams] = None) -> "Kernel":
        """create_shader_module + create_compute_pipeline (src/harness.rs:179-191)."""
        h = C.c_void_p()
        p = (params or KernelParams()).to_c()
        check(lib().b200mm_kernel_get(self._h, int(kernel_id), M, N, K, C.byref(p), C.byref(h)), self._h)
        return Kernel(self, h, int(kernel_id), (M, N, K))

    def launch(self, kern: "Kernel", A: "Buffer", B: "Buffer", Cb: "Buffer", grid: Optional[Sequence[int]] = None):
        """mm (src/harness.rs:250-287): asynchronous, in order on this context."""
        g = (C.c_uint32 * 3)(*grid) if grid is not None else None
        check(lib().b200mm_launch(self._h, kern.handle, A.handle, B.handle, Cb.handle, g), self._h)

    def mm_host(self, kern: "Kernel", hostA: np.ndarray, hostB: np.ndarray, hostC: np.ndarray, dA: "Buffer", dB: "Buffer", dC: "Buffer"):
        """End-to-end call with host buffers: H2D(A,B) + launch + D2H(C), blocking."""
        check(lib().b200mm_mm_host(self._h, kern.handle, _ptr(hostA), hostA.nbytes,
                                   _ptr(hostB), hostB.nbytes, _ptr(hostC),
                                   hostC.nbytes, dA.handle, dB.handle, dC.handle), self._h)

    def timer_begin(self):
        check(lib().b200mm_timer_begin(self._h), self._h)

    def timer_end(self) -> float:
        ms = C.c_float()
        check(lib().b200mm_timer_end(self._h, C.byref(ms)), self._h)
        return float(ms.value)

    def measure_fma_peak(self, packed: bool = True, iters: int = 4096, reps: int = 3) -> float:
        """Measurement tool: FP32 FMA-pipe ceiling (TFLOP/s) from a register-only FFMA2 / FFMA microbenchmark."""
        out = C.c_double()
        check(lib().b200mm_measure_fma_peak(self._h, 1 if packed else 0, iters, reps, C.byref(out)), self._h)
        return float(out.value)

    def flush_l2(self):
        check(lib().b200mm_flush_l2(self._h), self._h)

    def peer_barrier(self, local_flags: "Buffer", peer_ptrs: Sequence[int], rank: int, world: int):
        arr = (C.c_void_p * world)(*[C.c_void_p(p) for p in peer_ptrs])
        check(lib().b200mm_peer_barrier(self._h, local_flags.handle, arr, rank, world), self._h)

    def unshard_columns(self, gathered_ptr: int, c_ptr: int, M: int, N: int, world: int):
        check(lib().b200mm_unshard_columns(self._h, C.c_void_p(gathered_ptr), C.c_void_p(c_ptr), M, N, world), self._h)


class Buffer:
    def __init__(self, ctx: Context, handle: C.c_void_p, nbytes: int):
        self.ctx, self._h, self.nbytes = ctx, handle, nbytes

    @property
    def handle(self):
        return self._h

    @property
    def ptr(self) -> int:
        return lib().b200mm_buffer_device_ptr(self._h) or 0

    def write(self, array: np.ndarray, offset: int = 0):
        a = array if (isinstance(array, np.ndarray) and array.flags.c_contiguous) else np.ascontiguousarray(array)
        check(lib().b200mm_buffer_write(self.ctx.handle, self._h, offset, _ptr(a), a.nbytes), self.ctx.handle)

    def read(self, dtype=np.float32, count: Optional[int] = None, offset: int = 0) -> np.ndarray:
        """to_cpu (src/harness.rs:289-302): blocking read-back."""
        itemsize = np.dtype(dtype).itemsize
        n = count if count is not None else (self.nbytes - offset) // itemsize
        out = np.empty(n, dtype=dtype)
        check(lib().b200mm_buffer_read(self.ctx.handle, self._h, offset, _ptr(out), out.nbytes), self.ctx.handle)
        return out

    def read_into(self, out: np.ndarray, offset: int = 0):
        check(lib().b200mm_buffer_read(self.ctx.handle, self._h, offset, _ptr(out), out.nbytes), self.ctx.handle)

    def read_2d_into(self, out: np.ndarray, offset: int, src_pitch: int, width_bytes: int, rows: int):
        """Blocking read of `rows` rows of width_bytes (row r from offset + r*src_pitch) into the contiguous array `out`."""
        check(lib().b200mm_buffer_read_2d(self.ctx.handle, self._h, offset, src_pitch, _ptr(out), width_bytes,
                                          width_bytes, rows), self.ctx.handle)

    def fill_weights(self, seed: int, n: int, offset: int = 0):
        check(lib().b200mm_buffer_fill_weights(self.ctx.handle, self._h, seed, offset, n), self.ctx.handle)

    def fill_weights_2d(self, seed: int, rows: int, cols: int, src_ld: int, src_col0: int, offset: int = 0):
        check(lib().b200mm_buffer_fill_weights_2d(self.ctx.handle, self._h, seed, offset, rows, cols, src_ld, src_col0), self.ctx.handle)

    def ipc_export(self) -> bytes:
        hb = C.create_string_buffer(64)
        check(lib().b200mm_ipc_export(self.ctx.handle, self._h, hb), self.ctx.handle)
        return hb.raw

    def free(self):
        if self._h is not None and self._h.value:
            lib().b200mm_buffer_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()


class Kernel:
    def __init__(self, ctx: Context, handle: C.c_void_p, kernel_id: int, dims):
        self.ctx, self._h, self.kernel_id, self.dims = ctx, handle, kernel_id, dims

    @property
    def handle(self):
        return self._h

    @property
    def name(self) -> str:
        return lib().b200mm_kernel_name(self.kernel_id).decode()

    def geometry(self):
        g, b = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
        check(lib().b200mm_kernel_geometry(self._h, g, b))
        return tuple(g), tuple(b)

    @property
    def workspace_bytes(self) -> int:
        return int(lib().b200mm_kernel_workspace_bytes(self._h))

    def profile(self, enable: bool = True):
        check(lib().b200mm_kernel_profile_enable(self.ctx.handle, self._h, 1 if enable else 0), self.ctx.handle)

    def profile_read(self, max_n: int = 256):
        """Durations (ms) of the dominant device kernel for the launches since the last read."""
        buf = (C.c_float * max_n)()
        n = C.c_int()
        check(lib().b200mm_kernel_profile_read(self.ctx.handle, self._h, buf, max_n, C.byref(n)), self.ctx.handle)
        return [float(buf[i]) for i in range(n.value)]

    def set_peers(self, rank: int, world: int, peer_ptrs: Sequence[int], ldc: int, col_offset: int):
        arr = (C.c_void_p * max(world, 1))(*[C.c_void_p(p) for p in peer_ptrs])
        check(lib().b200mm_kernel_set_peers(self._h, rank, world, arr, ldc, col_offset))

    def set_peer_flags(self, peer_ptrs: Optional[Sequence[int]], pingpong_stride: int = 0, deferred: bool = False):
        """In-kernel cross-rank completion (GEMV): see b200mm_kernel_set_peer_flags."""
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(p) for p in peer_ptrs]) if peer_ptrs else None
        check(lib().b200mm_kernel_set_peer_flags(self._h, arr, pingpong_stride, 1 if deferred else 0))

    def peer_wait(self):
        """Closes a chain of deferred launches: stream-ordered wait until every rank's last launch has landed here."""
        check(lib().b200mm_kernel_peer_wait(self.ctx.handle, self._h), self.ctx.handle)

    @property
    def peer_epoch(self) -> int:
        return int(lib().b200mm_kernel_peer_epoch(self._h))

    def free(self):
        if self._h is not None and self._h.value:
            lib().b200mm_kernel_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()
