"""Binding of the host Workload helpers (src/workload.rs:1-69) -- the arithmetic runs in libb200mm.so."""
import ctypes as C
from dataclasses import dataclass

from ._lib import B200mmError, lib

MAX_WORKGROUP_SIZE_X = 256
MAX_WORKGROUP_SIZE_Y = 256
MAX_WORKGROUP_SIZE_Z = 64
MAX_COMPUTE_WORKGROUPS_PER_DIMENSION = 65535


@dataclass(frozen=True)
class WorkgroupCount:
    x: int
    y: int
    z: int


@dataclass(frozen=True)
class WorkgroupSize:
    x: int
    y: int
    z: int


@dataclass(frozen=True)
class Workload:
    count: WorkgroupCount
    size: WorkgroupSize

    @staticmethod
    def ceil(num: int, div: int) -> int:
        return int(lib().wgpumm_workload_ceil(num, div))

    @staticmethod
    def compute_dim(work_items: int, dim: str):
        """(workgroup_count, workgroup_size); raises like the reference's panic!("Compute limits exceeded")."""
        c, s = C.c_uint32(), C.c_uint32()
        rc = lib().wgpumm_compute_dim(work_items, {"X": 0, "Y": 1, "Z": 2}[dim], C.byref(c), C.byref(s))
        if rc != 0:
            raise B200mmError(rc, lib().wgpumm_last_panic().decode())
        return c.value, s.value


def entry_workload(name: str, M: int = 0, N: int = 0, K: int = 0):
    """(Workload, kernel_id) produced by entry point `name` (gemm_1 .. qgemv_1, sgemm_tc3x ...)."""
    g, b, kid = (C.c_uint32 * 3)(), (C.c_uint32 * 3)(), C.c_int()
    rc = lib().wgpumm_entry_workload(name.encode(), M, N, K, g, b, C.byref(kid))
    if rc != 0:
        raise B200mmError(rc, lib().wgpumm_last_panic().decode())
    return Workload(WorkgroupCount(*g), WorkgroupSize(*b)), kid.value
