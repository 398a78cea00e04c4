"""wgpu_mm_b200 -- B200-native replacement for wgpu-mm's SGEMM / GEMV hot path.

The product is ``lib/libb200mm.so``: hand-written sm_100a CUDA kernels behind the C ABI of
``include/b200mm.h`` plus the C++ host mirror of the reference crate (``include/wgpu_mm.hpp``).
This Python package is only the thinnest possible binding over that C ABI, used by the tests,
``bench.py`` and the multi-GPU launcher (``torch.distributed`` is plumbing, not the product).

There is no CPU fallback: importing works anywhere, but every compute call needs the library and a
CUDA device and raises otherwise.
"""
from ._lib import (  # noqa: F401
    B200mmError,
    KernelId,
    Flags,
    lib,
    lib_path,
    build,
    device_count,
)
from .api import Context, Buffer, Kernel, KernelParams  # noqa: F401
from . import workload, quant, gemm, gemv, harness  # noqa: F401

__all__ = [
    "B200mmError", "KernelId", "Flags", "lib", "lib_path", "build", "device_count",
    "Context", "Buffer", "Kernel", "KernelParams", "workload", "quant", "gemm", "gemv", "harness",
]
