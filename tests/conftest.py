import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def built_lib():
    """libb200mm.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    import wgpu_mm_b200 as w
    if not os.path.exists(w.lib_path()):
        w.build()
    return w.lib()


@pytest.fixture(scope="session")
def gpu_ctx(built_lib):
    import wgpu_mm_b200 as w
    if w.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    ctx = w.Context(0)
    yield ctx
    ctx.close()
