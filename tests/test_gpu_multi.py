"""Launches tests/multi_gpu_check.py under torchrun when the box has >= 2 GPUs (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_n_sharded_sgemm_two_ranks(built_lib):
    import wgpu_mm_b200 as w
    n = w.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, found {n}")
    world = 2
    env = dict(os.environ, CHECK_SIZE="1024")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all ranks OK" in r.stdout
