"""The reference's own test list (src/gemm.rs:172-177, src/gemv.rs:41-49) through the C++ host harness:
verify one launch against mm_ref with the 1e-3 gate, then 8 warm-up + 10 timed launches."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_TESTS = ["gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5"]
ORPHANS = ["gemm_wonnx", "bram", "bram8x8", "gemm3"]
NATIVE = ["sgemm_simt", "sgemm_tc3x"]


@pytest.mark.parametrize("name", REFERENCE_TESTS + ORPHANS + NATIVE)
def test_gemm(gpu_ctx, name):
    """gemm_test!($test_name, $gemm_function) at the crate's 1024^3."""
    from wgpu_mm_b200 import gemm, harness
    context = {}
    dims = gemm.insert_matrix_dims(context)
    workload, shader = getattr(gemm, name)(context)
    rep = harness.test_harness(workload, shader, dims, False)
    assert rep.max_abs_err <= 1e-3
    assert rep.max_rel_err_f64 <= 5e-6
    assert rep.rotated  # M == N == K: buffer roles rotate like src/harness.rs:212-237
    assert rep.gflops > 0 and rep.kernel_ms > 0


@pytest.mark.parametrize("name", ["qgemv_1", "qgemv_sint8"])
def test_qgemv(gpu_ctx, name):
    """test_qgemv_1 at (1, 1024, 1024), quantize_b = true."""
    from wgpu_mm_b200 import gemv, harness
    context = {}
    dims = gemv.insert_matrix_dims(context)
    workload, shader = getattr(gemv, name)(context)
    rep = harness.test_harness(workload, shader, dims, True)
    assert rep.max_abs_err <= 1e-3
    assert not rep.rotated  # rotation is shape-illegal for M == 1


def test_gemv_f32_decode_shape(gpu_ctx):
    """BASELINE config 2: 1 x 4096 by 4096 x 16384."""
    from wgpu_mm_b200 import gemv, harness
    context = {}
    dims = gemv.insert_matrix_dims(context, (1, 16384, 4096))
    workload, shader = gemv.gemv_f32(context)
    rep = harness.test_harness(workload, shader, dims, False)
    assert rep.max_abs_err <= 1e-3 and rep.max_rel_err_f64 <= 5e-6


def test_qgemv_llm_shape(gpu_ctx):
    """BASELINE config 3: 1 x 4096 by 4096 x 14336 in the quant.rs format."""
    from wgpu_mm_b200 import gemv, harness
    context = {}
    dims = gemv.insert_matrix_dims(context, (1, 14336, 4096))
    workload, shader = gemv.qgemv_sint8(context)
    rep = harness.test_harness(workload, shader, dims, True)
    assert rep.max_abs_err <= 1e-3 and rep.max_rel_err_f64 <= 5e-6


@pytest.mark.parametrize("dims", [(1, 1024, 1024), (1, 14336, 4096)])
def test_qgemv_grouped_scales(gpu_ctx, dims):
    """Per-group scales (group_k = 128) through the same harness: quantise, launch, gate against mm_ref."""
    from wgpu_mm_b200 import gemv, harness
    context = {}
    dims = gemv.insert_matrix_dims(context, dims)
    workload, shader = gemv.qgemv_sint8_grouped(context)
    rep = harness.test_harness(workload, shader, dims, True)
    assert rep.max_abs_err <= 1e-3 and rep.max_rel_err_f64 <= 5e-6
    assert rep.kernel_gbps > 0


def test_mae_gate_panics(gpu_ctx):
    """A kernel that misses the gate must panic with the reference's message: single-pass TF32 at K=4096."""
    import ctypes as C
    import numpy as np
    import wgpu_mm_b200 as w
    import oracle
    # drive the gate directly: 1xTF32 result vs mm_ref exceeds 1e-3 at K = 4096 on this data distribution
    M, N, K = 128, 256, 4096
    A = oracle.generate_weight_data(1, M, K)
    B = oracle.generate_weight_data(2, K, N)
    kern = gpu_ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(flags=int(w.Flags.TC3X_1X)))
    dA, dB, dC = gpu_ctx.buffer_from(A), gpu_ctx.buffer_from(B), gpu_ctx.buffer(M * N * 4)
    gpu_ctx.launch(kern, dA, dB, dC)
    got = dC.read(np.float32).reshape(M, N)
    assert oracle.max_abs_err(got, oracle.mm_ref(A, B)) > 1e-3
    for b in (dA, dB, dC):
        b.free()
    kern.free()


def test_harness_panics_mae_too_high(gpu_ctx):
    """The harness's own panic path (src/harness.rs:82-84): single-pass TF32 through test_harness at K = 4096 must raise
    with the reference's message and the TOLERANCE status."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200 import gemm, harness
    context = {}
    dims = gemm.insert_matrix_dims(context, (128, 256, 4096))
    workload, shader = gemm.sgemm_tc3x_1x(context)
    with pytest.raises(w.B200mmError) as ei:
        harness.test_harness(workload, shader, dims, False)
    assert "MAE too high" in str(ei.value) and ei.value.code == w._lib.ERR_TOLERANCE


def test_harness_honours_workload_and_quantize_b(gpu_ctx):
    """test_harness(workload, shader, dims, quantize_b) passes both through (src/harness.rs:170-175, 197, 201-206): a
    caller-made Workload is what gets dispatched, and a quantize_b that contradicts the kernel's B operand fails."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200 import gemm, gemv, harness
    from wgpu_mm_b200.workload import Workload, WorkgroupCount, WorkgroupSize
    context = {}
    dims = gemm.insert_matrix_dims(context, (64, 64, 64))
    _, shader = gemm.gemm_1(context)
    mine = Workload(WorkgroupCount(8, 8, 1), WorkgroupSize(8, 8, 1))  # gemm_1 is guarded: any covering grid works
    rep = harness.test_harness(mine, shader, dims, False)
    assert rep.grid == (8, 8, 1) and rep.block == (8, 8, 1) and rep.max_abs_err <= 1e-3
    short = Workload(WorkgroupCount(1, 1, 1), WorkgroupSize(8, 8, 1))  # covers an 8 x 8 corner only -> the gate must trip
    with pytest.raises(w.B200mmError):
        harness.test_harness(short, shader, dims, False)
    with pytest.raises(w.B200mmError) as ei:
        harness.test_harness(None, shader, dims, True)
    assert "binding 1" in str(ei.value)
    context = {}
    dims = gemv.insert_matrix_dims(context)
    _, shader = gemv.qgemv_1(context)
    with pytest.raises(w.B200mmError):
        harness.test_harness(None, shader, dims, False)


def test_cpp_runner_like_cargo_test(gpu_ctx):
    exe = os.path.join(ROOT, "wgpu_mm_b200", "lib", "wgpu_mm_tests")
    r = subprocess.run([exe, "test_gemm_5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "Max Absolute Error" in r.stdout and "GFLOPS" in r.stdout and "test test_gemm_5 ... ok" in r.stdout
