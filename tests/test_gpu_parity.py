"""GPU parity tests: every kernel is called through the C ABI (include/b200mm.h) and compared with the
CPU oracle on the same seeded inputs.

Tolerances (north_star + reference):
  GATE      max-abs-error <= 1e-3 vs mm_ref on U[-0.2,0.2) data       src/harness.rs:82 (the reference's gate)
  REL_F64   max |gpu - fp64| / max |fp64| <= 5e-6                      north_star "max relative error against an FP64 host GEMM"
  WGSL      faithful ports: BIT-EXACT against their oracle restatement (same per-output arithmetic order)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GATE = 1e-3
REL_F64 = 5e-6


def _run(ctx, kid, A, B, M, N, K, params=None, grid=None, b_dtype=np.float32):
    import wgpu_mm_b200 as w
    kern = ctx.kernel(kid, M, N, K, params)
    dA = ctx.buffer_from(np.ascontiguousarray(A, dtype=np.float32))
    dB = ctx.buffer_from(np.ascontiguousarray(B, dtype=b_dtype))
    batch = params.batch if (params and params.batch) else 1
    noise = np.full(batch * M * N, 123.25, dtype=np.float32)  # C is overwritten, never accumulated into (src/harness.rs:55)
    dC = ctx.buffer_from(noise)
    ctx.launch(kern, dA, dB, dC, grid)
    out = dC.read(np.float32).reshape(batch * M, N)
    for b in (dA, dB, dC):
        b.free()
    kern.free()
    return out


def _check(oracle, got, A, B, rel=REL_F64):
    ref = oracle.mm_ref(A, B)
    assert not np.isnan(got).any()
    assert oracle.max_abs_err(got, ref) <= GATE
    e, m = oracle.err_vs_f64(got, oracle.mm_f64(A, B))
    assert e / m <= rel, f"rel err vs fp64 {e / m:.3e}"
    return e / m


PORTS = ["gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5", "gemm_wonnx", "bram", "bram8x8", "gemm3"]


@pytest.mark.parametrize("name", PORTS)
@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 256, 96), (96, 64, 160)])
def test_wgsl_ports_bit_exact(gpu_ctx, oracle, name, shape):
    """Faithful ports launched with the entry point's Workload (grid AND block) vs the per-shader restatement."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200.workload import entry_workload
    M, N, K = shape
    A = oracle.generate_weight_data(1, M, K)
    B = oracle.generate_weight_data(2, K, N)
    wl, kid = entry_workload(name, M, N, K)
    prm = w.KernelParams(workgroup_size=(wl.size.x, wl.size.y, wl.size.z))
    got = _run(gpu_ctx, kid, A, B, M, N, K, prm, grid=(wl.count.x, wl.count.y, wl.count.z))
    want = oracle.wgsl_gemm(name, A, B)
    assert np.array_equal(got, want), f"{name}: max diff {np.abs(got - want).max():.3e}"
    _check(oracle, got, A, B)


def test_wgsl_ports_at_reference_shape(gpu_ctx, oracle):
    """1024^3 (src/gemm.rs:5-7), the shape every reference test runs: BASELINE config 0 through gemm.wgsl."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200.workload import entry_workload
    M = N = K = 1024
    A = oracle.generate_weight_data(3, M, K)
    B = oracle.generate_weight_data(4, K, N)
    for name in ("gemm_wonnx", "gemm_5", "bram"):
        wl, kid = entry_workload(name, M, N, K)
        prm = w.KernelParams(workgroup_size=(wl.size.x, wl.size.y, wl.size.z))
        got = _run(gpu_ctx, kid, A, B, M, N, K, prm, grid=(wl.count.x, wl.count.y, wl.count.z))
        assert np.array_equal(got, oracle.wgsl_gemm(name, A, B)), name
        _check(oracle, got, A, B)


@pytest.mark.parametrize("shape", [(128, 128, 16), (128, 128, 128), (256, 384, 512), (1024, 1024, 1024)])
def test_sgemm_simt_aligned(gpu_ctx, oracle, shape):
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(5, M, K)
    B = oracle.generate_weight_data(6, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_SIMT, A, B, M, N, K)  # default schedule: leftover tiles are split along K
    _check(oracle, got, A, B)
    again = _run(gpu_ctx, w.KernelId.SGEMM_SIMT, A, B, M, N, K)
    assert np.array_equal(got, again), "split-K parts are added in a fixed order: results must be deterministic"
    # B200MM_F_SEQUENTIAL_K: k-sequential fma per output == gemm_5.wgsl's order == oracle_wgsl_gemm_3 (same arithmetic)
    seq = _run(gpu_ctx, w.KernelId.SGEMM_SIMT, A, B, M, N, K, w.KernelParams(flags=int(w.Flags.SEQUENTIAL_K)))
    assert np.array_equal(seq, oracle.wgsl_gemm("gemm_3", A, B))


@pytest.mark.parametrize("shape", [(1, 4, 4), (5, 7, 3), (130, 257, 45), (127, 129, 17), (300, 100, 1000)])
def test_sgemm_simt_ragged(gpu_ctx, oracle, shape):
    """Edge tiles: arbitrary M, N, K through the guarded instantiation."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(7, M, K)
    B = oracle.generate_weight_data(8, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_SIMT, A, B, M, N, K)
    _check(oracle, got, A, B)


@pytest.mark.parametrize("bn", [256, 128])
@pytest.mark.parametrize("shape", [(128, 256, 32), (128, 256, 256), (256, 512, 1024), (1024, 1024, 1024), (384, 768, 96)])
def test_sgemm_tc3x(gpu_ctx, oracle, shape, bn):
    """tcgen05 3xTF32: FP32-accurate (passes the reference gate and the FP64 relative bound)."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(9, M, K)
    B = oracle.generate_weight_data(10, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(bn, 0, 0, 0)))
    _check(oracle, got, A, B)


@pytest.mark.parametrize("shape", [(256, 256, 256), (512, 768, 1024), (300, 520, 260), (1024, 1024, 1024), (130, 260, 36), (2304, 2048, 768), (257, 1001, 515)])
def test_sgemm_tc3x_cta_pairs(gpu_ctx, oracle, shape):
    """The 2-CTA instantiation (tune[0] = 512: 256 x 256 tiles on CTA pairs, tcgen05 cta_group::2), forced on shapes the default
    rule would give to single CTAs: ragged edges (a pair whose second CTA is entirely out of range), stream-K tails, and the
    padded path.  Same gates as the 1-CTA kernel, and bit-identical to it where the K-split schedule coincides."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(19, M, K)
    B = oracle.generate_weight_data(20, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(512, 0, 0, 0)))
    assert not (got == 123.25).any()
    _check(oracle, got, A, B)
    again = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(512, 0, 0, 0)))
    assert np.array_equal(got, again)
    # the st.global epilogue (tune[2] = 6) and the TMA-store epilogue (default) move the same values
    stg = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(512, 0, 6, 0)))
    assert np.array_equal(got, stg)
    if K <= 256:  # one chain per tile: no K-split anywhere, so the two kernels perform the same arithmetic
        single = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(513, 0, 0, 0)))
        assert np.array_equal(got, single)


@pytest.mark.parametrize("force", [512, 513])
@pytest.mark.parametrize("shape", [(256, 256, 256), (512, 768, 1024), (300, 520, 260), (1024, 1024, 1024), (2304, 2048, 768), (128, 4096, 2048),
                                   (4096, 4096, 512), (2560, 4096, 4096)])
def test_sgemm_tc3x_lo_tiles_computed_in_shared_memory(gpu_ctx, oracle, shape, force):
    """Tc3xCfg::SPLIT: the epilogue warps derive B_lo (tune[3] = 1, 4) or A_lo and B_lo (5: no pre-pass at all) from the landed
    tiles inside the GEMM; 2 / 3 take the lo operands from the split_lo pre-pass (2: A by row bands, 3: round 1), 0 picks by shape.
    Same lo values, same MMA order -> all forms are bit-identical, on CTA pairs (512) and single CTAs (513), with stream-K
    tails, k-splits, ragged edges and several waves per CTA."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(23, M, K)
    B = oracle.generate_weight_data(24, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(force, 0, 0, 5)))
    assert not (got == 123.25).any()
    rows = np.array(sorted({0, 1, 127, min(128, M - 1), M // 2 + 3, M - 1}))
    e, m = oracle.err_vs_f64(got[rows], oracle.mm_f64_rows(A, B, rows))
    assert e / m <= REL_F64
    for t3 in (0, 1, 2, 3, 4):
        other = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(force, 0, 0, t3)))
        assert np.array_equal(got, other), f"tune[3] = {t3}"
    again = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(tune=(force, 0, 0, 5)))
    assert np.array_equal(got, again)


def test_sgemm_tc3x_cta_pairs_at_the_baseline_shape(gpu_ctx, oracle):
    """4096^3 takes the pair kernel by default (512 pair tiles >= 74 SM pairs): sampled rows vs FP64 / mm_ref, checksum of all tiles."""
    import wgpu_mm_b200 as w
    M = N = K = 4096
    A = oracle.generate_weight_data(21, M, K)
    B = oracle.generate_weight_data(22, K, N)
    kern = gpu_ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
    grid, block = kern.geometry()
    assert grid[0] == 148 and block[0] == 384
    kern.free()
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    rows = np.array(sorted({0, 127, 128, 255, 256, 2051, 4095}))
    e, m = oracle.err_vs_f64(got[rows], oracle.mm_f64_rows(A, B, rows))
    assert e / m <= REL_F64
    assert oracle.max_abs_err(got[rows], oracle.mm_ref(A[rows], B)) <= GATE
    cs = A.astype(np.float64).sum(axis=0) @ B.astype(np.float64)
    assert np.abs(got.astype(np.float64).sum(axis=0) - cs).max() <= 1e-6 * M * np.abs(cs).max() + 1e-3


@pytest.mark.parametrize("shape", [(100, 36, 20), (129, 260, 36), (500, 1000, 252), (128, 4, 4), (128, 130, 66), (48, 4096, 4096), (1, 7, 5),
                                   (100, 37, 21), (257, 1001, 515), (300, 64, 1027), (13, 14336, 4096)])
def test_sgemm_tc3x_ragged(gpu_ctx, oracle, shape):
    """TMA zero-fills out-of-range rows / k and the epilogue guards the stores; N or K not a multiple of 4 goes through
    zero-padded staging copies (arbitrary M, N, K on the tensor-core path)."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(11, M, K)
    B = oracle.generate_weight_data(12, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    _check(oracle, got, A, B)


@pytest.mark.parametrize("shape", [(4096, 4096, 512), (4096, 14336, 512), (19000, 256, 512), (4224, 4096, 768), (2048, 2048, 256),
                                   (1792, 1792, 1792), (2304, 2304, 2304), (2560, 2048, 2560)])  # the last three: L2-resident stream-K / two-tile schedules
def test_sgemm_tc3x_stream_k_tail_with_idle_ctas(gpu_ctx, oracle, shape):
    """Shapes whose stream-K tail has fewer units than CTAs (some CTAs get an empty range) or exactly one chain per tile:
    the round-1 fix-up spun forever on a CTA that never publishes (ADVICE r1, 4096 x 4096 x 512).  Sampled rows vs FP64
    and mm_ref; two launches must agree bit for bit (fixed summation order)."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(13, M, K)
    B = oracle.generate_weight_data(14, K, N)
    rows = np.array(sorted({0, 1, 127, 128, M // 2 + 3, M - 129, M - 1}))
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    assert not np.isnan(got).any() and not (got == 123.25).any()
    e, m = oracle.err_vs_f64(got[rows], oracle.mm_f64_rows(A, B, rows))
    assert e / m <= REL_F64
    assert oracle.max_abs_err(got[rows], oracle.mm_ref(A[rows], B)) <= GATE
    # checksum of checksums over ALL tiles: column sums of C == (column sums of A) * B, in FP64
    cs = A.astype(np.float64).sum(axis=0) @ B.astype(np.float64)
    assert np.abs(got.astype(np.float64).sum(axis=0) - cs).max() <= 1e-6 * M * np.abs(cs).max() + 1e-3
    again = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    assert np.array_equal(got, again)


def _random_tc3x_shapes(n, seed=2024):
    """Seeded shapes across the regimes the default rules of setup_tc3x distinguish: 128- vs 256-column tiles, single CTAs vs pairs,
    k-split vs stream-K vs whole waves, B_lo from the pre-pass vs computed in shared memory, ragged edges, the padded path."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        kind = i % 6
        if kind == 0:    # small and squarish: 128 x 128 tiles
            M, N, K = rng.integers(1, 1536), 4 * rng.integers(1, 384), 4 * rng.integers(1, 384)
        elif kind == 1:  # skinny M against a big B: B_lo computed in shared memory
            M, N, K = rng.integers(1, 257), 4 * rng.integers(512, 1025), 4 * rng.integers(1024, 1537)
        elif kind == 2:  # L2-resident mid sizes: stream-K, pairs from 48 pair tiles on
            M, N, K = 4 * rng.integers(400, 640), 4 * rng.integers(400, 640), 4 * rng.integers(64, 512)
        elif kind == 3:  # more than one wave of pair tiles
            M, N, K = rng.integers(2500, 4200), 4 * rng.integers(700, 1050), 4 * rng.integers(32, 160)
        elif kind == 4:  # N or K not a multiple of 4: zero-padded staging copies
            M, N, K = rng.integers(1, 700), rng.integers(1, 900), rng.integers(1, 900)
        else:            # tall and narrow
            M, N, K = rng.integers(3000, 9000), 4 * rng.integers(1, 64), 4 * rng.integers(16, 256)
        out.append((int(M), int(N), int(K)))
    return out


@pytest.mark.parametrize("shape", _random_tc3x_shapes(36))
def test_sgemm_tc3x_default_rules_on_random_shapes(gpu_ctx, oracle, shape):
    """Whatever instantiation and schedule the default rules pick: sampled rows against FP64 and mm_ref, the FP64 checksum of
    checksums over ALL tiles, no unwritten element, and two launches bit-identical."""
    import wgpu_mm_b200 as w
    M, N, K = shape
    A = oracle.generate_weight_data(61, M, K)
    B = oracle.generate_weight_data(62, K, N)
    got = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    assert not np.isnan(got).any() and not (got == 123.25).any()
    rows = np.array(sorted({0, M // 3, M // 2, max(0, M - 129), M - 1}))
    e, m = oracle.err_vs_f64(got[rows], oracle.mm_f64_rows(A, B, rows))
    assert e / max(m, 1e-30) <= REL_F64
    assert oracle.max_abs_err(got[rows], oracle.mm_ref(A[rows], B)) <= GATE
    cs = A.astype(np.float64).sum(axis=0) @ B.astype(np.float64)
    assert np.abs(got.astype(np.float64).sum(axis=0) - cs).max() <= 1e-6 * M * np.abs(cs).max() + 1e-3
    again = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    assert np.array_equal(got, again)


@pytest.mark.parametrize("kid_name", ["SGEMM_TC3X", "SGEMM_SIMT"])
def test_split_k_fixup_survives_a_busy_device(gpu_ctx, oracle, kid_name):
    """The split-K / stream-K finisher only waits on lower-numbered CTAs, so it cannot deadlock when the grid is not fully
    co-resident: run 1024^3 (K-split over idle SMs in both kernels) while a second context keeps every SM busy."""
    import wgpu_mm_b200 as w
    kid = getattr(w.KernelId, kid_name)
    M = N = K = 1024
    A = oracle.generate_weight_data(15, M, K)
    B = oracle.generate_weight_data(16, K, N)
    quiet = _run(gpu_ctx, kid, A, B, M, N, K)
    _check(oracle, quiet, A, B)
    other = w.Context(0)
    S = 2048
    oa, ob, oc = other.buffer(S * S * 4), other.buffer(S * S * 4), other.buffer(S * S * 4)
    oa.fill_weights(1, S * S); ob.fill_weights(2, S * S)
    hog = other.kernel(w.KernelId.SGEMM_SIMT, S, S, S)
    kern = gpu_ctx.kernel(kid, M, N, K)
    dA, dB, dC = gpu_ctx.buffer_from(A), gpu_ctx.buffer_from(B), gpu_ctx.buffer_from(np.full(M * N, 123.25, dtype=np.float32))
    for _ in range(40):
        other.launch(hog, oa, ob, oc)  # ~0.3 ms each on its own stream: 512 CTAs, 2 per SM
    for _ in range(10):
        gpu_ctx.launch(kern, dA, dB, dC)
    busy = dC.read(np.float32).reshape(M, N)
    other.sync()
    assert np.array_equal(busy, quiet)
    for b in (oa, ob, oc, dA, dB, dC):
        b.free()
    hog.free(); kern.free(); other.close()


def test_sgemm_tc3x_const_b_reuses_the_split_only_while_b_stays(gpu_ctx, oracle):
    """B200MM_F_CONST_B (weights): B's tf32 lo part is computed on the first launch with a given B buffer and reused while the
    pointer stays the same; a different B buffer must be split again.  Results are bit-identical to the unflagged kernel."""
    import wgpu_mm_b200 as w
    M, N, K = 256, 512, 384
    A1, A2 = oracle.generate_weight_data(31, M, K), oracle.generate_weight_data(32, M, K)
    B1, B2 = oracle.generate_weight_data(33, K, N), oracle.generate_weight_data(34, K, N)
    plain = gpu_ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
    const = gpu_ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(flags=int(w.Flags.CONST_B)))
    dA1, dA2, dB1, dB2 = (gpu_ctx.buffer_from(x) for x in (A1, A2, B1, B2))
    dC, dR = gpu_ctx.buffer(M * N * 4), gpu_ctx.buffer(M * N * 4)
    for dA, dB, A, B in ((dA1, dB1, A1, B1), (dA2, dB1, A2, B1), (dA1, dB2, A1, B2), (dA2, dB2, A2, B2), (dA2, dB1, A2, B1)):
        gpu_ctx.launch(const, dA, dB, dC)
        gpu_ctx.launch(plain, dA, dB, dR)
        got, ref = dC.read(np.float32).reshape(M, N), dR.read(np.float32).reshape(M, N)
        assert np.array_equal(got, ref)
        _check(oracle, got, A, B)
    for b in (dA1, dA2, dB1, dB2, dC, dR):
        b.free()
    plain.free(); const.free()


def test_sgemm_tc3x_padded_path_is_deterministic_and_leaves_neighbours_alone(gpu_ctx, oracle):
    """N % 4 != 0: C has an odd pitch; the copy-back must write exactly M x N floats (canary behind C) and repeat bit for bit."""
    import wgpu_mm_b200 as w
    M, N, K = 70, 130, 66
    A = oracle.generate_weight_data(17, M, K)
    B = oracle.generate_weight_data(18, K, N)
    kern = gpu_ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
    dA, dB = gpu_ctx.buffer_from(A), gpu_ctx.buffer_from(B)
    dC = gpu_ctx.buffer_from(np.full(M * N + 64, 7.0, dtype=np.float32))
    gpu_ctx.launch(kern, dA, dB, dC)
    first = dC.read(np.float32)
    gpu_ctx.launch(kern, dA, dB, dC)
    second = dC.read(np.float32)
    assert (first[M * N:] == 7.0).all() and np.array_equal(first, second)
    _check(oracle, first[:M * N].reshape(M, N), A, B)
    for b in (dA, dB, dC):
        b.free()
    kern.free()


def test_sgemm_tc3x_single_pass_fails_the_gate_at_large_k(gpu_ctx, oracle):
    """1xTF32 cannot hold the reference's 1e-3 gate -- this is why the split exists."""
    import wgpu_mm_b200 as w
    M, N, K = 128, 256, 4096
    A = oracle.generate_weight_data(13, M, K)
    B = oracle.generate_weight_data(14, K, N)
    got1 = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K, w.KernelParams(flags=int(w.Flags.TC3X_1X)))
    got3 = _run(gpu_ctx, w.KernelId.SGEMM_TC3X, A, B, M, N, K)
    ref = oracle.mm_f64(A, B)
    e1 = np.abs(got1 - ref).max()
    e3 = np.abs(got3 - ref).max()
    assert e3 < 2e-5 and e1 > 20 * e3


def test_sgemm_linearity_full_size(gpu_ctx):
    """4096^3 (BASELINE config 1) through size-independent properties: a rank-1 structured product has a
    closed form, and C(A, B) restricted to sampled rows equals an FP64 GEMM of those rows."""
    import wgpu_mm_b200 as w
    import oracle
    M = N = K = 4096
    A = oracle.generate_weight_data(15, M, K)
    B = oracle.generate_weight_data(16, K, N)
    rows = np.array([0, 1, 127, 128, 2047, 4095])
    ref = oracle.mm_f64_rows(A, B, rows)
    for kid in (w.KernelId.SGEMM_TC3X, w.KernelId.SGEMM_SIMT):
        got = _run(gpu_ctx, kid, A, B, M, N, K)
        e, m = oracle.err_vs_f64(got[rows], ref)
        assert e / m <= REL_F64, (kid, e / m)
        # checksum of checksums: column sums of C == (column sums of A-rows) applied to B
        cs = got.astype(np.float64).sum(axis=0)
        want = A.astype(np.float64).sum(axis=0) @ B.astype(np.float64)
        assert np.abs(cs - want).max() / np.abs(want).max() < 1e-5


GEMV_SHAPES = [(64, 64), (512, 1024), (1000, 260), (4096, 16384), (36, 4)]


@pytest.mark.parametrize("variant", [0, 100, 1, 2, 3, 4, 6, 7])
@pytest.mark.parametrize("kn", GEMV_SHAPES)
def test_gemv_f32(gpu_ctx, oracle, kn, variant):
    """fp32 GEMV vs mm_ref with M == 1 (the reference has no fp32 GEMV shader)."""
    import wgpu_mm_b200 as w
    K, N = kn
    x = oracle.generate_weight_data(17, 1, K)
    W = oracle.generate_weight_data(18, K, N)
    got = _run(gpu_ctx, w.KernelId.GEMV_F32, x, W, 1, N, K, w.KernelParams(tune=(variant, 0, 0, 0)))
    _check(oracle, got, x, W)


@pytest.mark.parametrize("cluster_off", [0, 1])
@pytest.mark.parametrize("splits", [1, 2, 7, 8, 64])
def test_gemv_f32_split_k_is_deterministic(gpu_ctx, oracle, splits, cluster_off):
    import wgpu_mm_b200 as w
    K, N = 2048, 1024
    x = oracle.generate_weight_data(19, 1, K)
    W = oracle.generate_weight_data(20, K, N)
    # splits <= 8 are reduced inside a thread-block cluster (DSMEM); cluster_off = 1 or splits > 8 use the ticket path
    prm = w.KernelParams(tune=(0, splits, 0, cluster_off))
    a = _run(gpu_ctx, w.KernelId.GEMV_F32, x, W, 1, N, K, prm)
    b = _run(gpu_ctx, w.KernelId.GEMV_F32, x, W, 1, N, K, prm)
    assert np.array_equal(a, b)
    _check(oracle, a, x, W)


QSHAPES = [(1024, 1024), (64, 64), (4096, 14336), (200, 48), (36, 16)]


@pytest.mark.parametrize("variant", [0, 100, 1, 4, 5, 6, 11, 12, 13, 21, 25])
@pytest.mark.parametrize("kn", QSHAPES)
def test_qgemv_sint8(gpu_ctx, oracle, kn, variant):
    """Quantised GEMV in the src/quant.rs format with the reference's ABSMAX = 2.0 quirk."""
    import wgpu_mm_b200 as w
    K, N = kn
    x = oracle.generate_weight_data(21, 1, K)
    W = oracle.generate_weight_data(22, K, N)
    words, _absmax = oracle.sint8_quantize(W, K, N)  # true absmax discarded, like src/harness.rs:134
    prm = w.KernelParams(absmax=2.0, batch=1, tune=(variant, 0, 0, 0))
    got = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, x, words, 1, N, K, prm, b_dtype=np.uint32)
    ref = oracle.qgemv_ref(x, words, 1, N, K, 2.0)
    assert oracle.max_abs_err(got, ref) <= GATE
    e, m = oracle.err_vs_f64(got, oracle.qgemv_f64(x, words, 1, N, K, 2.0))
    assert e / m <= REL_F64
    # agreement with the WGSL-order restatement (blocked-by-4 dot): tolerance only, WGSL leaves the order open
    assert oracle.max_abs_err(got, oracle.wgsl_qgemv_1(x, words, N, K, 2.0)) <= 1e-4


@pytest.mark.parametrize("case", [("s8", 4096, 14336, 21, 2, 148), ("s8", 4096, 14336, 21, 1, 296), ("s8", 1024, 1024, 21, 2, 40), ("s8", 1000, 1040, 13, 2, 17),
                                  ("s8", 512, 4096, 21, 4, 74), ("f32", 1024, 2048, 5, 2, 20), ("f32", 4096, 16384, 5, 4, 148), ("f32", 300, 260, 4, 1, 20)])
def test_gemv_balanced_ragged_panels(gpu_ctx, oracle, case):
    """tune[3] = explicit panel count: the column groups are dealt evenly to that many (ragged) panels so that a grid can be sized
    to exactly one CTA slot per panel x split.  Same results as the natural partition, bit for bit per column (the row order of
    each column's sum does not depend on which panel owns it) as long as the K-split is the same."""
    import wgpu_mm_b200 as w
    kind, K, N, variant, splits, panels = case
    x = oracle.generate_weight_data(27, 1, K)
    W = oracle.generate_weight_data(28, K, N)
    if kind == "s8":
        words, _ = oracle.sint8_quantize(W, K, N)
        kid, B, dt = w.KernelId.QGEMV_SINT8, words, np.uint32
        want, f64 = oracle.qgemv_ref(x, words, 1, N, K, 2.0), oracle.qgemv_f64(x, words, 1, N, K, 2.0)
    else:
        kid, B, dt = w.KernelId.GEMV_F32, W, np.float32
        want, f64 = oracle.mm_ref(x, W), oracle.mm_f64(x, W)
    got = _run(gpu_ctx, kid, x, B, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(variant, splits, 0, panels)), b_dtype=dt)
    assert not (got == 123.25).any()
    assert oracle.max_abs_err(got, want) <= GATE
    e, m = oracle.err_vs_f64(got, f64)
    assert e / m <= REL_F64
    kern = gpu_ctx.kernel(kid, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(variant, splits, 0, panels)))
    assert kern.geometry()[0][0] == panels
    kern.free()
    natural = _run(gpu_ctx, kid, x, B, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(variant, splits, 0, 0)), b_dtype=dt)
    assert np.array_equal(got, natural)


def test_peer_flags_need_peers_first(gpu_ctx):
    """b200mm_kernel_set_peer_flags is only meaningful on a GEMV kernel that already has its peers (world >= 2)."""
    import ctypes as C
    import wgpu_mm_b200 as w
    kern = gpu_ctx.kernel(w.KernelId.GEMV_F32, 1, 1024, 1024)
    flags = gpu_ctx.buffer(64)
    with pytest.raises(w.B200mmError):
        kern.set_peer_flags([flags.ptr, flags.ptr], 0)
    with pytest.raises(w.B200mmError):
        kern.peer_wait()
    gemm = gpu_ctx.kernel(w.KernelId.SGEMM_SIMT, 128, 128, 128)
    with pytest.raises(w.B200mmError):
        gemm.set_peer_flags([flags.ptr, flags.ptr], 0)
    assert kern.peer_epoch == 0
    for k in (kern, gemm):
        k.free()
    flags.free()


def test_gemv_panel_count_must_fit_the_instantiation(gpu_ctx):
    import wgpu_mm_b200 as w
    with pytest.raises(w.B200mmError):  # 896 column groups on 16 panels = 56 per panel, the 128-column instantiation holds 8
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, 14336, 4096, w.KernelParams(absmax=2.0, batch=1, tune=(21, 2, 0, 16)))
    with pytest.raises(w.B200mmError):  # more panels than column groups
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, 256, 4096, w.KernelParams(absmax=2.0, batch=1, tune=(21, 2, 0, 17)))


def test_qgemv_true_absmax_and_extremes(gpu_ctx, oracle):
    """Sanity variant with the real absmax, and weights at +-127 (the int8 extremes the codec can emit)."""
    import wgpu_mm_b200 as w
    K, N = 256, 512
    x = oracle.generate_weight_data(23, 1, K)
    W = oracle.generate_weight_data(24, K, N)
    W[0, :] = 0.2
    W[1, :] = -0.2
    words, absmax = oracle.sint8_quantize(W, K, N)
    q = words.view(np.int8)
    assert q.max() == 127 and q.min() == -127
    prm = w.KernelParams(absmax=absmax, batch=1)
    got = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, x, words, 1, N, K, prm, b_dtype=np.uint32)
    e, m = oracle.err_vs_f64(got, oracle.qgemv_f64(x, words, 1, N, K, absmax))
    assert e / m <= REL_F64
    # dequantised GEMV approximates the unquantised one to quantisation accuracy
    full = oracle.mm_f64(x, W)
    assert np.abs(got - full).max() < 0.05


@pytest.mark.parametrize("kid_name", ["QGEMV_SINT8", "QGEMV_1"])
def test_qgemv_batched(gpu_ctx, oracle, kid_name):
    """global_id.y batch offsets (qgemv_1.wgsl:12-14) that the reference harness never dispatches."""
    import wgpu_mm_b200 as w
    K, N, batch = 256, 512, 3
    x = oracle.generate_weight_data(25, batch, K)
    Ws = [oracle.sint8_quantize(oracle.generate_weight_data(30 + b, K, N), K, N)[0] for b in range(batch)]
    Bq = np.concatenate(Ws)
    kid = getattr(w.KernelId, kid_name)
    prm = w.KernelParams(absmax=2.0, batch=batch, workgroup_size=(8, 1, 1) if kid_name == "QGEMV_1" else (0, 0, 0))
    got = _run(gpu_ctx, kid, x, Bq, 1, N, K, prm, b_dtype=np.uint32)
    want = oracle.wgsl_qgemv_1(x, Bq, N, K, 2.0, batch=batch)
    if kid_name == "QGEMV_1":
        assert np.array_equal(got, want)  # faithful port: bit-exact vs its restatement
    else:
        assert oracle.max_abs_err(got, want) <= 1e-4
    for b in range(batch):
        assert oracle.max_abs_err(got[b:b + 1], oracle.qgemv_ref(x[b:b + 1], Ws[b], 1, N, K, 2.0)) <= GATE


def test_qgemv_1_port_batch_guard(gpu_ctx, oracle):
    """workgroup_size_y = 2 with batch = 3 dispatches 4 rows of invocations; the 4th must not touch memory (guard canary)."""
    import wgpu_mm_b200 as w
    K, N, batch = 64, 64, 3
    x = oracle.generate_weight_data(26, batch, K)
    Bq = np.concatenate([oracle.sint8_quantize(oracle.generate_weight_data(40 + b, K, N), K, N)[0] for b in range(batch)])
    kern = gpu_ctx.kernel(w.KernelId.QGEMV_1, 1, N, K, w.KernelParams(absmax=2.0, batch=batch, workgroup_size=(8, 2, 1)))
    dx, dB = gpu_ctx.buffer_from(x), gpu_ctx.buffer_from(Bq)
    dy = gpu_ctx.buffer_from(np.full((batch + 1) * N, 7.0, dtype=np.float32))  # one canary row behind the batch
    gpu_ctx.launch(kern, dx, dB, dy)
    got = dy.read(np.float32).reshape(batch + 1, N)
    assert np.array_equal(got[:batch], oracle.wgsl_qgemv_1(x, Bq, N, K, 2.0, batch=batch))
    assert (got[batch] == 7.0).all()
    for b in (dx, dB, dy):
        b.free()
    kern.free()


GROUPED = [(1024, 1024, 128), (4096, 14336, 128), (4096, 14336, 256), (1000, 1040, 128), (64, 64, 128), (2048, 512, 2048),
           (1536, 256, 512), (4096, 4096, 4096),
           # GGUF-style small groups (round 2): a narrower pipeline window (UNROLL 2 / 1) so that it never straddles a group
           (1024, 1024, 64), (1024, 1024, 32), (4096, 14336, 64), (4096, 14336, 32), (1000, 1040, 32), (96, 48, 64), (40, 16, 32)]


@pytest.mark.parametrize("kng", GROUPED)
def test_qgemv_sint8_grouped_scales(gpu_ctx, oracle, kng):
    """Per-(row block, column) scales: B = weight words ++ f32 scales, params.group_k set.
    Checked against mm_ref / FP64 over the oracle's group-dequantised weights; ragged last group, K < group_k, panels
    that are not full (N % 256 != 0), one group per column (group_k >= K) included."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200.quant import sint8_quantize_grouped, split_grouped
    K, N, G = kng
    x = oracle.generate_weight_data(61, 1, K)
    W = oracle.generate_weight_data(62, K, N)
    W *= (1.0 + 7.0 * (np.arange(K, dtype=np.float32)[:, None] // G % 3)) * (1.0 + (np.arange(N, dtype=np.float32)[None, :] % 5))
    packed = sint8_quantize_grouped(W, K, N, G)
    words, scales = split_grouped(packed, K, N, G)
    owords, oscales = oracle.sint8_quantize_grouped(W, K, N, G)  # product codec == oracle codec, bit for bit
    assert np.array_equal(words, owords) and np.array_equal(scales, oscales)
    prm = w.KernelParams(batch=1, group_k=G)
    got = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, x, packed, 1, N, K, prm, b_dtype=np.uint32)
    assert not np.isnan(got).any()
    e, m = oracle.err_vs_f64(got, oracle.qgemv_grouped_f64(x, words, scales, 1, N, K, G))
    assert e / m <= REL_F64, f"rel err vs fp64 {e / m:.3e}"
    ref = oracle.qgemv_grouped_ref(x, words, scales, 1, N, K, G)
    assert oracle.max_abs_err(got, ref) <= GATE * max(1.0, m)  # the blown-up weights make |y| > 1: gate relative to max |y|


def test_qgemv_grouped_beats_global_scale_on_outliers(gpu_ctx, oracle):
    """Why the format exists: one outlier weight ruins the global-absmax codec (src/quant.rs:17) for every column."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200.quant import sint8_quantize_grouped
    K, N, G = 1024, 1024, 128
    x = oracle.generate_weight_data(63, 1, K)
    W = oracle.generate_weight_data(64, K, N)
    W[5, 7] = 40.0
    full = oracle.mm_f64(x, W)
    words, absmax = oracle.sint8_quantize(W, K, N)
    y_global = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, x, words, 1, N, K, w.KernelParams(absmax=absmax, batch=1), b_dtype=np.uint32)
    y_group = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, x, sint8_quantize_grouped(W, K, N, G), 1, N, K,
                   w.KernelParams(batch=1, group_k=G), b_dtype=np.uint32)
    others = np.arange(N) != 7  # the damage is confined to the outlier's own (row block, column)
    err_global = np.abs(y_global - full)[0, others].max()
    err_group = np.abs(y_group - full)[0, others].max()
    assert err_group < 0.02 and err_group * 10 < err_global, (err_group, err_global)
    assert np.abs(y_group - full)[0, 7] <= np.abs(y_global - full).max() * 1.5


def test_qgemv_grouped_batched_and_errors(gpu_ctx, oracle):
    import wgpu_mm_b200 as w
    from wgpu_mm_b200.quant import sint8_quantize_grouped, split_grouped
    K, N, G, batch = 512, 512, 128, 2
    x = oracle.generate_weight_data(65, batch, K)
    packs = [sint8_quantize_grouped(oracle.generate_weight_data(70 + b, K, N), K, N, G) for b in range(batch)]
    got = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, x, np.concatenate(packs), 1, N, K, w.KernelParams(batch=batch, group_k=G), b_dtype=np.uint32)
    for b in range(batch):
        words, scales = split_grouped(packs[b], K, N, G)
        e, m = oracle.err_vs_f64(got[b:b + 1], oracle.qgemv_grouped_f64(x[b:b + 1], words, scales, 1, N, K, G))
        assert e / m <= REL_F64
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, N, K, w.KernelParams(group_k=96))  # not 32, 64 or a multiple of 128
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 2, N, K, w.KernelParams(group_k=128, batch=2))  # batched needs M == 1
    # M > 1 with per-group scales: one pass per row of x (the grouped kernel is single-row)
    X3 = oracle.generate_weight_data(66, 3, K)
    got3 = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, X3, packs[0], 3, N, K, w.KernelParams(batch=1, group_k=G), b_dtype=np.uint32)
    words0, scales0 = split_grouped(packs[0], K, N, G)
    e3, m3 = oracle.err_vs_f64(got3, oracle.qgemv_grouped_f64(X3, words0, scales0, 3, N, K, G))
    assert e3 / m3 <= REL_F64
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.GEMV_F32, 1, N, K, w.KernelParams(group_k=128))  # fp32 weights carry no scales
    kern = gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, N, K, w.KernelParams(group_k=G))
    dA = gpu_ctx.buffer_from(x[:1].copy())
    dB = gpu_ctx.buffer_from(split_grouped(packs[0], K, N, G)[0].copy())  # weights without their scales: too short
    dC = gpu_ctx.buffer_from(np.zeros(N, dtype=np.float32))
    with pytest.raises(w.B200mmError):
        gpu_ctx.launch(kern, dA, dB, dC)
    for b in (dA, dB, dC):
        b.free()
    kern.free()


@pytest.mark.parametrize("case", [("f32", 1024, 2048, 0), ("s8", 4096, 14336, 0), ("s8", 1024, 1024, 128), ("s8", 200, 48, 0)])
def test_gemv_autotune_keeps_parity(gpu_ctx, oracle, case):
    """B200MM_F_AUTOTUNE picks the geometry / K-split count by measurement at kernel creation; whatever it picks must
    pass the same gates, launch after launch (the choice is fixed per kernel object)."""
    import wgpu_mm_b200 as w
    from wgpu_mm_b200.quant import sint8_quantize_grouped, split_grouped
    kind, K, N, G = case
    x = oracle.generate_weight_data(81, 1, K)
    W = oracle.generate_weight_data(82, K, N)
    flags = int(w.Flags.AUTOTUNE)
    before = gpu_ctx.launch_count
    if kind == "f32":
        kern = gpu_ctx.kernel(w.KernelId.GEMV_F32, 1, N, K, w.KernelParams(flags=flags))
        B, want64 = W, oracle.mm_f64(x, W)
    elif G:
        kern = gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, N, K, w.KernelParams(flags=flags, group_k=G))
        B = sint8_quantize_grouped(W, K, N, G)
        words, scales = split_grouped(B, K, N, G)
        want64 = oracle.qgemv_grouped_f64(x, words, scales, 1, N, K, G)
    else:
        kern = gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, N, K, w.KernelParams(flags=flags, absmax=2.0))
        B, _ = oracle.sint8_quantize(W, K, N)
        want64 = oracle.qgemv_f64(x, B, 1, N, K, 2.0)
    assert gpu_ctx.launch_count == before  # tuning launches are not counted as user launches
    dA, dB, dC = gpu_ctx.buffer_from(x), gpu_ctx.buffer_from(np.ascontiguousarray(B)), gpu_ctx.buffer_from(np.full(N, 7.5, dtype=np.float32))
    outs = []
    for _ in range(3):
        gpu_ctx.launch(kern, dA, dB, dC)
        outs.append(dC.read(np.float32).reshape(1, N))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    e, m = oracle.err_vs_f64(outs[0], want64)
    assert e / m <= REL_F64, f"rel err vs fp64 {e / m:.3e}"
    grid, block = kern.geometry()
    assert 1 <= grid[1] <= 8
    for b in (dA, dB, dC):
        b.free()
    kern.free()


def test_device_datagen_matches_oracle(gpu_ctx, oracle):
    n = 1 << 16
    buf = gpu_ctx.buffer(n * 4)
    for seed, off in ((1, 0), (0x5EED, 12345), (7, 1 << 33)):
        buf.fill_weights(seed, n, off)
        got = buf.read(np.float32)
        want = oracle.generate_weight_data(seed, 1, n, offset=off).reshape(-1)
        assert np.array_equal(got, want)
    buf.free()


def test_unshard_columns(gpu_ctx):
    M, N, world = 64, 256, 4
    full = np.arange(M * N, dtype=np.float32).reshape(M, N)
    panels = np.concatenate([full[:, r * (N // world):(r + 1) * (N // world)].reshape(-1) for r in range(world)])
    g = gpu_ctx.buffer_from(panels)
    c = gpu_ctx.buffer(M * N * 4)
    gpu_ctx.unshard_columns(g.ptr, c.ptr, M, N, world)
    assert np.array_equal(c.read(np.float32).reshape(M, N), full)
    g.free()
    c.free()


def test_launch_rejects_undersized_buffers(gpu_ctx):
    import wgpu_mm_b200 as w
    kern = gpu_ctx.kernel(w.KernelId.SGEMM_SIMT, 128, 128, 128)
    small = gpu_ctx.buffer(16)
    ok = gpu_ctx.buffer(128 * 128 * 4)
    with pytest.raises(w.B200mmError):
        gpu_ctx.launch(kern, small, ok, ok)
    small.free()
    ok.free()
    kern.free()


@pytest.mark.parametrize("kid_name", ["SGEMM_TC3X", "SGEMM_SIMT"])
@pytest.mark.parametrize("size", [1024, 2048])
def test_mm_host_end_to_end(gpu_ctx, oracle, kid_name, size):
    """b200mm_mm_host: host A, B -> device, GEMM, C -> host (pipelined over row panels for the SGEMM kernels)."""
    import wgpu_mm_b200 as w
    M = N = K = size
    A = oracle.generate_weight_data(41, M, K)
    B = oracle.generate_weight_data(42, K, N)
    Cm = np.full((M, N), 123.25, dtype=np.float32)
    kern = gpu_ctx.kernel(getattr(w.KernelId, kid_name), M, N, K)
    dA, dB, dC = gpu_ctx.buffer(M * K * 4), gpu_ctx.buffer(K * N * 4), gpu_ctx.buffer(M * N * 4)
    for _ in range(2):  # second call reuses the cached panel kernel and B split state
        Cm[:] = 123.25
        gpu_ctx.mm_host(kern, A, B, Cm, dA, dB, dC)
        assert not (Cm == 123.25).any()
        rows = np.array([0, 1, M // 2, M - 1])
        e, m = oracle.err_vs_f64(Cm[rows], oracle.mm_f64_rows(A, B, rows))
        assert e / m <= REL_F64
        assert oracle.max_abs_err(Cm[rows], oracle.mm_ref(A[rows], B)) <= GATE
        # different B on the second call: the panel kernel must re-split it
        B = oracle.generate_weight_data(43, K, N)
    for b in (dA, dB, dC):
        b.free()
    kern.free()


def test_mm_host_with_ragged_n_and_k_then_device_launch(gpu_ctx, oracle):
    """The pipelined host path on a shape whose kernel object owns an inner (zero-padded) kernel: N % 4 != 0, K % 4 != 0 with
    M = 2 x 256 rows.  mm_host must leave that object intact -- a device-buffer launch and the free afterwards used to hit a
    kernel that the panel set-up had freed."""
    import wgpu_mm_b200 as w
    M, N, K = 512, 1001, 515
    A = oracle.generate_weight_data(44, M, K)
    B = oracle.generate_weight_data(45, K, N)
    Cm = np.full((M, N), 123.25, dtype=np.float32)
    kern = gpu_ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
    dA, dB, dC = gpu_ctx.buffer(M * K * 4), gpu_ctx.buffer(K * N * 4), gpu_ctx.buffer(M * N * 4)
    gpu_ctx.mm_host(kern, A, B, Cm, dA, dB, dC)
    _check(oracle, Cm, A, B)
    dA2, dB2 = gpu_ctx.buffer_from(A), gpu_ctx.buffer_from(B)
    dC2 = gpu_ctx.buffer_from(np.full(M * N, 123.25, dtype=np.float32))
    gpu_ctx.launch(kern, dA2, dB2, dC2)
    got = dC2.read(np.float32).reshape(M, N)
    _check(oracle, got, A, B)  # (not bit-equal to Cm: the 256-row panels have their own k-split schedule)
    for b in (dA, dB, dC, dA2, dB2, dC2):
        b.free()
    kern.free()


@pytest.mark.parametrize("quant", [False, True])
def test_gemv_dependent_chain_with_pdl(gpu_ctx, oracle, quant):
    """x_{i+1} = y_i through the same square weight matrix, 24 launches back to back.  The GEMV kernels are launched with
    programmatic dependent launch and prefetch their weights BEFORE griddepcontrol.wait; anything that depends on the
    previous kernel (x, partials, y) must only be touched after it -- a violation shows up as a wrong chain result."""
    import wgpu_mm_b200 as w
    n = 1024
    x0 = oracle.generate_weight_data(51, 1, n)
    W = oracle.generate_weight_data(52, n, n) * np.float32(0.35)  # keeps the iterates O(0.1)
    if quant:
        words, absmax = oracle.sint8_quantize(W, n, n)
        dW = gpu_ctx.buffer_from(words)
        Wd = oracle.sint8_dequantize(words, absmax, n, n).astype(np.float64)
        kern = gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, n, n, w.KernelParams(absmax=absmax, batch=1))
    else:
        dW = gpu_ctx.buffer_from(W)
        Wd = W.astype(np.float64)
        kern = gpu_ctx.kernel(w.KernelId.GEMV_F32, 1, n, n)
    bufs = [gpu_ctx.buffer_from(x0), gpu_ctx.buffer(n * 4)]
    steps = 24
    for i in range(steps):
        gpu_ctx.launch(kern, bufs[i % 2], dW, bufs[(i + 1) % 2])
    got = bufs[steps % 2].read(np.float32).astype(np.float64)
    ref = x0.astype(np.float64).reshape(-1)
    for i in range(steps):
        ref = ref @ Wd
    scale = np.abs(ref).max()
    assert scale > 1e-12
    assert np.abs(got - ref).max() / scale < 1e-4, "dependent GEMV chain diverged: PDL ordering violated?"
    for b in bufs + [dW]:
        b.free()
    kern.free()


@pytest.mark.parametrize("m", [2, 3, 4, 5, 7, 8, 13, 16])
@pytest.mark.parametrize("kn", [(512, 1024), (4096, 4096), (1000, 260)])
def test_gemv_f32_skinny_m(gpu_ctx, oracle, m, kn):
    """Skinny GEMM: M rows of x share one pass over W; semantics = mm_ref with M rows."""
    import wgpu_mm_b200 as w
    K, N = kn
    X = oracle.generate_weight_data(61, m, K)
    W = oracle.generate_weight_data(62, K, N)
    got = _run(gpu_ctx, w.KernelId.GEMV_F32, X, W, m, N, K)
    _check(oracle, got, X, W)


@pytest.mark.parametrize("kn", [(1024, 2048), (4096, 14336)])
@pytest.mark.parametrize("m", [2, 3, 4, 6, 13, 16])
def test_qgemv_sint8_skinny_m(gpu_ctx, oracle, m, kn):
    """Every M <= 16: row counts without an instantiation of their own run in chunks of 4 / 2 / 1 rows."""
    import wgpu_mm_b200 as w
    K, N = kn
    if (K, N) == (4096, 14336) and m not in (3, 13):
        pytest.skip("BASELINE shape: the ragged row counts only")
    X = oracle.generate_weight_data(63, m, K)
    W = oracle.generate_weight_data(64, K, N)
    words, _ = oracle.sint8_quantize(W, K, N)
    got = _run(gpu_ctx, w.KernelId.QGEMV_SINT8, X, words, m, N, K, w.KernelParams(absmax=2.0, batch=1), b_dtype=np.uint32)
    ref = oracle.qgemv_ref(X, words, m, N, K, 2.0)
    assert oracle.max_abs_err(got, ref) <= GATE
    e, mx = oracle.err_vs_f64(got, oracle.qgemv_f64(X, words, m, N, K, 2.0))
    assert e / mx <= REL_F64


def test_gemv_rejects_unsupported_m(gpu_ctx):
    """The GEMV kernels stop at 16 rows of x (above that the SGEMM kernels take over); batched / grouped / peer-store launches
    keep to the natively instantiated row counts (per-group scales with M > 1 run one pass per row)."""
    import wgpu_mm_b200 as w
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.GEMV_F32, 17, 1024, 1024)
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 17, 1024, 1024, w.KernelParams(absmax=2.0))
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 3, 1024, 1024, w.KernelParams(absmax=2.0, batch=2))
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 3, 1024, 1024, w.KernelParams(group_k=128, batch=2))


@pytest.mark.parametrize("case", ["gemv_f32", "qgemv_sint8", "gemv_f32_m4", "sgemm_tc3x", "sgemm_simt"])
def test_repeatability_stress(gpu_ctx, oracle, case):
    """Race detector: the same launch repeated back to back must give bit-identical results every time.  Covers the
    cluster/DSMEM K-split reduction + PDL (GEMV), the stream-K owner/contributor fix-up with epoch flags (tc3x) and the
    split-K fix-up of small SIMT problems."""
    import wgpu_mm_b200 as w
    if case.startswith("gemv_f32"):
        M = 4 if case.endswith("m4") else 1
        K, N = 4096, 16384
        A = oracle.generate_weight_data(71, M, K)
        B = oracle.generate_weight_data(72, K, N)
        kern = gpu_ctx.kernel(w.KernelId.GEMV_F32, M, N, K)
        reps = 60
    elif case == "qgemv_sint8":
        M, K, N = 1, 4096, 14336
        A = oracle.generate_weight_data(73, M, K)
        B, _ = oracle.sint8_quantize(oracle.generate_weight_data(74, K, N), K, N)
        kern = gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, M, N, K, w.KernelParams(absmax=2.0, batch=1))
        reps = 100
    else:
        M = N = K = 1024  # 32 tiles x 4 chains: every tc3x tile is split over 4 CTAs; 64 SIMT tiles x 4 K-parts
        A = oracle.generate_weight_data(75, M, K)
        B = oracle.generate_weight_data(76, K, N)
        kern = gpu_ctx.kernel(getattr(w.KernelId, case.upper()), M, N, K)
        reps = 40
    dA, dB = gpu_ctx.buffer_from(A), gpu_ctx.buffer_from(B)
    outs = [gpu_ctx.buffer(M * N * 4) for _ in range(reps)]
    for o in outs:  # all launches are queued back to back, nothing in between
        gpu_ctx.launch(kern, dA, dB, o)
    first = outs[0].read(np.float32)
    assert np.isfinite(first).all()
    for i, o in enumerate(outs[1:], 1):
        assert np.array_equal(o.read(np.float32), first), f"launch {i} differs from launch 0"
    for b in outs + [dA, dB]:
        b.free()
    kern.free()


@pytest.mark.parametrize("kid_name", ["SGEMM_TC3X", "SGEMM_SIMT"])
def test_sgemm_long_k_accuracy(gpu_ctx, oracle, kid_name):
    """K = 16384 (the depth of BASELINE config 4): the chained TMEM accumulation must hold fp32-level accuracy."""
    import wgpu_mm_b200 as w
    M, N, K = 256, 512, 16384
    A = oracle.generate_weight_data(81, M, K)
    B = oracle.generate_weight_data(82, K, N)
    got = _run(gpu_ctx, getattr(w.KernelId, kid_name), A, B, M, N, K)
    rel = _check(oracle, got, A, B)
    assert rel <= REL_F64


def test_invalid_arguments_are_rejected(gpu_ctx):
    """Empty shapes, unknown kernels and bad launch grids fail with a status instead of undefined behaviour."""
    import wgpu_mm_b200 as w
    for dims in ((0, 64, 64), (64, 0, 64), (64, 64, 0)):
        with pytest.raises(w.B200mmError):
            gpu_ctx.kernel(w.KernelId.SGEMM_SIMT, *dims)
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(999, 64, 64, 64)
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.GEMM_5, 48, 64, 64)  # gemm_5.wgsl has no guards: needs M, N % 32 == 0
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.QGEMV_SINT8, 1, 100, 64, w.KernelParams(absmax=2.0))  # N % 16 != 0
    kern = gpu_ctx.kernel(w.KernelId.GEMM_1, 64, 64, 64, w.KernelParams(workgroup_size=(16, 16, 1)))
    buf = gpu_ctx.buffer(64 * 64 * 4)
    with pytest.raises(w.B200mmError):
        gpu_ctx.launch(kern, buf, buf, buf, grid=(0, 1, 1))  # "Compute limits exceeded"
    with pytest.raises(w.B200mmError):
        gpu_ctx.kernel(w.KernelId.GEMM_1, 64, 64, 64, w.KernelParams(workgroup_size=(64, 64, 1)))  # 4096 threads per workgroup
    buf.free()
    kern.free()
