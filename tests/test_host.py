"""CPU tests of the host side: the C ABI loads and exports what include/*.h declares, the C++ host
mirror (Workload, entry points, codec) behaves like the reference crate, and every compute call fails
loudly without a GPU (there is no CPU fallback)."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, macro):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))  # drop the macro definitions
    return re.findall(macro + r"\s+[^;(]*?\b(\w+)\s*\(", text)


def test_library_exports_every_declared_symbol(built_lib):
    names = _declared("b200mm.h", "B200MM_API") + _declared("wgpu_mm_c.h", "WGPUMM_API")
    assert len(names) >= 40
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/ but not exported by libb200mm.so"


def test_product_never_links_the_oracle(built_lib):
    import wgpu_mm_b200 as w
    out = subprocess.run(["nm", "-D", w.lib_path()], capture_output=True, text=True).stdout
    assert "oracle_" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "wgpu_mm_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "liboracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_sass_contains_blackwell_instructions(built_lib):
    """The tensor-core kernel must really be tcgen05 + TMA (B200_PROFILING.md: UTC*MMA, UTMALDG, LDTM)."""
    import wgpu_mm_b200 as w
    sass = subprocess.run(["cuobjdump", "-sass", w.lib_path()], capture_output=True, text=True).stdout
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "UTMALDG" in sass, "no TMA tensor load in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"
    assert "UTMASTG" in sass, "no TMA tensor store in SASS (pair kernel epilogue)"
    assert re.search(r"UTC\w*MMA\.2CTA", sass), "no cta_group::2 MMA in SASS"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", w.lib_path()], capture_output=True, text=True).stdout


def test_sass_of_the_in_kernel_split(built_lib):
    """Tc3xCfg::SPLIT = 2 (lo tiles computed in shared memory): the pair kernel issues TWO tensor loads per k-step (A, B) where
    the pre-pass form issues four (A, A_lo, B, B_lo), and carries the LDS.128 / STS.128 of the split plus the 2-CTA MMA and the
    TMA store."""
    import wgpu_mm_b200 as w
    sass = subprocess.run(["cuobjdump", "-sass", w.lib_path()], capture_output=True, text=True).stdout
    fns = {}
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if "sgemm_tc3x_kernel" in name:
            fns[name] = chunk
    pre = next(v for k, v in fns.items() if "Li256ELi5ELb0ELi16ELi256ELb1ELb1ELi0E" in k)
    split2 = next(v for k, v in fns.items() if "Li256ELi5ELb0ELi16ELi256ELb1ELb1ELi2E" in k)
    assert pre.count("UTMALDG") == 4 and split2.count("UTMALDG") == 2
    assert "LDS.128" in split2 and "STS.128" in split2 and "LDS.128" not in pre
    assert re.search(r"UTC\w*MMA\.2CTA", split2) and "UTMASTG" in split2


def test_no_gpu_fails_loudly(built_lib):
    import wgpu_mm_b200 as w
    if w.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(w.B200mmError) as e:
        w.Context(0)
    assert e.value.code == -3 and "No GPU found" in str(e.value)
    with pytest.raises(w.B200mmError) as e:
        w.harness.test_harness(None, "gemm_5", (64, 64, 64), False)
    assert "No GPU found" in str(e.value)


def test_workload_matches_reference_geometry(built_lib):
    """src/gemm.rs:16-150, src/gemv.rs:17-33 at the crate's shapes."""
    from wgpu_mm_b200.workload import entry_workload
    from wgpu_mm_b200 import KernelId
    expect = {
        "gemm_1": ((64, 64, 1), (16, 16, 1), KernelId.GEMM_1),
        "gemm_1v": ((64, 64, 1), (16, 4, 1), KernelId.GEMM_1V),
        "gemm_2": ((64, 64, 1), (256, 1, 1), KernelId.GEMM_2),
        "gemm_3": ((64, 64, 1), (256, 1, 1), KernelId.GEMM_3),
        "gemm_4": ((64, 64, 1), (128, 1, 1), KernelId.GEMM_4),
        "gemm_5": ((32, 32, 1), (64, 1, 1), KernelId.GEMM_5),
        "qgemv_1": ((32, 1, 1), (8, 1, 1), KernelId.QGEMV_1),
        # orphan shaders, geometry inferred in SURVEY 2.2
        "gemm_wonnx": ((256, 1, 1), (256, 1, 1), KernelId.GEMM_WONNX),
        "bram8x8": ((64, 32, 1), (4, 8, 1), KernelId.BRAM8X8),
        "bram": ((32, 32, 1), (8, 8, 1), KernelId.BRAM),
        "gemm3": ((8, 16, 1), (16, 16, 1), KernelId.GEMM3),
    }
    for name, (grid, block, kid) in expect.items():
        wl, k = entry_workload(name)
        assert (wl.count.x, wl.count.y, wl.count.z) == grid, name
        assert (wl.size.x, wl.size.y, wl.size.z) == block, name
        assert k == kid
    # x <-> N, y <-> M swap of gemm_4/gemm_5 (src/gemm.rs:109,139)
    wl, _ = entry_workload("gemm_5", 64, 256, 32)
    assert (wl.count.x, wl.count.y) == (8, 2)


def test_entry_points_fill_the_context(built_lib):
    from wgpu_mm_b200 import gemm, gemv
    ctx = {}
    assert gemm.insert_matrix_dims(ctx) == (1024, 1024, 1024)
    wl, shader = gemm.gemm_5(ctx)
    assert shader == "gemm_5" and ctx["workgroup_size_x"] == 64
    ctx = {}
    assert gemv.insert_matrix_dims(ctx) == (1, 1024, 1024)
    assert gemv.ABSMAX == 2.0
    wl, shader = gemv.qgemv_1(ctx)
    assert (wl.count.x, wl.size.x) == (32, 8)


def test_compute_dim_and_ceil(built_lib, oracle):
    from wgpu_mm_b200.workload import Workload
    from wgpu_mm_b200 import B200mmError
    assert Workload.ceil(1024, 16) == 64 and Workload.ceil(1025, 16) == 65 and Workload.ceil(1, 32) == 1
    for items in (1, 2, 65535, 65536, 1 << 20, 65535 * 256):
        assert Workload.compute_dim(items, "X") == oracle.compute_dim(items, "X")
    with pytest.raises(B200mmError) as e:
        Workload.compute_dim(65535 * 64 + 1, "Z")
    assert "Compute limits exceeded" in str(e.value)


def test_host_codec_golden_and_vs_oracle(built_lib, oracle):
    """Product codec (host/quant.cc) vs src/quant.rs:48-64 and vs the oracle, bit for bit."""
    from wgpu_mm_b200 import quant, B200mmError
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "test_qdq.json")))
    words, absmax = quant.sint8_quantize(np.array(g["matrix"], dtype=np.float32), 4, 4)
    assert [int(w) for w in words] == g["words"]
    deq = quant.sint8_dequantize(words, absmax, 4, 4).reshape(-1)
    assert np.all(np.abs(np.array(g["matrix"], dtype=np.float32) - deq) < 0.01)
    for (K, N, seed) in ((4, 4, 1), (64, 256, 2), (128, 1024, 3), (3, 4, 4)):
        W = oracle.generate_weight_data(seed, K, N)
        w1, a1 = quant.sint8_quantize(W, K, N)
        w2, a2 = oracle.sint8_quantize(W, K, N)
        assert a1 == a2 and np.array_equal(w1, w2)
        assert np.array_equal(quant.sint8_dequantize(w1, 2.0, K, N), oracle.sint8_dequantize(w2, 2.0, K, N))
    with pytest.raises(B200mmError):
        quant.sint8_quantize(np.zeros(6, dtype=np.float32), 2, 3)  # len % 4 != 0 (src/quant.rs:13)
    with pytest.raises(B200mmError):
        quant.sint8_quantize(np.zeros(8, dtype=np.float32), 4, 4)  # len != K*N (src/quant.rs:12)


def test_cpp_runner_reports_like_cargo(built_lib):
    exe = os.path.join(ROOT, "wgpu_mm_b200", "lib", "wgpu_mm_tests")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "test_qdq"], capture_output=True, text=True)
    assert r.returncode == 0 and "test test_qdq ... ok" in r.stdout
    r = subprocess.run([exe, "test_qdq_grouped"], capture_output=True, text=True)  # hand-computed words / scales of the grouped codec
    assert r.returncode == 0 and "test test_qdq_grouped ... ok" in r.stdout


@pytest.mark.parametrize("kng", [(256, 64, 128), (300, 16, 128), (64, 8, 128), (512, 32, 512)])
def test_grouped_codec_matches_oracle(oracle, kng):
    """Per-group scales (extension of src/quant.rs:17): product codec (host/quant.cc) == oracle restatement, bit for bit;
    the round trip is within half a quantisation step of each group's own scale."""
    from wgpu_mm_b200.quant import sint8_dequantize_grouped, sint8_quantize_grouped, split_grouped
    K, N, G = kng
    W = oracle.generate_weight_data(90, K, N)
    W[: K // 2] *= 9.0
    W[3, :] = 0.0
    packed = sint8_quantize_grouped(W, K, N, G)
    words, scales = split_grouped(packed, K, N, G)
    owords, oscales = oracle.sint8_quantize_grouped(W, K, N, G)
    assert np.array_equal(words, owords) and np.array_equal(scales, oscales)
    assert scales.shape == (-(-K // G), N)
    back = sint8_dequantize_grouped(packed, K, N, G)
    assert np.array_equal(back, oracle.sint8_dequantize_grouped(owords, oscales, K, N, G))
    step = np.repeat(scales, G, axis=0)[:K] / 127.0
    assert (np.abs(back - W) <= 0.5 * step + 1e-7).all()


def test_grouped_codec_one_group_equals_per_column_global(oracle):
    """group_k >= K: one scale per column; a column whose absmax equals the matrix absmax packs exactly like src/quant.rs."""
    from wgpu_mm_b200.quant import sint8_quantize, sint8_quantize_grouped, split_grouped
    K, N = 64, 16
    W = oracle.generate_weight_data(91, K, N)
    W[0, :] = 0.25  # every column now has the same absmax = the global one
    words, scales = split_grouped(sint8_quantize_grouped(W, K, N, 128), K, N, 128)
    gwords, absmax = sint8_quantize(W, K, N)
    assert absmax == 0.25 and (scales == 0.25).all() and np.array_equal(words, gwords)


def test_grouped_codec_zero_group_and_asserts():
    from wgpu_mm_b200 import B200mmError
    from wgpu_mm_b200.quant import sint8_dequantize_grouped, sint8_quantize_grouped, split_grouped
    W = np.zeros((128, 8), dtype=np.float32)
    packed = sint8_quantize_grouped(W, 128, 8, 128)  # 0/0 = NaN -> `as i32` = 0, like the reference on an all-zero matrix
    words, scales = split_grouped(packed, 128, 8, 128)
    assert not words.any() and not scales.any()
    assert not sint8_dequantize_grouped(packed, 128, 8, 128).any()
    with pytest.raises(B200mmError):
        sint8_quantize_grouped(W, 128, 6, 128)
    with pytest.raises(B200mmError):
        sint8_quantize_grouped(W, 128, 8, 0)
    with pytest.raises(B200mmError):
        sint8_dequantize_grouped(packed[:-1], 128, 8, 128)


TC3X_SHAPES = [(4096, 4096, 4096), (16384, 2048, 16384), (128, 4096, 4096), (256, 4096, 4096), (512, 4096, 4096), (1024, 4096, 4096),
               (2048, 4096, 4096), (1000, 520, 300), (128, 256, 256), (128, 128, 16), (4096, 4096, 100), (300, 36, 4100), (19000, 256, 512),
               (1792, 1792, 1792), (2304, 2304, 2304), (2560, 2560, 2560), (2048, 4096, 2048)]  # L2-resident: stream-K instead of idle SMs / tiny tails


@pytest.mark.parametrize("shape", TC3X_SHAPES)
@pytest.mark.parametrize("cfg", [(256, 16), (256, 32), (128, 32), (512, 16)])  # bn = 512: the 2-CTA kernel (256 x 256 tiles on SM pairs)
@pytest.mark.parametrize("pure", [0, 1])
def test_tc3x_schedule_covers_every_unit_once(shape, cfg, pure):
    """The SGEMM_TC3X work schedule (hybrid waves + stream-K, or the uniform k-split when tiles < SMs) as the kernel's own
    segment iterator walks it, run on the host for all CTAs: every (tile, chain) unit exactly once, grid <= SMs."""
    from wgpu_mm_b200 import lib
    M, N, K = shape
    bn, bk = cfg
    sms = 148
    out = (C.c_int * 6)()
    assert lib().b200mm_tc3x_schedule(M, N, K, bn, bk, sms, pure, out) == 0
    grid, full_waves, cpt, k_split, tiles, sk_units = list(out)
    if bn == 512:  # grid counts CTA pairs
        assert tiles == -(-M // 256) * -(-N // 256) and cpt == -(-(-(-K // bk)) // (256 // bk))
        units = sms // 2
    else:
        assert tiles == -(-M // 128) * -(-N // bn) and cpt == -(-(-(-K // bk)) // (256 // bk))
        units = sms
    assert 1 <= grid <= units
    cover = np.zeros(tiles * cpt, dtype=np.uint16)
    mseg, mch = C.c_int(), C.c_int()
    assert lib().b200mm_tc3x_schedule_cover(M, N, K, bn, bk, sms, pure, cover.ctypes.data_as(C.c_void_p), cover.size, C.byref(mseg), C.byref(mch)) == 0
    assert (cover == 1).all(), f"units visited {np.bincount(cover)} times"
    assert full_waves * grid * cpt + sk_units == tiles * cpt
    # balance: whole-tile waves plus an equal share (rounded up) of the stream-K units
    assert mch.value <= full_waves * cpt + -(-sk_units // grid)
    if k_split and not pure:
        # fewer tiles than SMs: one segment per CTA, every tile cut into the same k_split slices
        assert tiles < units and cpt % k_split == 0 and grid == tiles * k_split and tiles * k_split <= units
        assert mseg.value == 1 and mch.value == cpt // k_split
        assert all(cpt % d or tiles * d > units for d in range(k_split + 1, cpt + 1))  # k_split is the largest admissible divisor


# shapes where stream-K units < CTAs, so some CTAs have an empty range (the advisor's hang list) + the regular ones
TC3X_REPLAY_SHAPES = TC3X_SHAPES + [(4096, 4096, 512), (16384, 16384, 512), (4096, 14336, 512), (4096, 4096, 256), (4224, 4096, 768),
                                    (19072, 512, 1024), (16384, 16384, 16384)]


@pytest.mark.parametrize("shape", TC3X_REPLAY_SHAPES)
@pytest.mark.parametrize("cfg", [(256, 16), (256, 32), (128, 32), (512, 16)])
@pytest.mark.parametrize("pure", [0, 1])
def test_tc3x_stream_k_fixup_protocol_replay(shape, cfg, pure):
    """Replays the owner/contributor dependency graph of the stream-K tail with the kernel's own contributor rule: a finisher
    waits only on lower-numbered CTAs that really publish a part of its tile (never on a CTA with an empty range -- the round-1
    hang at 4096 x 4096 x 512), and the published parts plus its own chains cover the tile exactly."""
    from wgpu_mm_b200 import lib
    M, N, K = shape
    bn, bk = cfg
    for sms in (148, 132, 7):
        bad, mw = C.c_int(-1), C.c_int()
        assert lib().b200mm_tc3x_schedule_replay(M, N, K, bn, bk, sms, pure, C.byref(bad), C.byref(mw)) == 0
        assert bad.value == 0, f"{bad.value} protocol violations at {shape} {cfg} sms={sms}"
        assert mw.value <= max(1, -(-(-(-K // bk)) // (256 // bk)))  # at most chains_per_tile - 1 parts per tile (+ slack for 1)


def test_tc3x_schedule_reference_points():
    """The cases quoted in DESIGN.md 4.1."""
    from wgpu_mm_b200 import lib
    out = (C.c_int * 6)()
    lib().b200mm_tc3x_schedule(4096, 4096, 4096, 256, 16, 148, 0, out)
    assert list(out) == [148, 3, 16, 0, 512, 68 * 16]  # 3 whole waves + 68 tiles split 148 ways
    lib().b200mm_tc3x_schedule(256, 4096, 4096, 256, 16, 148, 0, out)
    assert list(out)[:5] == [128, 0, 16, 4, 32]  # the 256-row panels of the host-buffer path: 32 tiles x 4 k-slices
    lib().b200mm_tc3x_schedule(1024, 4096, 4096, 256, 16, 148, 0, out)
    assert list(out)[:5] == [128, 1, 16, 1, 128]  # one tile per CTA, in lock-step
    lib().b200mm_tc3x_schedule(1792, 1792, 1792, 512, 16, 148, 0, out)
    assert list(out) == [74, 0, 7, 0, 49, 49 * 7]  # 49 pair tiles of 7 chains: no k-split divides, operands fit in L2 -> stream-K over all pairs
    lib().b200mm_tc3x_schedule(2304, 2304, 2304, 512, 16, 148, 0, out)
    assert list(out) == [74, 0, 9, 0, 81, 81 * 9]  # 81 pair tiles = one wave + 7, L2-resident: no whole wave, stream-K over everything
    lib().b200mm_tc3x_schedule(2048, 2048, 2048, 512, 16, 148, 0, out)
    assert list(out)[:5] == [64, 1, 8, 1, 64]  # 64 pair tiles for 74 pairs = 86 % of the machine: one whole tile per pair, no stream-K
    assert lib().b200mm_tc3x_schedule(0, 4096, 4096, 256, 16, 148, 0, out) != 0
    assert lib().b200mm_tc3x_schedule(128, 128, 128, 192, 16, 148, 0, out) != 0


def test_tc3x_default_rules(built_lib):
    """The shape rules of setup_tc3x, device-free (b200mm_tc3x_plan): the cases DESIGN.md 4.1 quotes, on a 148-SM device.
    plan = (tile columns, BK, pairs, TMA store, lo operands computed in shared memory, all of A in the pre-pass, grid, whole waves, k-slices)."""
    from wgpu_mm_b200 import lib

    def plan(M, N, K, tune=None, flags=0):
        out = (C.c_int * 9)()
        t = (C.c_uint32 * 4)(*tune) if tune else None
        assert lib().b200mm_tc3x_plan(M, N, K, 148, t, flags, out) == 0
        return tuple(out)

    assert plan(4096, 4096, 4096) == (256, 16, 1, 1, 0, 0, 148, 3, 0)            # pairs, 3 whole waves + stream-K tail, split pre-pass
    assert plan(16384, 2048, 16384)[:5] == (256, 16, 1, 1, 0)                     # the 8-GPU panel
    assert plan(1024, 1024, 1024) == (128, 32, 0, 0, 0, 0, 128, 0, 2)             # the reference's test shape: 128 x 128 tiles, 2 k-slices
    assert plan(512, 512, 512)[:3] == (128, 32, 0)
    assert plan(128, 4096, 4096) == (128, 32, 0, 0, 1, 1, 128, 0, 4)              # skinniest: 128 x 128 tiles + B_lo in shared memory
    assert plan(128, 14336, 4096)[:6] == (256, 16, 0, 0, 1, 1)                    # skinny, many columns: 256-column tiles + B_lo in shared memory
    assert plan(256, 4096, 4096)[:6] == (256, 16, 0, 0, 1, 1)                     # (the pipelined host path pins its panels to tune[3] = 2)
    assert plan(512, 4096, 4096)[:6] == (256, 16, 0, 0, 0, 0)                     # M > 256: pre-pass
    assert plan(2048, 2048, 2048)[:3] == (256, 16, 1) and plan(2048, 2048, 2048)[6:] == (128, 1, 1)   # L2-resident: pairs from 48 pair tiles
    assert plan(2304, 2304, 2304)[6:] == (148, 0, 0)                              # 81 pair tiles, L2-resident: stream-K over everything
    assert plan(3072, 3072, 3072)[6:] == (148, 1, 0)                              # beyond L2: whole waves + tail
    assert plan(768, 4096, 4096)[2] == 1 and plan(768, 4096, 4096)[6:] == (148, 0, 0)   # 48 pair tiles would idle a third of the SMs
    assert plan(640, 4096, 4096)[2] == 0                                          # 256-row tiles would pad 640 rows to 768
    assert plan(1024, 4096, 4096)[2] == 1
    assert plan(1000, 1001, 515)[:2] == plan(1000, 1004, 516)[:2]                 # ragged N / K: the padded shape's plan
    # overrides
    assert plan(4096, 4096, 4096, (513, 0, 0, 0))[2] == 0 and plan(1024, 1024, 1024, (512, 0, 0, 0))[2] == 1
    assert plan(4096, 4096, 4096, (0, 0, 0, 5))[4:6] == (2, 0) and plan(4096, 4096, 4096, (0, 0, 0, 3))[4:6] == (0, 1)
    assert plan(4096, 4096, 4096, (0, 0, 6, 0))[3] == 0 and plan(4096, 4096, 4096, (0, 0, 32, 0))[1] == 32
    assert plan(4096, 4096, 4096, None, 1)[:5] == (256, 32, 0, 0, 0)              # B200MM_F_TC3X_1X: single-pass kernel
    out = (C.c_int * 9)()
    assert lib().b200mm_tc3x_plan(0, 1, 1, 148, None, 0, out) != 0


@pytest.mark.parametrize("rows_per_split", [128, 256, 1024, 2048, 4096])
@pytest.mark.parametrize("group_k", [32, 64, 128, 256, 512])
def test_gemv_blocked_row_order_model(rows_per_split, group_k):
    """Arithmetic model of gemv.cuh's BLOCKED row order (the index math restated, WARPS = 8, LPR = 16 -> RPW = 2, UNROLL = 4): the
    loop counts in the interleaved coordinate kk and prow() maps it to the physical row.  Every row of the slab is visited
    exactly once, a thread's rows ascend, one loop iteration (2 * UNROLL loads) never straddles a quantisation group, and at
    most one group boundary is crossed between consecutive iterations -- the three facts the per-group fold relies on."""
    WARPS, RPW, UNROLL = 8, 2, 4
    RSTEP = WARPS * RPW
    k_beg = 3 * rows_per_split  # some split in the middle of K
    k_end = k_beg + rows_per_split
    seen = np.zeros(rows_per_split, dtype=np.int32)
    for warp in range(WARPS):
        for riw in range(RPW):
            k0 = k_beg + warp * RPW + riw
            p0 = k_beg + warp * (rows_per_split // WARPS) + riw
            prow = lambda kk: p0 + (kk - k0) // WARPS
            last_group, last_row = None, -1
            k = k0
            while k < k_end:
                rows = [prow(k) + u * RPW for u in range(2 * UNROLL) if k + u * RSTEP < k_end]
                groups = {r // group_k for r in rows}
                assert len(groups) == 1, "an iteration straddles a group"
                g = groups.pop()
                assert g == prow(k) // group_k  # group_step() looks at the iteration's first row only
                assert last_group is None or g - last_group in (0, 1)
                assert rows[0] > last_row and rows == sorted(rows)
                last_group, last_row = g, rows[-1]
                for r in rows:
                    assert k_beg <= r < k_end
                    seen[r - k_beg] += 1
                k += 2 * UNROLL * RSTEP
    assert (seen == 1).all()


# ---- rust/ : the reference-side crate cannot be compiled here (no cargo), so it is checked mechanically ----
RUST = os.path.join(ROOT, "rust", "src")


def _rust(name):
    return open(os.path.join(RUST, name)).read()


def test_rust_shim_binds_every_exported_symbol(built_lib):
    """rust/src/ffi.rs declares every function the two headers export (and nothing the library lacks), and its kernel-id,
    status and flag constants equal the header's."""
    ffi = _rust("ffi.rs")
    declared = _declared("b200mm.h", "B200MM_API") + _declared("wgpu_mm_c.h", "WGPUMM_API")
    bound = re.findall(r"pub fn ((?:b200mm|wgpumm)_\w+)\(", ffi)
    assert sorted(set(declared) - set(bound)) == [], "exported by the headers but not bound in rust/src/ffi.rs"
    assert sorted(set(bound) - set(declared)) == [], "bound in rust/src/ffi.rs but not declared in include/"
    for n in bound:
        assert hasattr(built_lib, n)
    hdr = open(os.path.join(ROOT, "include", "b200mm.h")).read()
    consts = dict(re.findall(r"\b(B200MM_(?:K|ERR)_\w+|B200MM_OK)\s*=\s*(-?\d+)", hdr))
    consts.update({k: str(int(v, 16)) for k, v in re.findall(r"#define\s+(B200MM_F_\w+)\s+(0x[0-9a-fA-F]+|0)u", hdr)})
    rust_consts = {k: str(int(v, 0)) for k, v in re.findall(r"pub const (B200MM_\w+): \w+ = (-?(?:0x[0-9a-fA-F]+|\d+));", ffi)}
    assert len(consts) >= 25
    assert rust_consts == consts
    # struct layouts: same field order and count as the C definitions
    c_fields = re.findall(r"(\w+)(?:\[\d\])?;", re.search(r"typedef struct b200mm_kernel_params \{(.*?)\} b200mm_kernel_params;", hdr, re.S).group(1).replace("/*", "\n/*"))
    r_fields = re.findall(r"pub (\w+):", re.search(r"pub struct b200mm_kernel_params \{(.*?)\n\}", ffi, re.S).group(1))
    assert r_fields == ["workgroup_size", "absmax", "batch", "flags", "tune", "group_k"] and all(f in c_fields for f in r_fields)


def test_rust_crate_has_every_upstream_public_item_with_a_body():
    """src/lib.rs:1-10, src/gemm.rs:9-150, src/gemv.rs:8-33, src/quant.rs:7-43, src/harness.rs:170, src/workload.rs of the reference:
    every public item exists in rust/src with a body (no commented-out stubs), plus the upstream test names."""
    lib = _rust("lib.rs")
    for mod in ("pub mod gemm;", "pub mod gemv;", "pub mod quant;", "mod harness;", "pub use harness::*;", "pub use launch_shape::*;"):
        assert mod in lib
    gemm, gemv, quant, harness, shape = (_rust(f) for f in ("gemm.rs", "gemv.rs", "quant.rs", "harness.rs", "launch_shape.rs"))
    m = re.search(r"entry_point!\(([^)]*)\);", gemm)
    assert m and [s.strip() for s in m.group(1).split(",")] == ["gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5", "gemm_wonnx", "bram",
                                                                "bram8x8", "gemm3", "sgemm_simt", "sgemm_tc3x"]
    assert "pub fn insert_matrix_dims(context: &mut Context) -> (usize, usize, usize) {" in gemm
    assert "pub fn insert_matrix_dims(context: &mut Context) -> (usize, usize, usize) {" in gemv
    assert "pub const ABSMAX: f32 = 2.0;" in gemv
    for fn in ("qgemv_1", "qgemv_sint8", "gemv_f32"):
        assert re.search(rf"pub fn {fn}\(tera: &mut Tera, context: &mut Context\) -> \(Workload, String\) \{{", gemv)
    assert re.search(r"pub fn sint8_quantize<F: QuantFloat>\(matrix: &\[F\], K: usize, N: usize\) -> \(Vec<u32>, F\) \{", quant)
    assert "pub fn sint8_dequantize(quantized_matrix: &[u32], absmax: f32, K: usize, N: usize) -> Vec<f32> {" in quant
    assert "pub async fn test_harness(workload: Workload, shader: String, dims: (usize, usize, usize), quantize_b: bool) {" in harness
    assert 'panic!("MAE too high")' in harness and "fn mm_ref(" in harness
    for item in ("pub struct WorkgroupCount(pub u32, pub u32, pub u32);", "pub struct WorkgroupSize(pub u32, pub u32, pub u32);", "pub struct Workload {",
                 "pub enum WorkloadDim {", "pub fn compute_dim(work_items: usize, dim: WorkloadDim) -> (u32, u32) {", "pub fn ceil(num: usize, div: usize) -> usize {",
                 "MAX_COMPUTE_WORKGROUPS_PER_DIMENSION: usize = 65535"):
        assert item in shape, item
    tests = set(re.findall(r"gemm_test!\((test_\w+),", gemm)) | set(re.findall(r"pub (?:async )?fn (test_\w+)\(", gemm + gemv + quant))
    assert {"test_gemm_1", "test_gemm_1v", "test_gemm_2", "test_gemm_3", "test_gemm_4", "test_gemm_5", "test_qgemv_1", "test_qdq"} <= tests
    for f in os.listdir(RUST):  # balanced braces: the cheapest syntax check available without rustc
        src = re.sub(r'"(?:[^"\\]|\\.)*"', '""', re.sub(r"//.*", "", _rust(f)))
        assert src.count("{") == src.count("}") and src.count("(") == src.count(")") and src.count("[") == src.count("]"), f


def test_rust_entry_table_matches_the_compiled_mirror(built_lib):
    """The geometry table in rust/src/gemm.rs / gemv.rs, evaluated here, equals what the C++ mirror (host/entry_points.cc)
    produces through wgpumm_entry_workload -- at the reference shapes and at a BASELINE shape."""
    from wgpu_mm_b200.workload import entry_workload
    hdr = open(os.path.join(ROOT, "include", "b200mm.h")).read()
    ids = {k: int(v) for k, v in re.findall(r"\b(B200MM_K_\w+)\s*=\s*(\d+)", hdr)}
    table = []
    pat = re.compile(r'Entry \{ name: "(\w+)",\s*kernel_id: (\w+),.*?size: \((\d+), (\d+), (\d+)\),\s*grid_x: \(Dim::(\w+), ([\d *]+)\),\s*grid_y: \(Dim::(\w+), ([\d *]+)\) \}', re.S)
    for name, kid, sx, sy, sz, dx, vx, dy, vy in pat.findall(_rust("gemm.rs")):
        table.append((name, ids[kid], (int(sx), int(sy), int(sz)), (dx, eval(vx)), (dy, eval(vy))))
    for name, kid, sx, sy, sz, per in re.findall(r'Entry::new\("(\w+)", (\w+), \((\d+), (\d+), (\d+)\), ([\d *]+)\)', _rust("gemv.rs")):
        table.append((name, ids[kid], (int(sx), int(sy), int(sz)), ("N", eval(per)), ("One", 1)))
    assert len(table) == 15
    ceil = lambda a, b: -(-a // b)
    for name, kid, size, gx, gy in table:
        gemv = kid in (ids["B200MM_K_QGEMV_1"], ids["B200MM_K_GEMV_F32"], ids["B200MM_K_QGEMV_SINT8"])
        for (M, N, K) in ([(1, 1024, 1024), (1, 14336, 4096)] if gemv else [(1024, 1024, 1024), (4096, 4096, 4096), (256, 512, 128)]):
            ext = {"M": M, "N": N, "MN": M * N, "One": 1}
            want = (ceil(ext[gx[0]], gx[1]), ceil(ext[gy[0]], gy[1]), 1)
            if name == "sgemm_tc3x":
                want = (min(148, ceil(M, 128) * ceil(N, 256)), 1, 1)
            wl, got_id = entry_workload(name, M, N, K)
            assert got_id == kid, name
            assert (wl.count.x, wl.count.y, wl.count.z) == want, (name, M, N, K)
            assert (wl.size.x, wl.size.y, wl.size.z) == size, name
