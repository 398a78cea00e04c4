"""CPU tests of the oracle (test infrastructure) against the reference's golden vector and itself."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_qdq_golden(oracle):
    """src/quant.rs:48-64 -- the reference's only known-answer test, transcribed in tests/golden/test_qdq.json."""
    g = json.load(open(os.path.join(GOLD, "test_qdq.json")))
    m = np.array(g["matrix"], dtype=np.float32)
    words, absmax = oracle.sint8_quantize(m, g["K"], g["N"])
    assert len(words) == 4
    assert [int(w) for w in words] == g["words"]
    assert absmax == pytest.approx(g["absmax"], rel=1e-6)
    deq = oracle.sint8_dequantize(words, absmax, g["K"], g["N"]).reshape(-1)
    assert np.all(np.abs(m - deq) < g["roundtrip_tolerance"])


def test_qdq_grouped_hand_computed(oracle):
    """Per-group scales are an extension (the reference has one global absmax, src/quant.rs:17), so nothing in the reference pins
    them; the oracle is pinned by hand instead, on the test_qdq matrix with 2-row groups: scales = column maxima per row pair,
    0.1/1.0*127 = 12.7 -> 13, 0.5/1.2*127 = 52.9 -> 53, and every entry of rows 1 and 3 is its own group's absmax -> +-127.
    group_k = K must reproduce the per-column global codec, and with one scale for everything the golden words of test_qdq."""
    g = json.load(open(os.path.join(GOLD, "test_qdq.json")))
    m = np.array(g["matrix"], dtype=np.float32)
    words, scales = oracle.sint8_quantize_grouped(m, 4, 4, 2)
    assert scales.tolist() == [[1.0, 1.0, np.float32(1.2), np.float32(1.2)]] * 2
    b = lambda q: q & 0xFF
    row0 = b(13) | b(-13) << 8 | b(53) << 16 | b(-53) << 24
    assert [int(w) for w in words] == [row0, 0x817F817F, row0, 0x817F817F]
    deq = oracle.sint8_dequantize_grouped(words, scales, 4, 4, 2).reshape(-1)
    assert np.all(np.abs(m - deq) < 0.006)
    # a matrix whose columns all share the global absmax: grouped (one group) == the reference codec == golden words
    m2 = m.reshape(4, 4).copy()
    m2[3] = [1.2, -1.2, 1.2, -1.2]  # now every column contains the global absmax
    words1, scales1 = oracle.sint8_quantize_grouped(m2, 4, 4, 4)
    gw, absmax = oracle.sint8_quantize(m2, 4, 4)
    assert (scales1 == absmax).all() and np.array_equal(words1, gw)
    assert [int(w) for w in gw[:3]] == g["words"][:3]  # rows 0-2 are still the reference's golden words


def test_weight_stream_golden(oracle):
    g = json.load(open(os.path.join(GOLD, "weight_stream.json")))
    for c in g["cases"]:
        v = oracle.generate_weight_data(c["seed"], 1, 16, offset=c["offset"]).reshape(-1)
        assert [int(x) for x in v.view(np.uint32)] == c["bits"]


def test_mm_ref_small_golden(oracle):
    g = json.load(open(os.path.join(GOLD, "mm_ref_small.json")))
    for c in g["cases"]:
        A = oracle.generate_weight_data(c["seed_a"], c["M"], c["K"])
        B = oracle.generate_weight_data(c["seed_b"], c["K"], c["N"])
        for fn in (oracle.mm_ref, oracle.mm_ref_literal):
            got = fn(A, B).reshape(-1).view(np.uint32)
            assert [int(x) for x in got] == c["c_bits"]


def test_weight_distribution(oracle):
    """src/harness.rs:112-116: Uniform[-10,10)/50 => values in [-0.2, 0.2)."""
    v = oracle.generate_weight_data(123, 512, 512)
    assert v.min() >= -0.2 and v.max() < 0.2
    assert abs(float(v.mean())) < 2e-3
    assert abs(float(v.std()) - 0.4 / np.sqrt(12)) < 2e-3
    # counter based: a sub-range equals the same range of the full stream
    w = oracle.generate_weight_data(123, 1, 100, offset=512 * 7 + 5).reshape(-1)
    assert np.array_equal(w, v.reshape(-1)[512 * 7 + 5: 512 * 7 + 105])


@pytest.mark.parametrize("shape", [(1, 64, 128), (33, 40, 52), (128, 128, 128), (7, 1030, 19)])
def test_mm_ref_fast_order_is_bit_identical(oracle, shape):
    M, N, K = shape
    A = oracle.generate_weight_data(1, M, K)
    B = oracle.generate_weight_data(2, K, N)
    assert np.array_equal(oracle.mm_ref(A, B), oracle.mm_ref_literal(A, B))


def test_mm_ref_matches_numpy_f64(oracle):
    A = oracle.generate_weight_data(1, 64, 256)
    B = oracle.generate_weight_data(2, 256, 96)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    assert np.abs(oracle.mm_ref(A, B) - ref).max() < 2e-6
    assert np.abs(oracle.mm_f64(A, B) - ref).max() < 1e-12
    rows = np.array([0, 5, 63])
    assert np.abs(oracle.mm_f64_rows(A, B, rows) - ref[rows]).max() < 1e-12


@pytest.mark.parametrize("name", ["gemm_1", "gemm_1v", "gemm_2", "gemm_3", "gemm_4", "gemm_5", "gemm_wonnx", "bram", "gemm3"])
def test_wgsl_restatements_compute_the_product(oracle, name):
    """Every shader restatement (with its own dispatch geometry) must equal A*B; orders differ by ulps."""
    M, N, K = 64, 96, 128
    A = oracle.generate_weight_data(11, M, K)
    B = oracle.generate_weight_data(12, K, N)
    ref = oracle.mm_f64(A, B)
    got = oracle.wgsl_gemm(name, A, B)
    assert not np.any(got == 123.25), "restatement left outputs unwritten (geometry error)"
    assert np.abs(got - ref).max() < 5e-7
    # the reference's gate (src/harness.rs:82) holds with a wide margin
    assert oracle.max_abs_err(got, oracle.mm_ref(A, B)) <= 1e-3


def test_sequential_mul_add_shaders_equal_mm_ref_bitwise(oracle):
    """gemm_1/1v/2/bram/gemm3 accumulate k-sequentially with mul,add -- same order as mm_ref."""
    A = oracle.generate_weight_data(21, 32, 64)
    B = oracle.generate_weight_data(22, 64, 32)
    ref = oracle.mm_ref(A, B)
    for name in ("gemm_1", "gemm_1v", "gemm_2", "bram", "gemm3"):
        assert np.array_equal(oracle.wgsl_gemm(name, A, B), ref), name
    # the fma shaders agree with each other bitwise
    g3 = oracle.wgsl_gemm("gemm_3", A, B)
    assert np.array_equal(oracle.wgsl_gemm("gemm_4", A, B), g3)
    assert np.array_equal(oracle.wgsl_gemm("gemm_5", A, B), g3)


def test_quant_properties(oracle):
    K, N = 64, 128
    W = oracle.generate_weight_data(5, K, N)
    words, absmax = oracle.sint8_quantize(W, K, N)
    assert absmax == pytest.approx(float(np.abs(W).max()))
    q = words.view(np.int8)
    assert q.min() >= -127 and q.max() <= 127  # 0x80 is never emitted
    deq = oracle.sint8_dequantize(words, absmax, K, N)
    assert np.abs(deq - W).max() <= absmax / 127 / 2 * 1.0001
    # idempotence: quantising the dequantised matrix reproduces the words
    words2, _ = oracle.sint8_quantize(deq, K, N)
    assert np.array_equal(words, words2)
    # little-endian packing, element 0 in the low byte (src/quant.rs:21-25)
    expect0 = int(np.round(W[0, 0] / np.float32(absmax) * np.float32(127))) & 0xFF
    assert int(words[0]) & 0xFF == expect0


def test_quant_round_half_away_from_zero(oracle):
    # absmax = 127 so x/absmax*127 == x exactly: .5 cases must round away from zero like f32::round
    m = np.array([127.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.49], dtype=np.float32)
    words, absmax = oracle.sint8_quantize(m, 2, 4)
    q = words.view(np.int8)
    assert absmax == 127.0
    assert list(q) == [127, 1, -1, 2, -2, 3, -3, 0]


@pytest.mark.parametrize("batch", [1, 3])
def test_qgemv_restatement_vs_reference_semantics(oracle, batch):
    """qgemv_1.wgsl restatement vs mm_ref(A, dequant(Bq, ABSMAX)) (src/harness.rs:42-48,58)."""
    N, K, ABSMAX = 256, 128, 2.0
    x = oracle.generate_weight_data(31, batch, K)
    outs = []
    Ws = []
    for b in range(batch):
        W = oracle.generate_weight_data(40 + b, K, N)
        Ws.append(oracle.sint8_quantize(W, K, N)[0])
    Bq = np.concatenate(Ws)
    got = oracle.wgsl_qgemv_1(x, Bq, N, K, ABSMAX, batch=batch)
    for b in range(batch):
        ref = oracle.qgemv_ref(x[b:b + 1], Ws[b], 1, N, K, ABSMAX)
        assert oracle.max_abs_err(got[b:b + 1], ref) < 1e-4
        f64 = oracle.qgemv_f64(x[b:b + 1], Ws[b], 1, N, K, ABSMAX)
        e, m = oracle.err_vs_f64(got[b:b + 1], f64)
        assert e / m < 5e-6


def test_compute_dim(oracle):
    """src/workload.rs:48-68."""
    assert oracle.compute_dim(1, "X") == (1, 1)
    assert oracle.compute_dim(65535, "X") == (65535, 1)
    assert oracle.compute_dim(65536, "X") == (32768, 2)
    assert oracle.compute_dim(1 << 20, "X") == (61681, 17)
    assert oracle.compute_dim(65535 * 64, "Z") == (65535, 64)
    with pytest.raises(RuntimeError):
        oracle.compute_dim(65535 * 64 + 1, "Z")
    with pytest.raises(RuntimeError):
        oracle.compute_dim(65535 * 256 + 1, "X")


def test_tolerance_table_3xtf32(oracle):
    """1xTF32 fails the reference gate, a round-to-nearest 3xTF32 split passes with margin."""
    K = 1024
    A = oracle.generate_weight_data(1, 48, K)
    B = oracle.generate_weight_data(2, K, 48)

    def rna_tf32(x):
        b = x.view(np.uint32).astype(np.uint64)
        b = (b + 0x1000) & 0xFFFFE000
        return b.astype(np.uint32).view(np.float32)

    ref = oracle.mm_f64(A, B)
    ah, bh = rna_tf32(A), rna_tf32(B)
    al, bl = rna_tf32(A - ah), rna_tf32(B - bh)
    one = ah.astype(np.float64) @ bh.astype(np.float64)
    three = one + ah.astype(np.float64) @ bl.astype(np.float64) + al.astype(np.float64) @ bh.astype(np.float64)
    assert np.abs(one - ref).max() > 2e-4
    assert np.abs(three - ref).max() < 1e-6
