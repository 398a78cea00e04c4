"""Writes tests/golden/*.json.

The reference holds exactly one golden vector for this path -- test_qdq, src/quant.rs:48-64 -- and it
is transcribed here from the reference source (input, expected words, tolerance).  The reference's
toolchain (nightly Rust + wgpu + a Vulkan ICD) is absent from the image, so no further vectors can be
produced by running it; the remaining fixtures pin OUR seeded data stream (the reference's RNG is
unseeded, src/harness.rs:111) so that the CPU oracle, the C++ host harness and the CUDA generator
cannot drift apart silently.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    qdq = {
        "source": "src/quant.rs:48-64 (test_qdq), transcribed from the reference",
        "K": 4, "N": 4,
        "matrix": [0.1, -0.1, 0.5, -0.5, 1.0, -1.0, 1.2, -1.2] * 2,
        "words": [3409310987, 2172622442, 3409310987, 2172622442],
        "absmax": 1.2,
        "roundtrip_tolerance": 0.01,
    }
    json.dump(qdq, open(os.path.join(HERE, "test_qdq.json"), "w"), indent=1)

    stream = {
        "source": "oracle_generate_weight_data_at (seeded restatement of src/harness.rs:103-121)",
        "cases": [],
    }
    for seed, offset in ((1, 0), (0x5EED, 0), (7, 1 << 33)):
        v = oracle.generate_weight_data(seed, 1, 16, offset=offset).reshape(-1)
        stream["cases"].append({"seed": seed, "offset": offset, "bits": [int(x) for x in v.view(np.uint32)]})
    json.dump(stream, open(os.path.join(HERE, "weight_stream.json"), "w"), indent=1)

    # small seeded GEMM / qGEMV cases with their mm_ref outputs (bit patterns), for regression of the oracle itself
    cases = []
    for (M, N, K, seed) in ((8, 12, 16, 3), (5, 8, 20, 4), (1, 32, 64, 5)):
        A = oracle.generate_weight_data(seed, M, K)
        B = oracle.generate_weight_data(seed + 100, K, N)
        Cm = oracle.mm_ref_literal(A, B)
        cases.append({"M": M, "N": N, "K": K, "seed_a": seed, "seed_b": seed + 100,
                      "c_bits": [int(x) for x in Cm.reshape(-1).view(np.uint32)]})
    json.dump({"source": "oracle_mm_ref_literal (src/harness.rs:17-28)", "cases": cases},
              open(os.path.join(HERE, "mm_ref_small.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
