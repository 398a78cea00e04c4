#!/usr/bin/env python
"""bench.py -- headline benchmark of the SGEMM / GEMV hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ...]

N = 1   workload = SGEMM 4096 x 4096 x 4096 fp32 (BASELINE configs[1]) through the tcgen05 3xTF32 kernel.
        A "step" is one C = A*B.  Three (A, B, C) sets (576 MiB > 126 MB L2) are rotated between steps.
N > 1   workload = SGEMM 16384^3, B and C N-sharded over the ranks (BASELINE configs[4]); every rank ends
        the step holding the full C ("fused": the epilogue stores each tile to every peer over NVLink;
        "nccl": panel GEMM + NCCL all-gather + interleave kernel).  One process per GPU under torchrun.

One JSON line on stdout (rank 0).  `value` = whole-job TFLOP/s (2*M*N*K flop per step, FP32-equivalent)
with inputs resident in HBM; `e2e` = the same metric through b200mm_mm_host with pinned HOST buffers
(H2D of A and B and D2H of C inside the timed region); `roofline` is for the dominant kernel
(sgemm_tc3x_kernel) timed by its own CUDA-event pair inside each step; `extras` carries the other
BASELINE configs (SIMT SGEMM, fp32 GEMV, sint8 GEMV) measured the same way, each with its own roofline.

--impl reference times the reference's own CPU implementation of the path -- mm_ref, src/harness.rs:17-28,
the code the reference crate itself executes on the host (its WGSL shaders need wgpu + a Vulkan ICD, absent
here) -- restated in oracle/oracle.c, on all host cores, on a bounded row-slice of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_traffic():
    """DRAM bytes per launch measured by ncu (committed under profiles/); None when a kernel has no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0, "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int = 0):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0 - 0.05) and (t1 is None or t <= t1 + 0.15)] or [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path (mm_ref) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_mm_ref_sample(M, N, K, target_s=12.0, reps=1):
    """Times oracle.mm_ref (src/harness.rs:17-28 restated, OpenMP over all cores) on a row-slice of the M x N x K
    workload sized for ~target_s seconds.  Returns (tflops, rows, seconds, cores)."""
    import oracle
    oracle.build()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can get
    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = oracle.num_threads()
    B = oracle.generate_weight_data(2, K, N)
    probe_rows = max(8, min(M, cores * 2))
    A = oracle.generate_weight_data(1, probe_rows, K)
    t = time.perf_counter(); oracle.mm_ref(A, B); dt = time.perf_counter() - t
    rows = int(min(M, max(probe_rows, probe_rows * target_s / max(dt, 1e-4))))
    rows = max(cores, rows // cores * cores)
    A = oracle.generate_weight_data(1, rows, K)
    best = None
    for _ in range(reps):
        t = time.perf_counter(); oracle.mm_ref(A, B); dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return 2.0 * rows * N * K / best / 1e12, rows, best, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    N_gpus = args.gpus
    M, N, K = (4096, 4096, 4096) if N_gpus == 1 else (16384, 16384, 16384)
    import oracle
    oracle.build()
    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = oracle.num_threads()
    # each step = one bounded sample; size it so warmup+steps finish within a few minutes
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    per_step = min(12.0, max(1.0, budget_s))
    tf, rows, dt, cores = cpu_mm_ref_sample(M, N, K, target_s=per_step)
    Bm = oracle.generate_weight_data(2, K, N)
    Am = oracle.generate_weight_data(1, rows, K)
    for _ in range(args.warmup):
        oracle.mm_ref(Am, Bm)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.mm_ref(Am, Bm)
    el = time.perf_counter() - t0
    value = 2.0 * rows * N * K * args.steps / el / 1e12
    sample = f"rows 0..{rows - 1} of C ({rows} x {N} x {K} slice of the {M}^3 product) per step, fp32 mm_ref order"
    line = {
        "impl": "reference", "metric": "sgemm_fp32_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": N_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if N_gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic U[-10,10)/50, seeded",
        "config": {"workload": f"sgemm {M}x{N}x{K} fp32 row-major", "reference_path": "mm_ref (src/harness.rs:17-28), CPU"},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# single-GPU arm
# --------------------------------------------------------------------------------------------------
def time_kernel_steps(ctx, kern, sets, steps, warmup):
    """W warm-up + K timed steps rotating over `sets` of (A,B,C); returns (total_ms, per-launch dominant-kernel ms list)."""
    for i in range(warmup):
        a, b, c = sets[i % len(sets)]
        ctx.launch(kern, a, b, c)
    ctx.sync()
    kern.profile(True)
    n0 = ctx.launch_count
    ctx.timer_begin()
    for i in range(steps):
        a, b, c = sets[i % len(sets)]
        ctx.launch(kern, a, b, c)
    total = ctx.timer_end()
    time_kernel_steps.launches = ctx.launch_count - n0  # device kernels launched inside the timed region
    per = kern.profile_read(256)
    kern.profile(False)
    return total, per


def time_back_to_back(ctx, kern, sets, iters, warmup):
    """Average duration (ms) of `iters` launches issued back to back on the stream, CUDA events around the whole batch.
    Used for the microsecond-scale GEMV kernels: they are launched with programmatic dependent launch, which an event
    record between two launches would defeat, and a per-launch event pair has ~2 us granularity."""
    for i in range(warmup):
        ctx.launch(kern, *sets[i % len(sets)])
    ctx.sync()
    ctx.timer_begin()
    for i in range(iters):
        ctx.launch(kern, *sets[i % len(sets)])
    return ctx.timer_end() / iters


def make_sets(ctx, M, N, K, nsets, seed0, quant=False, oracle=None, group_k=0):
    sets = []
    for s in range(nsets):
        a = ctx.buffer(M * K * 4); a.fill_weights(seed0 + 10 * s + 1, M * K)
        if quant:
            import wgpu_mm_b200 as w
            Wm = np.empty((K, N), dtype=np.float32)
            tmp = ctx.buffer(K * N * 4); tmp.fill_weights(seed0 + 10 * s + 2, K * N); tmp.read_into(Wm.reshape(-1)); tmp.free()
            words = w.quant.sint8_quantize_grouped(Wm, K, N, group_k) if group_k else w.quant.sint8_quantize(Wm, K, N)[0]
            b = ctx.buffer_from(words)
        else:
            b = ctx.buffer(K * N * 4); b.fill_weights(seed0 + 10 * s + 2, K * N)
        c = ctx.buffer(M * N * 4); c.fill_weights(seed0 + 10 * s + 3, M * N)
        sets.append((a, b, c))
    ctx.sync()
    return sets


def free_sets(sets):
    for t in sets:
        for b in t:
            b.free()


def run_single(args):
    import wgpu_mm_b200 as w
    peaks = load_peaks()
    ctx = w.Context(0)
    info = ctx.device_info()
    steps, warmup = args.steps, max(3, args.warmup)
    M = N = K = 4096
    flop = 2.0 * M * N * K

    # ---------------- headline: tcgen05 3xTF32 SGEMM 4096^3 ----------------
    tune = (args.tc_bn, 0, args.tc_bk, 0)
    kern = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=tune))
    sets = make_sets(ctx, M, N, K, 3, 100)
    sampler = ClockSampler(0); sampler.start(); time.sleep(0.3)
    t0 = sampler.mark()
    total_ms, per = time_kernel_steps(ctx, kern, sets, steps, warmup)
    t1 = sampler.mark()
    launches = time_kernel_steps.launches  # 2 device kernels per step: split_lo (A and B), tcgen05 GEMM
    clocks = sampler.stop(t0, t1)
    ms_per_step = total_ms / steps
    value = flop / (ms_per_step * 1e-3) / 1e12
    kern_ms = float(np.mean(per)) if per else ms_per_step
    tc_achieved = flop / (kern_ms * 1e-3) / 1e12
    # tensor roofline: tf32 runs at half the bf16 rate and the split executes 3 MMAs per product
    tc_peak = peaks["bf16_tflops"] / 2.0 / 3.0
    traffic = load_traffic()
    roofline = {"bound": "tensor", "achieved": tc_achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": tc_achieved / tc_peak,
                "traffic": traffic.get("sgemm_tc3x_kernel@4096"), "kernel": "sgemm_tc3x_kernel", "kernel_ms": kern_ms,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({peaks['_source']}, burst) / 2 (tf32:bf16 rate) / 3 (3xTF32 MMAs per product); "
                               f"tensor-pipe view: {3 * tc_achieved:.1f} of {peaks['bf16_tflops'] / 2:.1f} TF32 TFLOP/s"}

    # ---------------- e2e: host buffers through b200mm_mm_host ----------------
    import ctypes as C
    hA, hB, hC = (C.c_void_p() for _ in range(3))
    for h, nb in ((hA, M * K * 4), (hB, K * N * 4), (hC, M * N * 4)):
        w._lib.check(w.lib().b200mm_host_alloc(nb, C.byref(h)))
    npA = np.ctypeslib.as_array((C.c_float * (M * K)).from_address(hA.value))
    npB = np.ctypeslib.as_array((C.c_float * (K * N)).from_address(hB.value))
    npC = np.ctypeslib.as_array((C.c_float * (M * N)).from_address(hC.value))
    sets[0][0].read_into(npA); sets[0][1].read_into(npB)
    dA, dB, dC = sets[1]
    e2e_steps = max(3, min(steps, 10))
    for _ in range(2):
        ctx.mm_host(kern, npA, npB, npC, dA, dB, dC)
    te = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.mm_host(kern, npA, npB, npC, dA, dB, dC)  # blocking: returns when C is in host memory
    e2e_s = (time.perf_counter() - te) / e2e_steps
    e2e = {"value": flop / e2e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": M * K * 4 + K * N * 4, "d2h_bytes_per_step": M * N * 4,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "api": "b200mm_mm_host (pinned host A,B -> device, split_lo + tcgen05 GEMM, C -> pinned host; copies pipelined over 16 row panels)"}
    checksum = float(npC[:4096].astype(np.float64).sum())
    for h in (hA, hB, hC):
        w.lib().b200mm_host_free(h)
    kern.free()

    extras = {}
    if not args.no_extras:
        # ---------------- SIMT FP32 SGEMM 4096^3 (comparison point) ----------------
        ks = ctx.kernel(w.KernelId.SGEMM_SIMT, M, N, K)
        tot, per = time_kernel_steps(ctx, ks, sets, max(3, steps // 2), 3)
        ms = float(np.mean(per))
        simt_peak = info["sm_count"] * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        extras["sgemm_simt_4096"] = {"tflops": flop / (ms * 1e-3) / 1e12, "kernel_ms": ms,
                                     "roofline": {"bound": "fma", "achieved": flop / (ms * 1e-3) / 1e12, "peak": simt_peak, "unit": "TFLOP/s",
                                                  "frac": flop / (ms * 1e-3) / 1e12 / simt_peak,
                                                  "peak_source": f"{info['sm_count']} SM x 128 lanes x 2 flop x sm_max_mhz (derived, not measured)"}}
        ks.free()
    free_sets(sets)

    if not args.no_extras:
        try:  # a failing extra must never cost the headline line
            # ---------------- GEMV fp32 1x4096 * 4096x16384: 4 weight sets = 1 GiB rotated (> L2) ----------------
            Kv, Nv = 4096, 16384
            gsets = make_sets(ctx, 1, Nv, Kv, 4, 300)
            AT = int(w.Flags.AUTOTUNE)  # geometry / K-split count measured once at kernel creation (no-op when --gemv-variant is given)
            kg = ctx.kernel(w.KernelId.GEMV_F32, 1, Nv, Kv, w.KernelParams(tune=(args.gemv_variant, 0, 0, 0), flags=AT))
            ms = time_back_to_back(ctx, kg, gsets, 200, 20)
            tot = ms * 40
            gbytes = 4.0 * Kv * Nv + 4 * Kv + 4 * Nv
            extras["gemv_f32_4096x16384"] = {"gbps": gbytes / (ms * 1e-3) / 1e9, "kernel_us": ms * 1e3, "geometry": list(kg.geometry()), "timing": "200 back-to-back PDL launches, 4 weight sets (1 GiB) rotated",
                                             "roofline": {"bound": "hbm", "achieved": gbytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                          "frac": gbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "frac_of_nominal_8000_gbs": gbytes / (ms * 1e-3) / 1e9 / 8000.0, "traffic": traffic.get("gemv_stream_kernel<GemvF32>@4096x16384"),
                                                          "algorithmic_bytes": gbytes, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['_source']})"}}
            kg.free(); free_sets(gsets)
            # ---------------- qGEMV sint8 1x4096 * 4096x14336: 8 weight sets = 470 MB rotated (> L2) ----------------
            Kq, Nq = 4096, 14336
            qsets = make_sets(ctx, 1, Nq, Kq, 8, 500, quant=True)
            kq = ctx.kernel(w.KernelId.QGEMV_SINT8, 1, Nq, Kq, w.KernelParams(absmax=2.0, batch=1, tune=(args.gemv_variant, 0, 0, 0), flags=AT))
            ms = time_back_to_back(ctx, kq, qsets, 400, 40)
            tot = ms * 80
            qbytes = 1.0 * Kq * Nq + 4 * Kq + 4 * Nq
            extras["qgemv_sint8_4096x14336"] = {"gbps": qbytes / (ms * 1e-3) / 1e9, "kernel_us": ms * 1e3, "geometry": list(kq.geometry()), "timing": "400 back-to-back PDL launches, 8 weight sets (470 MB) rotated",
                                                "roofline": {"bound": "hbm", "achieved": qbytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                             "frac": qbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "frac_of_nominal_8000_gbs": qbytes / (ms * 1e-3) / 1e9 / 8000.0, "traffic": traffic.get("gemv_stream_kernel<GemvS8>@4096x14336"),
                                                             "algorithmic_bytes": qbytes, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['_source']})"}}
            kq.free(); free_sets(qsets)
            # ---------------- same shape with per-group scales (group_k = 128): weights + 1.8 MB of scales ----------------
            gk = 128
            qsets = make_sets(ctx, 1, Nq, Kq, 8, 600, quant=True, group_k=gk)
            kq = ctx.kernel(w.KernelId.QGEMV_SINT8, 1, Nq, Kq, w.KernelParams(batch=1, group_k=gk, flags=AT))
            ms = time_back_to_back(ctx, kq, qsets, 400, 40)
            gqbytes = qbytes + 4.0 * (Kq // gk) * Nq
            extras["qgemv_sint8_g128_4096x14336"] = {"gbps": gqbytes / (ms * 1e-3) / 1e9, "kernel_us": ms * 1e3, "algorithmic_bytes": gqbytes, "geometry": list(kq.geometry()),
                                                     "frac_of_hbm_peak": gqbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                                     "timing": "400 back-to-back PDL launches, 8 weight sets rotated"}
            kq.free(); free_sets(qsets)
        except Exception as exc:  # noqa: BLE001
            extras.setdefault("errors", {})["gemv"] = repr(exc)

    if not args.no_extras:
        try:  # a failing extra must never cost the headline line
            # ---------------- SGEMM 16384^3 on ONE GPU: the strong-scaling baseline of the N-sharded runs ----------------
            Mb = 16384
            kb = ctx.kernel(w.KernelId.SGEMM_TC3X, Mb, Mb, Mb, w.KernelParams(tune=tune))
            bsets = make_sets(ctx, Mb, Mb, Mb, 1, 700)
            tot, per = time_kernel_steps(ctx, kb, bsets, 3, 2)
            bflop = 2.0 * Mb * Mb * Mb
            extras["sgemm_tc3x_16384_1gpu"] = {"tflops": bflop / (tot / 3 * 1e-3) / 1e12, "ms_per_step": tot / 3, "kernel_ms": float(np.mean(per)),
                                               "note": "same kernel and schedule as the N-sharded runs; sustained clocks (power cap) apply"}
            kb.free(); free_sets(bsets)
        except Exception as exc:  # noqa: BLE001
            extras.setdefault("errors", {})["sgemm_16384_1gpu"] = repr(exc)

    if not args.no_extras:
        try:  # a failing extra must never cost the headline line
            # ---------------- BASELINE configs[0]: the reference's own test shape (1024^3) through the host harness ----------------
            # verify (max-abs-err <= 1e-3 vs mm_ref) -> 8 warm-up -> 10 timed launches + read-back, exactly src/harness.rs:170-248;
            # "gflops" is the reference-style number (wall clock incl. the D2H of C), "kernel_gflops" the CUDA-event one
            from wgpu_mm_b200 import harness as hz
            h = {}
            for entry in ("gemm_wonnx", "gemm_5", "sgemm_simt", "sgemm_tc3x"):
                r = hz.test_harness(None, entry, (1024, 1024, 1024), False)
                h[entry] = {"reference_style_gflops": r.gflops, "kernel_gflops": r.kernel_gflops, "max_abs_err": r.max_abs_err,
                            "max_rel_err_f64": r.max_rel_err_f64}
            r = hz.test_harness(None, "qgemv_1", (1, 1024, 1024), True)
            h["qgemv_1"] = {"kernel_gbps": r.kernel_gbps, "max_abs_err": r.max_abs_err}
            extras["harness_1024_reference_shapes"] = h
        except Exception as exc:  # noqa: BLE001
            extras.setdefault("errors", {})["harness_1024"] = repr(exc)

    # ---------------- CPU baseline (reported, not the target) ----------------
    cpu = None
    if not args.no_cpu:
        tf, rows, dt, cores = cpu_mm_ref_sample(M, N, K, target_s=12.0)
        cpu = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port",
               "sample": f"mm_ref (src/harness.rs:17-28 restated, OpenMP) on rows 0..{rows - 1} of the 4096^3 product, {dt:.1f} s"}
        # the gemm.wgsl (+gemm_macro.wgsl) per-invocation restatement at 1024^3 on all host cores --
        # the stand-in for "the WGSL path on a software Vulkan adapter", which cannot run here (no wgpu / lavapipe)
        import oracle
        A1 = oracle.generate_weight_data(1, 1024, 1024); B1 = oracle.generate_weight_data(2, 1024, 1024)
        oracle.wgsl_gemm("gemm_wonnx", A1, B1)
        t = time.perf_counter()
        for _ in range(5):
            oracle.wgsl_gemm("gemm_wonnx", A1, B1)
        dtw = (time.perf_counter() - t) / 5
        cpu["gemm_wgsl_restatement_1024_gflops"] = 2.0 * 1024 ** 3 / dtw / 1e9

    line = {
        "metric": "sgemm_fp32_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32x3 (fp32-accurate, fp32 accumulate)",
        "data": "synthetic U[-10,10)/50, seeded, generated on device",
        "config": {"workload": "sgemm 4096x4096x4096 fp32 row-major (BASELINE configs[1])", "kernel": "sgemm_tc3x (split_lo + tcgen05 GEMM per step)",
                   "tile": f"128x{args.tc_bn}x{args.tc_bk or 16}", "l2": "3 rotating (A,B,C) sets = 576 MiB of operands, larger than the 126 MB L2",
                   "device": info["name"], "sm_count": info["sm_count"]},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "extras": extras, "c_checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    ctx.close()


# --------------------------------------------------------------------------------------------------
# multi-GPU arm: 16384^3, N-sharded, one process per GPU
# --------------------------------------------------------------------------------------------------
def run_multi(args):
    import torch
    import torch.distributed as dist
    import wgpu_mm_b200 as w
    from wgpu_mm_b200 import shard

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    ctx = w.Context(local)
    steps, warmup = args.steps, max(3, args.warmup)
    M = N = K = args.size
    plan = shard.ShardPlan(N, world, rank)
    job = shard.ShardedSgemm(ctx, M, N, K, plan, mode=args.mode, kernel_id=w.KernelId.SGEMM_TC3X, seed=100, tc_bn=args.tc_bn)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        job.step()
    job.barrier()
    t0 = sampler.mark()
    l0 = ctx.launch_count
    ctx.timer_begin()
    for _ in range(steps):
        job.step()
    ms = ctx.timer_end()
    job.barrier()
    t1 = sampler.mark()
    launches = ctx.launch_count - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    per = job.kernel_times()
    flop = 2.0 * M * N * K
    value = flop * steps / (ms * 1e-3) / 1e12
    # every rank must hold the same full C: compare a row checksum across ranks (no oracle involved)
    cs = torch.tensor([job.checksum()], dtype=torch.float64, device="cuda")
    cs_all = [torch.zeros_like(cs) for _ in range(world)]
    dist.all_gather(cs_all, cs)
    consistent = all(float(c.item()) == float(cs_all[0].item()) for c in cs_all) and float(cs_all[0].item()) != 0.0
    e2e_s, h2d, d2h = job.e2e(2) if not args.no_e2e else (None, None, None)
    job.close()
    # ---- the GEMV configs, N-sharded over the same ranks (BASELINE configs[2], [3]) ----
    extras = {}
    if not args.no_extras:
        for name, Kv, Nv, quant in (("gemv_f32_4096x16384", 4096, 16384, False), ("qgemv_sint8_4096x14336", 4096, 14336, True)):
            if Nv % (16 * world):
                continue
            gplan = shard.ShardPlan(Nv, world, rank)
            gj = shard.ShardedGemv(ctx, Kv, Nv, gplan, quant=quant, mode=args.mode)
            for _ in range(10):
                gj.step()
            gj.barrier()
            gj.kernel_times()
            ctx.timer_begin()
            for _ in range(50):
                gj.step()
            gms = ctx.timer_end()
            gj.barrier()
            kt = gj.kernel_times()
            tt = torch.tensor([gms / 50, float(np.median(kt))], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_bytes = (Kv * Nv if quant else 4 * Kv * Nv) + 4 * Kv + 4 * Nv
            step_us, kern_us = float(tt[0]) * 1e3, float(tt[1]) * 1e3
            extras[name] = {"kernel_gbps_aggregate": total_bytes / (kern_us * 1e-6) / 1e9, "kernel_us_max_over_ranks": kern_us,
                            "step_us_incl_gather": step_us, "step_gbps_aggregate": total_bytes / (step_us * 1e-6) / 1e9,
                            "roofline": {"bound": "hbm", "achieved": total_bytes / (kern_us * 1e-6) / 1e9 / world, "peak": peaks["hbm_gbs"],
                                         "unit": "GB/s per GPU", "frac": total_bytes / (kern_us * 1e-6) / 1e9 / world / peaks["hbm_gbs"]}}
            gj.close()
    if rank == 0:
        clocks = sampler.stop(t0, t1)
        kern_ms = float(np.mean(per)) if per else ms / steps
        per_gpu = flop / world / (kern_ms * 1e-3) / 1e12
        tc_peak = peaks["bf16_tflops_sustained"] / 2.0 / 3.0
        line = {
            "metric": "sgemm_fp32_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "tf32x3 (fp32-accurate, fp32 accumulate)", "data": "synthetic U[-10,10)/50, seeded, generated on device",
            "config": {"workload": f"sgemm {M}x{N}x{K} fp32, B and C N-sharded over {world} GPUs (BASELINE configs[4])", "parallelism": f"n-shard x{world}",
                       "gather": args.mode, "l2": "per-rank operands (A 1 GiB + B panel) exceed the 126 MB L2"},
            "roofline": {"bound": "tensor", "achieved": per_gpu, "peak": tc_peak, "unit": "TFLOP/s", "frac": per_gpu / tc_peak, "traffic": None,
                         "kernel": "sgemm_tc3x_kernel (per GPU)", "kernel_ms": kern_ms,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['_source']}) / 2 / 3"},
            "cpu_baseline": None,
            "e2e": ({"value": flop / e2e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "ms_per_step": e2e_s * 1e3, "api": "per rank: pinned host A + B panel -> device, sharded step, full C -> pinned host"}
                    if e2e_s else None),
            "gpu_launches": int(launches), "clocks": clocks, "ranks_hold_identical_c": bool(consistent), "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fused", choices=["fused", "nccl"], help="multi-GPU gather: peer stores from the epilogue, or NCCL all-gather")
    ap.add_argument("--size", type=int, default=16384, help="multi-GPU problem size (M=N=K)")
    ap.add_argument("--tc-bn", type=int, default=256, choices=[128, 256])
    ap.add_argument("--tc-bk", type=int, default=0, choices=[0, 16, 32], help="k-block of the tcgen05 kernel (0 = library default)")
    ap.add_argument("--gemv-variant", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        if world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
        return run_multi(args)
    return run_single(args)


if __name__ == "__main__":
    main()
