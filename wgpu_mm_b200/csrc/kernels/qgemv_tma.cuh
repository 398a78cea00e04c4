// TMA-staged sint8 GEMV: y[1 x N] = x[1 x K] * dequant(Wq[K x N]) for the src/quant.rs packing (replaces
// shaders/gemv/qgemv_1.wgsl:10-39 at batch 1; the register-streaming kernel in gemv.cuh remains for batches and as
// the fp32 path).
//
// Why a second kernel: the sint8 shape of BASELINE config 3 is only 58.7 MB (8.4 us of streaming at 7 TB/s) and
// needs ~6 us of dequantisation issue time.  With register-staged loads the same warps alternate between waiting for
// DRAM and dequantising, so the two serialise (measured 15.6 us).  Here ONE producer thread streams 64-row x 256-byte
// weight tiles into a shared-memory ring with TMA (cp.async.bulk.tensor, mbarrier full/empty), so up to STAGES x 16 KB
// per CTA stay in flight no matter what the 8 consumer warps are doing, and the consumers spend their issue slots on
// PRMT / FADD2 / FFMA2 only.  Grid = (N/256 panels, K-splits <= 8); the K-splits of a panel form a thread-block
// cluster and are reduced in rank order through distributed shared memory; launched with programmatic dependent
// launch: the weight tiles of the next GEMV are requested before griddepcontrol.wait.
//
// Measured on B200 (tools/sweep_gemv.py, variant 200): 15.8 us = 3.71 TB/s, the same as the register-streaming kernel
// (15.6 us).  The K-scaling law t = 6.9 us + bytes / 6.7 TB/s holds for both, i.e. the shape is bound by per-launch
// fixed cost (grid launch, first-tile latency, cluster reduction tail) rather than by how the bytes are moved, so this
// kernel is kept as a selectable variant (tune[0] = 200) and not the default.
#pragma once
#include "gemv.cuh"
#include "sgemm_tc3x.cuh"  // ptx:: mbarrier / TMA wrappers

namespace b200mm {

struct QgemvTmaCfg {
    static constexpr int CONSUMER_WARPS = 8, THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr int PANEL = 256;          // columns = bytes per tile row
    static constexpr int STAGE_ROWS = 64, STAGES = 4;
    static constexpr uint32_t STAGE_BYTES = STAGE_ROWS * PANEL;
};

__global__ void __launch_bounds__(QgemvTmaCfg::THREADS)
qgemv_tma_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ x, float* __restrict__ y, int K, int N,
                 int rows_per_split, float out_scale, const __grid_constant__ PeerStore peers) {
    using Cfg = QgemvTmaCfg;
    constexpr int WARPS = Cfg::CONSUMER_WARPS, PANEL = Cfg::PANEL, SR = Cfg::STAGE_ROWS, STAGES = Cfg::STAGES;
    constexpr int COLS = 16, LPR = 16;  // 16 lanes x 16 B cover one 256-byte row; a warp takes 2 rows per step
    extern __shared__ __align__(128) uint8_t smem_q[];
    const uint32_t ring = (smem_u32(smem_q) + 127u) & ~127u;
    uint8_t* ring_gen = smem_q + (ring - smem_u32(smem_q));
    const uint32_t bar_base = ring + STAGES * Cfg::STAGE_BYTES;  // full[STAGES], empty[STAGES]
    float* xs = reinterpret_cast<float*>(ring_gen + STAGES * Cfg::STAGE_BYTES + 16 * STAGES);
    // the reduction scratch aliases the (by then drained) tile ring, so three CTAs fit in one SM's shared memory
    float* red = reinterpret_cast<float*>(ring_gen);  // WARPS * PANEL floats
    float* cta_part = red + WARPS * PANEL;             // PANEL floats

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int panel = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    const int k_beg = split * rows_per_split, k_end = min(K, k_beg + rows_per_split);
    const int n_stages = (k_end - k_beg + SR - 1) / SR;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(bar_base + 8u * s, 1);
            ptx::mbar_init(bar_base + 8u * (STAGES + s), WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == WARPS) {
        // ===================== producer =====================
        if (lane == 0) {
            ptx::prefetch_tmap(&tmW);
            for (int s = 0; s < n_stages; ++s) {
                const int slot = s % STAGES;
                const uint32_t phase = (s / STAGES) & 1;
                ptx::mbar_wait(bar_base + 8u * (STAGES + slot), phase ^ 1);
                ptx::mbar_arrive_expect_tx(bar_base + 8u * slot, Cfg::STAGE_BYTES);
                // rows beyond K are zero-filled by TMA; rows beyond k_end (next split's) are skipped by the consumers
                ptx::tma_load_2d(ring + slot * Cfg::STAGE_BYTES, &tmW, bar_base + 8u * slot, panel * PANEL, k_beg + s * SR);
            }
        }
    } else {
        // ===================== consumers =====================
        const int lir = lane % LPR, riw = lane / LPR;
        asm volatile("griddepcontrol.wait;" ::: "memory");  // x (and y) may come from the previous kernel
        for (int i = tid; i < rows_per_split; i += WARPS * 32) xs[i] = (k_beg + i < k_end) ? x[k_beg + i] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float acc[COLS];
#pragma unroll
        for (int j = 0; j < COLS; ++j) acc[j] = 0.f;
        for (int s = 0; s < n_stages; ++s) {
            const int slot = s % STAGES;
            const uint32_t phase = (s / STAGES) & 1;
            ptx::mbar_wait(bar_base + 8u * slot, phase);
            const uint8_t* tile = ring_gen + slot * Cfg::STAGE_BYTES;
            const int r0 = s * SR + warp * 2 + riw;  // row index inside the split
            uint4 w[SR / (2 * WARPS)];
#pragma unroll
            for (int i = 0; i < SR / (2 * WARPS); ++i)
                w[i] = *reinterpret_cast<const uint4*>(tile + (size_t)(warp * 2 + riw + i * 2 * WARPS) * PANEL + lir * 16);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_base + 8u * (STAGES + slot));  // tile is in registers: release the slot
#pragma unroll
            for (int i = 0; i < SR / (2 * WARPS); ++i) {
                const int r = r0 + i * 2 * WARPS;
                if (k_beg + r < k_end) GemvS8::fma(acc, w[i], xs[r]);
            }
        }
        // rows-in-warp -> one partial per column, then warps -> CTA partial (fixed orders)
#pragma unroll
        for (int j = 0; j < COLS; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
        asm volatile("bar.sync 1, 256;" ::: "memory");  // every consumer has its last tile in registers: the ring is free
        if (riw == 0) {
#pragma unroll
            for (int j = 0; j < COLS; ++j) red[(warp * COLS + j) * LPR + lir] = acc[j];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = tid; c < PANEL; c += WARPS * 32) {
            float s = 0.f;
#pragma unroll
            for (int wv = 0; wv < WARPS; ++wv) s += red[(wv * COLS + c % COLS) * LPR + c / COLS];
            cta_part[c] = s;
        }
    }
    // K-splits -> y inside the cluster (rank = split), see gemv.cuh
    if (splits > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    if (split == 0 && warp < WARPS) {
        const uint32_t local = smem_u32(cta_part);
        for (int c = tid; c < PANEL; c += WARPS * 32) {
            const int gc = panel * PANEL + c;
            float s = 0.f;
            for (int r = 0; r < splits; ++r) {
                uint32_t remote;
                float v;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local + 4u * c), "r"(r));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
                s += v;
            }
            if (gc < N) store_y(y, gc, s * out_scale, peers);
        }
    }
    if (splits > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace b200mm
