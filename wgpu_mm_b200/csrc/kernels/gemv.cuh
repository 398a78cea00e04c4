// HBM-streaming GEMV kernels: y[1 x N] = x[1 x K] * W[K x N], W row-major (N contiguous), for
//   - fp32 weights  (reference semantics: mm_ref with M == 1, src/harness.rs:17-28; there is no fp32
//                    GEMV shader in the reference), and
//   - sint8 weights packed 4 per u32 along N in the src/quant.rs:20-26 format, dequantised in registers
//                   (replaces shaders/gemv/qgemv_1.wgsl:10-39, including its global_id.y batch offsets :12-14).
//
// The reference shader gives each invocation 4 outputs and lets it walk all of K serially, so only N/4
// threads exist and the kernel is latency-bound.  Here the matrix is cut into
// column panels x K-splits so that >= 2 CTAs per SM each keep UNROLL independent 128-bit loads in
// flight per thread; a warp covers one contiguous 512 B (or 2 x 256 B) row segment per load, the CTA's
// warps take interleaved rows, and the K-splits are combined in a FIXED order by the last CTA to
// finish a panel (atomic ticket), so the result is deterministic run to run.  No tensor cores: the
// path is bandwidth-bound (algorithmic bytes = weights + x + y).
//
// Dequantisation: q in [-127,127] -> float by a byte-permute into the mantissa of 2^23 (the I2F pipe
// would be the bottleneck at HBM rate), then packed FADD2 / FFMA2 on column pairs, accumulation of
// x[k]*q in fp32, and ONE multiply by absmax/127 per output at the end.  The reference multiplies every weight
// by absmax first (q/127*absmax, src/quant.rs:36-39); factoring the scale out changes the rounding by
// <= 2 ulp per output, far inside the 1e-3 gate, and is reported against FP64 in the tests.
#pragma once
#include "common.cuh"
#include "sgemm_simt.cuh"  // PeerStore

namespace b200mm {

struct GemvF32 {  // one 128-bit load = 4 columns
    using Vec = float4;
    static constexpr int COLS = 4;
    static __device__ __forceinline__ Vec load(const void* p) { return ldg_stream_f4(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void fma(float (&acc)[COLS], const Vec& w, float xk) {
        acc[0] = fmaf(xk, w.x, acc[0]);
        acc[1] = fmaf(xk, w.y, acc[1]);
        acc[2] = fmaf(xk, w.z, acc[2]);
        acc[3] = fmaf(xk, w.w, acc[3]);
    }
};

struct GemvS8 {  // one 128-bit load = 16 int8 = 16 columns
    using Vec = uint4;
    static constexpr int COLS = 16;
    static __device__ __forceinline__ Vec load(const void* p) { return ldg_stream_u4(reinterpret_cast<const uint4*>(p)); }
    // Blackwell packed fp32: one FADD2 / FFMA2 handles two columns (SASS: FADD2, FFMA2), so a word of 4 weights
    // costs 1 LOP3 + 4 PRMT + 2 FADD2 + 2 FFMA2 = 2.25 instructions per element instead of 3.25.
    static __device__ __forceinline__ void add2(float& a0, float& a1, float m) {
        asm("{\n\t.reg .b64 ra, rm;\n\tmov.b64 ra, {%0,%1};\n\tmov.b64 rm, {%2,%2};\n\tadd.rn.f32x2 ra, ra, rm;\n\tmov.b64 {%0,%1}, ra;\n\t}"
            : "+f"(a0), "+f"(a1)
            : "f"(m));
    }
    static __device__ __forceinline__ void fma2(float& a0, float& a1, float x, float q0, float q1) {
        asm("{\n\t.reg .b64 ra, rx, rq;\n\tmov.b64 ra, {%0,%1};\n\tmov.b64 rx, {%2,%2};\n\tmov.b64 rq, {%3,%4};\n\t"
            "fma.rn.f32x2 ra, rx, rq, ra;\n\tmov.b64 {%0,%1}, ra;\n\t}"
            : "+f"(a0), "+f"(a1)
            : "f"(x), "f"(q0), "f"(q1));
    }
    static __device__ __forceinline__ void fma4(float* acc, uint32_t w, float xk) {
        // bytes are two's complement; flip the sign bit so each byte is q+128 in [1,255], drop it into the
        // mantissa of 2^23 and subtract 2^23+128: exact integer -> float without the I2F pipe.
        const uint32_t u = w ^ 0x80808080u;
        const float neg_magic = -8388736.0f;  // -(2^23 + 128)
        float q0 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7540));
        float q1 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7541));
        float q2 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7542));
        float q3 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7543));
        add2(q0, q1, neg_magic);
        add2(q2, q3, neg_magic);
        fma2(acc[0], acc[1], xk, q0, q1);
        fma2(acc[2], acc[3], xk, q2, q3);
    }
    static __device__ __forceinline__ void fma(float (&acc)[COLS], const Vec& w, float xk) {
        fma4(acc + 0, w.x, xk);
        fma4(acc + 4, w.y, xk);
        fma4(acc + 8, w.z, xk);
        fma4(acc + 12, w.w, xk);
    }
};

// N-sharded multi-GPU GEMV: this rank's y slice is also written to the same position of the full y on every
// peer (CUDA-IPC mappings over NVLink) -- the fused form of the all-gather of y slices.
__device__ __forceinline__ void store_y(float* y, int gc, float v, const PeerStore& peers) {
    if (peers.world == 0) {
        y[gc] = v;
    } else {
        for (int d = 0; d < peers.world; ++d) peers.c[d][peers.col0 + gc] = v;
    }
}

// Debug timeline (tools/trace_gemv.py): when the upper bits of `use_cluster` carry a slot number, `tickets` is a trace buffer and
// every CTA records globaltimer stamps at its milestones: [slot][cta][8] u64.  Zero cost when no slot is given (one predicate).
__device__ __forceinline__ void gemv_trace(unsigned long long* t, int idx) {
    if (t != nullptr && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        t[idx] = now;
    }
}

// Launch: grid (panels, splits, batch), block WARPS*32.
//   LPR   lanes per row segment (32 or 16): a warp reads 32/LPR rows per load instruction
//   panel = LPR * COLS columns;   rows of a split are dealt round-robin to (warp, row-in-warp) slots.
// partial: [batch][splits][N] floats, tickets: [batch][panels] u32 (zeroed once; self-resetting).
// MROWS > 1: skinny GEMM -- MROWS rows of x (row-major MROWS x K) share one pass over W; y is MROWS x N.
// GROUPED (sint8 only): the K*N weight bytes are followed by ceil(K/group_k) x N f32 scales, one per
//   (block of group_k rows, column).  A thread's pipeline window of 2*UNROLL*RSTEP rows is aligned so that it never
//   straddles a group (the host enforces group_k % window == 0 and rows_per_split % window == 0); the integer-valued
//   partial dot product of a group is folded into the running total with ONE fma per column when the group changes, so
//   the inner loop is the ungrouped one.  The CTA's scales (<= rows_per_split/group_k + 1 groups x PANEL columns) are
//   staged in shared memory up front, off the critical path.
// BLOCKED (grouped scales only): warp w streams the CONTIGUOUS rows [w, w + 1) * rows_per_split / WARPS of the CTA's slab instead of
// every WARPS-th group of RPW rows.  A thread's rows then stay inside one quantisation group for group_k / RPW loads instead of
// 2 * UNROLL, so the per-group fold (scale * partial into the running total) runs once per group and warp -- for cfg4 with
// group_k = 128 and 1024-row slabs exactly once, at the end -- instead of after every pipeline window.  The loop still counts in the
// interleaved coordinate kk; prow() maps it to the physical row.  Needs whole splits (K % rows_per_split == 0, host-checked).
template <class T, int WARPS, int UNROLL, int LPR, int MROWS = 1, bool GROUPED = false, int MINB = 1, int EAGER_ = -1, bool BLOCKED = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
gemv_stream_kernel(const float* __restrict__ x, const void* __restrict__ W, float* __restrict__ y,
                   float* __restrict__ partial, unsigned int* __restrict__ tickets, int K, int N, int rows_per_split,
                   float out_scale, size_t x_batch_stride, size_t w_batch_stride_bytes, size_t y_batch_stride,
                   const __grid_constant__ PeerStore peers, int use_cluster, int group_k) {
    static_assert(!GROUPED || MROWS == 1, "grouped scales: single-row kernel only");
    static_assert(!BLOCKED || GROUPED, "the blocked row order exists for the grouped-scale kernels");
    constexpr int COLS = T::COLS;
    constexpr int RPW = 32 / LPR;        // rows per warp per load
    constexpr int RSTEP = WARPS * RPW;   // rows per CTA per load
    constexpr int PANEL = LPR * COLS;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                                 // MROWS * rows_per_split floats
    float* red = sm + MROWS * rows_per_split;       // WARPS * MROWS * PANEL floats
    __shared__ unsigned int s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lir = lane % LPR, riw = lane / LPR;
    const int panel = blockIdx.x, split = blockIdx.y, batch = blockIdx.z, splits = gridDim.y;
    // Balanced column panels: the N / COLS column groups (one 128-bit load each) are dealt to the gridDim.x panels as evenly as
    // possible, panel p = groups [p*G/P, (p+1)*G/P).  With the natural panel count ceil(G / LPR) this is the plain LPR-group
    // partition; with MORE panels (the host's choice, e.g. exactly one CTA slot per panel x split so that every SM carries the same
    // load) the panels are ragged and a few lanes of each row segment idle.  N % COLS == 0 is required (N % 4 == 0 in the
    // reference); the host guarantees ceil(G / P) <= LPR.
    const int n_groups = N / COLS;
    const int grp0 = (int)((long long)panel * n_groups / (int)gridDim.x);
    const int pcol0 = grp0 * COLS;                                                                  // first column of this panel
    const int pcols = ((int)((long long)(panel + 1) * n_groups / (int)gridDim.x) - grp0) * COLS;    // its width (<= PANEL)
    const int col = pcol0 + lir * COLS;
    const bool col_ok = lir * COLS < pcols;

    x += batch * x_batch_stride;
    y += batch * y_batch_stride;
    const char* Wb = reinterpret_cast<const char*>(W) + batch * w_batch_stride_bytes;
    const size_t row_pitch = (size_t)N / COLS * sizeof(typename T::Vec);  // bytes per weight row

    const int k_beg = split * rows_per_split;
    const int k_end = min(K, k_beg + rows_per_split);
    const char* wp = Wb + (size_t)col / COLS * sizeof(typename T::Vec);
    int k = k_beg + warp * RPW + riw;
    // physical weight row of the interleaved coordinate kk (kk - k is a multiple of RSTEP = WARPS * RPW)
    const int k0 = k, p0 = k_beg + warp * (rows_per_split / WARPS) + riw;
    auto prow = [&](int kk) -> int {
        if constexpr (BLOCKED) return p0 + (int)((unsigned)(kk - k0) / (unsigned)WARPS);
        else return kk;
    };
    constexpr int PSTEP = BLOCKED ? RPW : RSTEP;  // physical rows between consecutive loads of a thread

    // Software pipeline: the loads of batch i+1 are issued before batch i is consumed, and the very first batch is
    // issued BEFORE x is staged, so its DRAM latency overlaps the staging barrier (every serialized ~1 us matters in
    // a kernel whose ideal duration is ~10 us).
    // Programmatic dependent launch: let the next kernel in the stream start as soon as SM resources free up, and do
    // not touch anything a previous kernel may still produce (x, partials, tickets, y) before griddepcontrol.wait.
    // The weights never depend on the previous kernel, so their first loads overlap its tail (no-ops without PDL).
    unsigned long long* trace = nullptr;
    if (use_cluster >> 8) {
        const size_t ctas = (size_t)gridDim.x * gridDim.y * gridDim.z;
        const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        trace = reinterpret_cast<unsigned long long*>(tickets) + ((size_t)((use_cluster >> 8) - 1) * ctas + cta) * 8;
        use_cluster &= 0xff;
    }
    gemv_trace(trace, 0);  // CTA started
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    typename T::Vec wa[UNROLL], wb[UNROLL];
    auto issue = [&](typename T::Vec (&w)[UNROLL], int kk) {
        const int pk = prow(kk);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (col_ok && kk + u * RSTEP < k_end) w[u] = T::load(wp + (size_t)(pk + u * PSTEP) * row_pitch);
    };
    // fp32: both register buffers are in flight before anything dependent is touched (58.7 MB shape: 12.8 -> 10.7 us).
    // sint8 keeps one (the longer live ranges cost it an occupancy step: measured 15.6 -> 19 us with both).
    constexpr bool EAGER = EAGER_ < 0 ? (T::COLS == 4) : (EAGER_ != 0);
    issue(wa, k);
    if constexpr (EAGER) issue(wb, k + UNROLL * RSTEP);
    // (Measured and rejected, round 2: prefetch.global.L2 of the CTA's whole weight slab at this point.  The median CTA finished
    // streaming 1.5 us earlier, but the early-resident CTAs' prefetches competed with the draining launch and the slowest CTAs
    // got slower: 12.9 -> 14.5 us at cfg4, profiles/r2_gemv_timeline.txt.)
    // grouped scales of this CTA's rows x columns -> shared memory (weights: independent of the previous kernel)
    float* ss = red + (WARPS * MROWS + 8 * MROWS) * PANEL;  // behind the 8 receive slots of the cluster reduction
    const int g0 = GROUPED ? k_beg / group_k : 0;
    const int ngl = GROUPED ? rows_per_split / group_k + 2 : 0;
    if constexpr (GROUPED) {
        const float* sc = reinterpret_cast<const float*>(Wb + (size_t)K * N);
        const int groups = (K + group_k - 1) / group_k;
        for (int i = tid; i < ngl * PANEL; i += WARPS * 32) {
            const int g = g0 + i / PANEL, gc = pcol0 + i % PANEL;
            ss[i] = (g < groups && i % PANEL < pcols) ? __ldg(sc + (size_t)g * N + gc) : 0.f;
        }
    }
    gemv_trace(trace, 1);  // first loads issued
    asm volatile("griddepcontrol.wait;" ::: "memory");
    gemv_trace(trace, 2);  // previous grid complete
    // split arrive / wait: tells the cluster this CTA is running (its shared memory may be written remotely from here on);
    // the matching wait sits in front of the first remote store, ~10 us later
    if (use_cluster && gridDim.y > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    // N-sharded chain with deferred completion: x may derive from the y the peers stored during the PREVIOUS launch
    if (peers.world > 1 && peers.flags[0] != nullptr && peers.deferred && peers.epoch > 1) {
        peer_wait_epoch(peers, peers.epoch - 1);
        __syncthreads();
    }

    // x -> shared memory with 128-bit loads (K % 4 == 0, splits start on multiples of 4): this sits between the PDL wait and the
    // first FMA, so a scalar strided loop (rows_per_split / threads dependent L2 round trips) was ~1 us of a ~12 us kernel
#pragma unroll
    for (int m = 0; m < MROWS; ++m) {
        const float4* xg = reinterpret_cast<const float4*>(x + (size_t)m * K + k_beg);
        float4* xs4 = reinterpret_cast<float4*>(xs + m * rows_per_split);
        const int n4 = rows_per_split >> 2, valid4 = (k_end - k_beg) >> 2;  // k_end - k_beg is a multiple of 4
        for (int i = tid; i < n4; i += WARPS * 32) xs4[i] = i < valid4 ? __ldg(xg + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    gemv_trace(trace, 3);  // x staged

    float acc[MROWS][COLS];
#pragma unroll
    for (int m = 0; m < MROWS; ++m)
#pragma unroll
        for (int j = 0; j < COLS; ++j) acc[m][j] = 0.f;
    // GROUPED: acc[0] is the current group's partial; the scaled running total lives in shared memory ([j/4][thread] float4
    // slots, conflict-free) so that the streaming loop keeps the register budget -- and the occupancy -- of the ungrouped kernel.
    float4* tot4 = reinterpret_cast<float4*>(ss + ngl * PANEL);
    int cur_g = 0;
    if constexpr (GROUPED) {
#pragma unroll
        for (int j = 0; j < COLS / 4; ++j) tot4[j * (WARPS * 32) + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        cur_g = prow(k) / group_k;
    }
    auto fold = [&](int g) {
        if constexpr (GROUPED) {
            const float4* s4 = reinterpret_cast<const float4*>(ss + min(max(g - g0, 0), ngl - 1) * PANEL + lir * COLS);
#pragma unroll
            for (int j = 0; j < COLS / 4; ++j) {
                float4 t = tot4[j * (WARPS * 32) + tid];
                const float4 sc = s4[j];
                t.x = fmaf(acc[0][4 * j + 0], sc.x, t.x);
                t.y = fmaf(acc[0][4 * j + 1], sc.y, t.y);
                t.z = fmaf(acc[0][4 * j + 2], sc.z, t.z);
                t.w = fmaf(acc[0][4 * j + 3], sc.w, t.w);
                tot4[j * (WARPS * 32) + tid] = t;
                acc[0][4 * j + 0] = acc[0][4 * j + 1] = acc[0][4 * j + 2] = acc[0][4 * j + 3] = 0.f;
            }
        }
    };
    // entering the window whose first row (for this thread) is kk: warp-uniform branch.  Windows advance by 128 rows <= group_k,
    // so at most one group boundary is crossed per step and a compare against the next boundary replaces the division.
    int next_boundary = GROUPED ? (cur_g + 1) * group_k : 0;
    auto group_step = [&](int kk) {
        if constexpr (GROUPED) {
            if (kk >= next_boundary) {
                fold(cur_g);
                ++cur_g;
                next_boundary += group_k;
            }
        }
    };

    auto consume = [&](const typename T::Vec (&w)[UNROLL], int kk) {
        const int pk = prow(kk);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (col_ok && kk + u * RSTEP < k_end) {
#pragma unroll
                for (int m = 0; m < MROWS; ++m) T::fma(acc[m], w[u], xs[m * rows_per_split + pk + u * PSTEP - k_beg]);
            }
    };
    // Fast path: while a whole double batch (and the loads it issues ahead) is in range, no per-vector bounds
    // bookkeeping -- in the sint8 kernel that bookkeeping was ~14 of 52 instructions per 16 weights.
    auto issue_fast = [&](typename T::Vec (&w)[UNROLL], int kk) {
        const int pk = prow(kk);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) w[u] = T::load(wp + (size_t)(pk + u * PSTEP) * row_pitch);
    };
    auto consume_fast = [&](const typename T::Vec (&w)[UNROLL], int kk) {
        const int pk = prow(kk);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int m = 0; m < MROWS; ++m) T::fma(acc[m], w[u], xs[m * rows_per_split + pk + u * PSTEP - k_beg]);
        }
    };
    if (col_ok) {
        constexpr int AHEAD = (EAGER ? 4 : 3) * UNROLL - 1;  // furthest row-step touched by one fast iteration
        for (; k + AHEAD * RSTEP < k_end; k += 2 * UNROLL * RSTEP) {
            group_step(prow(k));
            if constexpr (EAGER) {
                consume_fast(wa, k);
                issue_fast(wa, k + 2 * UNROLL * RSTEP);
                consume_fast(wb, k + UNROLL * RSTEP);
                issue_fast(wb, k + 3 * UNROLL * RSTEP);
            } else {
                issue_fast(wb, k + UNROLL * RSTEP);
                consume_fast(wa, k);
                issue_fast(wa, k + 2 * UNROLL * RSTEP);
                consume_fast(wb, k + UNROLL * RSTEP);
            }
        }
    }
    // guarded remainder (same buffer invariants: wa holds rows k.., and for EAGER wb holds rows k + UNROLL*RSTEP..)
    for (; k < k_end; k += 2 * UNROLL * RSTEP) {
        group_step(prow(k));
        if constexpr (EAGER) {
            consume(wa, k);
            issue(wa, k + 2 * UNROLL * RSTEP);
            consume(wb, k + UNROLL * RSTEP);
            issue(wb, k + 3 * UNROLL * RSTEP);
        } else {
            issue(wb, k + UNROLL * RSTEP);
            consume(wa, k);
            issue(wa, k + 2 * UNROLL * RSTEP);
            consume(wb, k + UNROLL * RSTEP);
        }
    }

    gemv_trace(trace, 4);  // warp 0 finished streaming
    if constexpr (GROUPED) {
        fold(cur_g);
#pragma unroll
        for (int j = 0; j < COLS / 4; ++j) {
            const float4 t = tot4[j * (WARPS * 32) + tid];
            acc[0][4 * j + 0] = t.x;
            acc[0][4 * j + 1] = t.y;
            acc[0][4 * j + 2] = t.z;
            acc[0][4 * j + 3] = t.w;
        }
    }

    // rows-in-warp -> one partial per column (lanes lir, lir+LPR, ... hold the same columns)
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
        for (int m = 0; m < MROWS; ++m)
#pragma unroll
            for (int j = 0; j < COLS; ++j) acc[m][j] += __shfl_xor_sync(0xffffffffu, acc[m][j], off);
    }
    if (riw == 0) {
#pragma unroll
        for (int m = 0; m < MROWS; ++m)
#pragma unroll
            for (int j = 0; j < COLS; ++j) red[((warp * MROWS + m) * COLS + j) * LPR + lir] = acc[m][j];  // [warp][m][j][lane]: conflict-free
    }
    __syncthreads();

    // warps -> CTA partial (fixed order), then K-splits (fixed order)
    const size_t pbase = ((size_t)batch * splits) * N;
    if (use_cluster && splits > 1) {
        // The K-splits of one panel form a thread-block cluster (cluster dims (1, splits, 1)).  Every CTA PUSHES its partial
        // into slot `split` of the leader's (rank 0) shared memory through distributed shared memory, one cluster barrier makes
        // the pushes visible, and the leader sums the slots in rank order from its OWN shared memory and writes y.  Round 1 let
        // the leader pull the partials instead, which needed a second barrier to keep the peers' shared memory alive and put
        // a DSMEM read latency on the leader's critical path (tools/trace_gemv.py: 1.6 us from "partial ready" to "exit" of a
        // ~13 us launch).  This replaces partial store + __threadfence + atomic ticket + reload (three serialized global round
        // trips, ~3 us) of the non-cluster path.
        float* recv = red + WARPS * MROWS * PANEL;  // leader: [8 slots][MROWS * PANEL] floats
        // the start-of-kernel arrive (below the PDL wait) pairs with this wait: every CTA of the cluster is known to be
        // running before its shared memory is written remotely (already complete by now: no cost)
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        uint32_t leader;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader) : "r"(smem_u32(recv)), "r"(0));
        leader += 4u * (uint32_t)(split * MROWS * PANEL);
        for (int c = tid; c < MROWS * PANEL; c += WARPS * 32) {
            const int m = c / PANEL, cc = c % PANEL;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) s += red[((w * MROWS + m) * COLS + cc % COLS) * LPR + cc / COLS];
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(leader + 4u * c), "f"(s) : "memory");
        }
        gemv_trace(trace, 5);  // CTA partial pushed, entering the cluster barrier
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        gemv_trace(trace, 6);  // all partials of the panel have landed at the leader
        if (split == 0) {
            for (int c = tid; c < MROWS * PANEL; c += WARPS * 32) {
                const int m = c / PANEL, gc = pcol0 + c % PANEL;
                float s = 0.f;
                for (int r = 0; r < splits; ++r) s += recv[r * MROWS * PANEL + c];  // fixed order: deterministic
                if (c % PANEL < pcols) store_y(y + (size_t)m * N, gc, s * out_scale, peers);
            }
        }
        gemv_trace(trace, 7);  // exit
        if (split == 0) peer_signal_and_wait(peers);
        return;
    }
    for (int c = tid; c < MROWS * PANEL; c += WARPS * 32) {
        const int m = c / PANEL, cc = c % PANEL;
        const int gc = pcol0 + cc;
        if (cc >= pcols) continue;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += red[((w * MROWS + m) * COLS + cc % COLS) * LPR + cc / COLS];
        if (splits == 1)
            store_y(y + (size_t)m * N, gc, s * out_scale, peers);
        else
            partial[pbase + (size_t)split * N + gc] = s;  // ticket path: MROWS == 1 only (the host never selects it otherwise)
    }
    if (splits == 1) {
        peer_signal_and_wait(peers);
        return;
    }

    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&tickets[batch * gridDim.x + panel], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)splits - 1) return;
    __threadfence();
    for (int c = tid; c < PANEL; c += WARPS * 32) {
        const int gc = pcol0 + c;
        if (c >= pcols) continue;
        // fixed summation order, but the L2 loads are issued 8 at a time (a one-at-a-time loop costs splits x ~0.4 us)
        float s = 0.f;
        for (int sp0 = 0; sp0 < splits; sp0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (sp0 + u < splits) ? __ldcg(&partial[pbase + (size_t)(sp0 + u) * N + gc]) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        store_y(y, gc, s * out_scale, peers);
    }
    if (tid == 0) tickets[batch * gridDim.x + panel] = 0u;  // ready for the next launch
    peer_signal_and_wait(peers);
}

}  // namespace b200mm
