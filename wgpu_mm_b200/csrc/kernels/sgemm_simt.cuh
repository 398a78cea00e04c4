// Warp-tiled FP32 SGEMM on the FMA pipe (north_star's "accuracy fallback and comparison point").
//
// It carries over the idea of the reference's fastest wired shader, shaders/gemm/gemm_5.wgsl:15-86
// (2-D register tiling: a block tile staged in shared memory, each thread owning a TM x TN patch and
// doing TM*TN fma per TM+TN shared loads), resized for a B200 SM instead of a 64-thread workgroup:
//
//   gemm_5.wgsl            32 x 32 x 16 block tile, 4 x 4 per thread, 64 threads, 2 barriers / k-tile
//   this kernel           128 x 128 x 16 block tile, 8 x 8 per thread, 256 threads, 1 barrier / k-tile,
//                          shared memory double-buffered, 128-bit global and shared accesses,
//                          A staged transposed so both fragments are float4 reads.
//
// Accumulation per output is k-sequential fma in fp32, i.e. exactly gemm_5.wgsl's order
// (threadResults = fma(regM, regN, threadResults), :74-76): with B200MM_F_SEQUENTIAL_K results agree with the
// oracle_wgsl_gemm_5 restatement bit for bit (always the case when there are at least 2 tiles per SM).  Small problems
// are split along K over otherwise idle SMs; the parts are added in a fixed order, so the result is deterministic but
// rounds differently from a single k-sequential chain.
//
// Data layout: A (M x K), B (K x N), C (M x N) row-major f32 in HBM; nothing is transposed or padded
// in global memory.  Shapes that are not multiples of the tile go through the guarded instantiation.
//
// Roofline: FMA pipe, 148 SM x 128 lanes x 2 flop x f_clk.  Algorithmic work 2*M*N*K flop.
#pragma once
#include "common.cuh"

namespace b200mm {

struct SimtCfg {
    static constexpr int BM = 128, BN = 128, BK = 16;
    static constexpr int THREADS = 256;
    static constexpr int TM = 8, TN = 8;
    static constexpr int AS_LD = BM + 4;  // padded: transposed stores of A hit different banks
    static constexpr int BS_LD = BN;
    static constexpr int SMEM_FLOATS = 2 * (BK * AS_LD + BK * BS_LD);
};

// Optional fused all-gather: every finished row segment is also stored to the same (row, col) of the
// full C on each peer GPU (NVLink peer mappings), see b200mm_kernel_set_peers.
// Schedule.  One CTA per tile (hardware block scheduler, 2 CTAs resident per SM).  When a problem has fewer tiles
// than resident CTAs, each tile is cut into `split` K-parts (CTA = tile * split + part) so that all SMs have work.
// The LAST part owns the tile: it adds the partial tiles the lower parts parked in `partial` (flag = epoch) in part order
// (deterministic) and stores C.  The owner only waits on lower-numbered CTAs, which the hardware dispatches first, so
// the wait cannot deadlock even when the launch is not fully co-resident.
// (Measured and rejected: a persistent loop -- extra live state pushed the main loop over the 128-register budget and
// ptxas sank the prefetch; splitting only the partial last wave of a large problem -- no gain.)
struct SimtSched {
    int tiles_m, tiles_n, group_m, tile_offset, split;
    float4* partial;        // [gridDim.x][BM*BN/4]
    unsigned int* flags;    // [gridDim.x]
    unsigned int epoch;
};

struct PeerStore {
    float* c[8];
    int world;     // 0 = disabled
    int rank;      // own rank: its slot in c[] is the local full C
    size_t ldc;    // leading dimension of the full C
    size_t col0;   // first column of this rank's panel in the full C
    // In-kernel cross-rank completion (b200mm_kernel_set_peer_flags; GEMV kernels): when flags[0] != nullptr the last CTA to
    // finish publishes `epoch` into slot `rank` of every rank's flag array and waits until all `world` local slots carry
    // it, so the kernel completes on the stream only when every rank's slice of this step has landed here -- no barrier launch.
    unsigned int* flags[8];   // flags[d] = rank d's array of >= world + 1 words (own entry = local); word `world` = local CTA counter
    unsigned int epoch;       // 1, 2, 3, ... per launch
    unsigned int signal_ctas; // CTAs that store outputs (each bumps the local counter once per launch)
    unsigned int deferred;    // 0: the last CTA also WAITS for every rank's epoch before the kernel ends (kernel completion =>
                              //    all slices have landed here).  1: it only publishes; the NEXT launch of the kernel object waits
                              //    for epoch - 1 right after griddepcontrol.wait, before it touches x, so the NVLink latency hides
                              //    behind the next launch's ramp-up and weight prefetch (b200mm_kernel_peer_wait closes a chain).
};

// every rank has published `epoch` into the local flag array (threads 0 .. world-1 poll one slot each)
__device__ __forceinline__ void peer_wait_epoch(const PeerStore& peers, unsigned int epoch) {
    if ((int)threadIdx.x < peers.world) {
        const unsigned int* local = peers.flags[peers.rank];
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        unsigned int seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local + threadIdx.x) : "memory");
            if ((int)(seen - epoch) >= 0) break;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ull) __trap();  // 10 s: a peer never arrived -- fail loudly instead of hanging the GPU
        } while (true);
    }
}

// Called by every thread of a CTA that has issued its last peer stores of the launch.  Returns after the cross-rank
// wait in the (single) CTA that turned out to be the last one; immediately in all others.
__device__ __forceinline__ void peer_signal_and_wait(const PeerStore& peers) {
    if (peers.world <= 1 || peers.flags[0] == nullptr) return;
    // Ordering without a system-scope fence per thread (hundreds of concurrent MEMBAR.SYS serialise device-wide: measured
    // +15 us per launch): the CTA's stores -> bar.sync -> thread 0: gpu-scope fence + counter (release pattern); the last
    // CTA's thread 0 reads the counter + gpu-scope fence (acquire pattern) -> bar.sync -> ONE sys-scope release per peer.
    // Causality order is transitive across these scopes, so every counted store is visible before the flag.
    __syncthreads();
    __shared__ unsigned int s_last;
    unsigned int* local = peers.flags[peers.rank];
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(local + peers.world, 1u);  // monotonic across launches: never reset
        __threadfence();
        s_last = (prev + 1u == peers.epoch * peers.signal_ctas) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    if ((int)threadIdx.x < peers.world) {
        const int d = threadIdx.x;
        __threadfence_system();  // cumulativity: everything the counted CTAs stored is ordered before the flag
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.flags[d] + peers.rank), "r"(peers.epoch) : "memory");
    }
    if (!peers.deferred) peer_wait_epoch(peers, peers.epoch);
}

// Closes a chain of deferred launches: one warp waits until every rank has published `epoch`.
__global__ void peer_wait_kernel(PeerStore peers, unsigned int epoch) { peer_wait_epoch(peers, epoch); }

// Blackwell packed fp32: one FFMA2 performs two IEEE fmas (d.xy = a.xy * b.xy + c.xy).  The outer product is
// issue-bound (ncu: 78 % issue-active vs 70 % FMA-pipe-active with scalar FFMA), so halving the FMA instruction
// count matters; results are bit-identical to scalar fmaf.
__device__ __forceinline__ void ffma2(float& c0, float& c1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 rc, ra, rb;\n\tmov.b64 rc, {%0,%1};\n\tmov.b64 ra, {%2,%2};\n\tmov.b64 rb, {%3,%4};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "+f"(c0), "+f"(c1)
        : "f"(a), "f"(b0), "f"(b1));
}

template <bool GUARD>
__global__ void __launch_bounds__(SimtCfg::THREADS, 2)
sgemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                  int ldc, const __grid_constant__ PeerStore peers, const __grid_constant__ SimtSched sched) {
    using Cfg = SimtCfg;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, AS_LD = Cfg::AS_LD, BS_LD = Cfg::BS_LD;
    __shared__ __align__(16) float smem[Cfg::SMEM_FLOATS];
    float* const As0 = smem;                    // As[buf] = As0 + buf * BK * AS_LD   (stored k-major: As[k][m])
    float* const Bs0 = smem + 2 * BK * AS_LD;   // Bs[buf] = Bs0 + buf * BK * BS_LD   (Bs[k][n])

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // 8 warps as 2 (m) x 4 (n); warp tile 64 x 32; lanes as 8 (m) x 4 (n); thread patch = 2 x 2 blocks of 4 x 4
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int tm = (lane >> 2) * 4, tn = (lane & 3) * 4;
    // global -> register staging: A tile 128 x 16 = 512 float4 (2 per thread), B tile 16 x 128 = 512 float4
    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;    // rows a_row and a_row + 64
    const int b_row = tid >> 5, b_col = (tid & 31) * 4;  // rows b_row and b_row + 8
    const int KT = (K + BK - 1) / BK;

    // grouped rasterisation (bands of 16 tile rows) so the tiles in flight share operand panels in L2
    auto tile_coords = [&](int t, int& bm, int& bn) {
        const int GROUP_M = sched.group_m;
        const int band_tiles = GROUP_M * sched.tiles_n;
        const int band = t / band_tiles;
        const int first_m = band * GROUP_M;
        const int rows = min(GROUP_M, sched.tiles_m - first_m);
        const int r = t - band * band_tiles;
        bm = (first_m + r % rows) * BM;
        bn = (r / rows) * BN;
    };

    {
        const int tile = sched.tile_offset + (int)blockIdx.x / sched.split;
        const int part = (int)blockIdx.x % sched.split, nparts = sched.split;
        int block_m, block_n;
        tile_coords(tile, block_m, block_n);
        const int kt_begin = (int)((long long)part * KT / nparts), kt_end = (int)((long long)(part + 1) * KT / nparts);

        const float* Ag = A + (size_t)(block_m + a_row) * K + a_kq;
        const float* Bg = B + (size_t)b_row * N + block_n + b_col;
        float4 ra[2], rb[2];
        // A goes through registers (it is transposed on the way into shared memory); in the aligned instantiation B is
        // copied global -> shared by cp.async straight into the NEXT buffer, so the prefetch needs no registers and cannot
        // be sunk below the FMA block by the compiler (which it did, at the 128-register cap, with a register-staged B).
        auto load_tile = [&](int k0, int nbuf) {
            if constexpr (!GUARD) {
                ra[0] = __ldg(reinterpret_cast<const float4*>(Ag + k0));
                ra[1] = __ldg(reinterpret_cast<const float4*>(Ag + (size_t)64 * K + k0));
                const uint32_t dst = smem_u32(Bs0 + nbuf * (BK * BS_LD) + b_row * BS_LD + b_col);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(Bg + (size_t)k0 * N) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 8 * BS_LD * 4), "l"(Bg + (size_t)(k0 + 8) * N) : "memory");
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[4];
                    const int r = block_m + a_row + 64 * h;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = k0 + a_kq + j;
                        v[j] = (r < M && k < K) ? A[(size_t)r * K + k] : 0.f;
                    }
                    ra[h] = make_float4(v[0], v[1], v[2], v[3]);
                    const int kb = k0 + b_row + 8 * h;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = block_n + b_col + j;
                        v[j] = (kb < K && c < N) ? B[(size_t)kb * N + c] : 0.f;
                    }
                    rb[h] = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        };
        auto store_tile = [&](int buf) {
            float* as = As0 + buf * (BK * AS_LD);
            float* bs = Bs0 + buf * (BK * BS_LD);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = a_row + 64 * h;
                as[(a_kq + 0) * AS_LD + r] = ra[h].x;
                as[(a_kq + 1) * AS_LD + r] = ra[h].y;
                as[(a_kq + 2) * AS_LD + r] = ra[h].z;
                as[(a_kq + 3) * AS_LD + r] = ra[h].w;
                if constexpr (GUARD) *reinterpret_cast<float4*>(&bs[(b_row + 8 * h) * BS_LD + b_col]) = rb[h];
            }
            if constexpr (!GUARD) asm volatile("cp.async.wait_all;" ::: "memory");
        };

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        if (kt_begin < kt_end) {
            load_tile(kt_begin * BK, 0);
            store_tile(0);
        }
        __syncthreads();

        for (int kt = kt_begin; kt < kt_end; ++kt) {
            const int buf = (kt - kt_begin) & 1;
            if (kt + 1 < kt_end) load_tile((kt + 1) * BK, buf ^ 1);  // global loads in flight across the whole k-tile
            const float* as = As0 + buf * (BK * AS_LD) + wm + tm;
            const float* bs = Bs0 + buf * (BK * BS_LD) + wn + tn;
            float4 fa[2][2], fb[2][2];  // register double buffer for the fragments
            fa[0][0] = *reinterpret_cast<const float4*>(as);
            fa[0][1] = *reinterpret_cast<const float4*>(as + 32);
            fb[0][0] = *reinterpret_cast<const float4*>(bs);
            fb[0][1] = *reinterpret_cast<const float4*>(bs + 16);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const int cur = k & 1, nxt = cur ^ 1;
                if (k + 1 < BK) {
                    fa[nxt][0] = *reinterpret_cast<const float4*>(as + (k + 1) * AS_LD);
                    fa[nxt][1] = *reinterpret_cast<const float4*>(as + (k + 1) * AS_LD + 32);
                    fb[nxt][0] = *reinterpret_cast<const float4*>(bs + (k + 1) * BS_LD);
                    fb[nxt][1] = *reinterpret_cast<const float4*>(bs + (k + 1) * BS_LD + 16);
                }
                const float a[8] = {fa[cur][0].x, fa[cur][0].y, fa[cur][0].z, fa[cur][0].w,
                                    fa[cur][1].x, fa[cur][1].y, fa[cur][1].z, fa[cur][1].w};
                const float b[8] = {fb[cur][0].x, fb[cur][0].y, fb[cur][0].z, fb[cur][0].w,
                                    fb[cur][1].x, fb[cur][1].y, fb[cur][1].z, fb[cur][1].w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; j += 2) ffma2(acc[i][j], acc[i][j + 1], a[i], b[j], b[j + 1]);
            }
            if (kt + 1 < kt_end) store_tile(buf ^ 1);
            __syncthreads();
        }

        if (part != nparts - 1) {
            // K-part of a split tile: park the partial tile for the owner (thread-major float4 layout, coalesced)
            float4* slot = sched.partial + (size_t)blockIdx.x * (BM * BN / 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    slot[(i * 2 + h) * Cfg::THREADS + tid] = make_float4(acc[i][4 * h + 0], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
            __threadfence();
            __syncthreads();
            if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sched.flags + blockIdx.x), "r"(sched.epoch) : "memory");
            return;
        }
        for (int pp = 0; pp < nparts - 1; ++pp) {
            // owner (last part): add the lower parts in part order (deterministic); they have lower block indices
            const unsigned int src = blockIdx.x - (unsigned)(nparts - 1) + (unsigned)pp;
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sched.flags + src) : "memory");
            } while (seen != sched.epoch);
            const float4* slot = sched.partial + (size_t)src * (BM * BN / 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 w = __ldcg(slot + (i * 2 + h) * Cfg::THREADS + tid);
                    acc[i][4 * h + 0] += w.x;
                    acc[i][4 * h + 1] += w.y;
                    acc[i][4 * h + 2] += w.z;
                    acc[i][4 * h + 3] += w.w;
                }
        }

        // epilogue: rows {tm..tm+3, 32+tm..}, cols {tn..tn+3, 16+tn..} of the warp tile; float4 stores
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = block_m + wm + tm + (i & 3) + (i >> 2) * 32;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = block_n + wn + tn + h * 16;
                const float4 v = make_float4(acc[i][4 * h + 0], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
                if constexpr (!GUARD) {
                    if (peers.world == 0) {
                        *reinterpret_cast<float4*>(&C[(size_t)r * ldc + c]) = v;
                    } else {
#pragma unroll 1
                        for (int p = 0; p < peers.world; ++p)
                            *reinterpret_cast<float4*>(&peers.c[p][(size_t)r * peers.ldc + peers.col0 + c]) = v;
                    }
                } else {
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (r < M && c + j < N) C[(size_t)r * ldc + c + j] = e[j];
                }
            }
        }
    }
}

}  // namespace b200mm
