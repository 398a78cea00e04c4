// FP32-accurate SGEMM on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma
// (kind::tf32) -> TMEM accumulators -> tcgen05.ld epilogue, with the 3xTF32 split
//
//     x = hi + lo,  hi = trunc_tf32(x) -- applied by the tensor core itself to the raw operand --, lo = tf32(x - hi)
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        (the lo*lo term is below fp32 resolution)
//
// This is north_star's main SGEMM kernel; it replaces the register-tiled WGSL shaders
// (shaders/gemm/gemm_5.wgsl:15-86 and the orphan bram/gemm3 kernels) for C = A*B with
// A (M x K), B (K x N), C (M x N) row-major f32 (src/harness.rs:17-28 fixes the layout).
//
// Two kernels per GEMM:
//   1. split_lo_kernel     elementwise pass, HBM-bound: reads B and the FIRST row bands of A once, writes B_lo and that part of
//                          A_lo (kind::tf32 ignores the low 13 mantissa bits of its 32-bit operands: the raw operand is
//                          consumed as hi, lo must be materialised).  The later row bands of A are split inside kernel 2.
//   2. sgemm_tc3x_kernel   persistent, warp-specialised, 384 threads, one CTA per SM, BK = 16:
//        warp 0    TMA producer: per k-block loads A / A_lo (128 x BK, K-major, SWIZZLE_128B or _64B) and
//                  B / B_lo (BK x BN, N-major: B is K x N row-major, so it is consumed as an MN-major
//                  operand straight from its natural layout -- no transpose anywhere; a 3-D tensor map
//                  (n%32, k, n/32) with SWIZZLE_128B_ATOM_32B lands the canonical MN-major atoms)
//        warp 1    MMA issuer: one elected thread issues 3 x (BK/8) tcgen05.mma per k-block into a
//                  fp32 accumulator in TMEM; tcgen05.commit releases the smem stage
//        warp 2    TMEM allocator, then SPLITTER: computes A_lo of the later 2048-row bands of A while earlier bands are being
//                  multiplied (per-band ready counters, the producer waits on them; cooperative launch)
//        warp 3    replicator (multi-GPU fused all-gather only): streams finished tiles to the peers over NVLink
//        warps 4-11 epilogue (two warpgroups, one per column half): after every 256 k tcgen05.ld the finished
//                  chain from TMEM and fold it into fp32 register accumulators with round-to-nearest adds (the
//                  tensor core's own accumulation truncates); at the end of the tile transpose through
//                  swizzled shared memory and store 128-byte row segments to C
//                  With Tc3xCfg::SPLIT the same warps also derive the lo tiles from the landed hi tiles in shared memory
//                  (split_try: a cooperative job done wherever they would otherwise wait) -- the default for skinny M
//      Two chain accumulators ping-pong in TMEM (2 x BN columns), so folding chain i overlaps the MMAs of
//      chain i+1, across tile boundaries as well.  Tiles are scheduled in full waves plus a stream-K tail
//      (Tc3xArgs) and rasterised in bands of 2048 rows.
//      Two instantiations: 128 x 256 tiles on single CTAs, and (Tc3xCfg::CTA2, the default for big GEMMs) 256 x 256 tiles on CTA
//      PAIRS -- a cluster of two CTAs on one TPC, tcgen05 cta_group::2: each CTA stages its own 128 rows of A and half of the
//      B tile, the pair's leader issues the MMAs for both tensor cores, each CTA drains its own accumulator and hands its C chunks
//      to the TMA (cp.async.bulk.tensor store from SWIZZLE_128B staging).
//
// Roofline: tensor pipe.  Algorithmic work 2*M*N*K flop; the tensor pipe executes 3x that in TF32.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sgemm_simt.cuh"  // PeerStore

namespace b200mm {

// ------------------------------------------------------------------------------------------------
// operand split
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// The tensor core TRUNCATES the low 13 mantissa bits of a kind::tf32 operand (measured,
// tools/probe_tc.py), so the raw fp32 operand already acts as hi = trunc(x) and only lo = tf32(x - trunc(x))
// has to be materialised (x - trunc(x) is exact in fp32).  One launch covers A and B: reads 2 x 64 MB and
// writes 2 x 64 MB at 4096^3 (a version that also materialised hi wrote 4 x 64 MB).
// The rounding of lo to tf32 is left to the consumer as well: adding half a tf32 ulp (bit 12) to the magnitude makes the tensor core's
// truncation of the low 13 bits a round-to-nearest, ties away -- the value cvt.rna.tf32.f32 would produce -- in 3 instructions per
// element instead of 5 (|x - hi| <= 2^-11 |x| cannot overflow; this code also runs inside the GEMM, Tc3xCfg::SPLIT).
__device__ __forceinline__ float split_lo1(float x) {
#ifdef B200MM_SPLIT_RNA_ADD
    return __uint_as_float(__float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)) + 0x1000u);
#else
    return to_tf32_rna(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));  // (the default until the add form is validated on the GPU)
#endif
}
__device__ __forceinline__ float4 split_lo4(const float4& v) { return make_float4(split_lo1(v.x), split_lo1(v.y), split_lo1(v.z), split_lo1(v.w)); }

// Only the first `a4` float4 of A are split here: the rows the first wave of tiles needs.  The remaining row bands of A are
// split INSIDE sgemm_tc3x_kernel by an otherwise idle warp per CTA while earlier bands are being multiplied (Tc3xArgs::split).
__global__ void split_lo_kernel(const float4* __restrict__ a, float4* __restrict__ a_lo, size_t a4,
                                const float4* __restrict__ b, float4* __restrict__ b_lo, size_t b4) {
    // programmatic dependent launch: the GEMM that follows may become resident and run its prologue (tensor-map fetch, barrier
    // init, TMEM allocation) while this pass is still streaming; it touches nothing of ours before its griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < a4 + b4; i += stride) {
        const bool is_a = i < a4;
        const float4 v = __ldg(is_a ? a + i : b + (i - a4));
        const float4 l = split_lo4(v);
        if (is_a)
            a_lo[i] = l;
        else
            b_lo[i - a4] = l;
    }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
namespace ptx {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {  // never suspends (try_wait may, for a while)
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// cta_group::2 forms (CTA pair on one TPC): the TMA lands in the issuing CTA's shared memory but signals the mbarrier of the
// pair's leader (bit 24 of a shared::cluster address is the CTA rank inside the pair; clearing it addresses the leader).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs: 128 rows each] * B[smem of both CTAs: N/2 columns each]; issued by the leader only
__device__ __forceinline__ void mma_tf32_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// arrives on the mbarrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the leader's copy of a barrier from the peer CTA
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(0));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMA store of one shared-memory box to global memory (bulk async-group completion); out-of-range parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // at most N groups still READING their shared-memory source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {  // at most N groups not yet complete (writes visible)
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate; one thread issues for the CTA.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// ------------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (tcgen05): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), base_offset [49,52)=0, layout [61,64).
//   layout 2 = SWIZZLE_128B          16-byte chunks XOR (row % 8); 8-row atoms.  Used for the K-major A tiles.
//   layout 1 = SWIZZLE_128B_BASE32B  32-byte chunks XOR (row % 4); 4-row atoms.  The ONLY layout the tensor core
//                                    accepts for MN-major 32-bit (tf32) operands -- with layout 2 the MMA silently
//                                    produces zeros (measured, tools/probe_tc.py).  TMA side: SWIZZLE_128B_ATOM_32B.
//   layout 4 = SWIZZLE_64B           16-byte chunks XOR ((row / 2) % 4); 8-row x 64-byte atoms (K-major A tiles when BK = 16).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32B = 1, kLayoutSw64 = 4;
__host__ __device__ constexpr uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(layout & 7u) << 61);
}
// Instruction descriptor, kind::tf32: D=f32 [4,6)=1, A fmt [7,10)=2 (tf32), B fmt [10,13)=2,
// A major bit15 (0 = K-major), B major bit16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Tc3xArgs {
    float* C;
    int M, N, K, ldc;
    int tiles_m, tiles_n;
    // Hybrid schedule.  Phase 1 (data-parallel): full_waves rounds of whole tiles, tile = wave * gridDim.x + blockIdx.x,
    // so the CTAs of a wave walk K in lock-step and share operand panels in L2.  Phase 2 (stream-K) covers the
    // remaining tiles (num_tiles % gridDim.x, or all of them when there are fewer tiles than CTAs): their
    // (tile, chain) units, in tile-major order, are cut into gridDim.x equal contiguous ranges.  A CTA whose range
    // ENDS inside a tile ("contributor": its last segment has c1 != chains_per_tile) parks its partial accumulators in
    // partial[blockIdx.x] and raises flags[blockIdx.x] = epoch; the CTA that holds the LAST chain of the tile
    // ("finisher") adds the contributions of the preceding CTAs in CTA order (deterministic) and stores C.  A CTA
    // therefore only ever waits on LOWER-numbered CTAs, which the hardware dispatches first: the wait cannot deadlock
    // even when the grid is not fully co-resident (another kernel holding SMs), and CTAs with an empty unit range
    // (sk_units < gridDim.x) are never waited on.  4096^3: 512 tiles on 148 SMs = 3 waves + 68 tiles split 148 ways
    // instead of a 4th wave that leaves 80 SMs idle.
    int chains_per_tile;
    int full_waves;
    long long sk_units;     // stream-K units = remaining tiles x chains_per_tile
    float4* partial;        // [gridDim.x][BN/8][256] float4
    unsigned int* flags;    // [gridDim.x]
    unsigned int epoch;     // bumped by the host on every launch, so flags never need clearing
    PeerStore peers;
    // In-kernel operand split of A (hides most of the split_lo pass behind the MMAs; at 16384^3 x 8 GPUs the replicated A
    // made that pass the Amdahl term of the step).  A is cut into bands of kTc3xBandRows rows = the bands of the tile
    // rasterisation.  Bands [0, prebands) were split by split_lo_kernel before this launch; warp 2 of every CTA splits its
    // 1/gridDim.x share of bands prebands, prebands + 1, ... in order and then adds 1 to band_cnt[band]; the TMA producer of
    // a tile in band b >= prebands first waits until band_cnt[b] == epoch * gridDim.x.  Needs all CTAs co-resident
    // (cooperative launch).  prebands >= number of bands: nothing to do in the kernel.
    const float* A;
    float* A_lo;
    unsigned int* band_cnt;  // [bands], monotonic across launches
    int prebands;
};

constexpr int kTc3xGroupM = 16;                     // tile rows per rasterisation band
constexpr int kTc3xBandRows = kTc3xGroupM * 128;    // rows of A per band

struct SegIter {  // identical iteration in the producer, issuer and epilogue roles (and on the host: b200mm_tc3x_schedule_cover)
    long long u, u1;
    int cpt, wave, full_waves, sk_tile0, block, grid;
    __host__ __device__ SegIter(int chains_per_tile, int full_waves_, long long sk_units, int block_, int grid_)
        : cpt(chains_per_tile), wave(0), full_waves(full_waves_), block(block_), grid(grid_) {
        sk_tile0 = full_waves * grid;
        u = (long long)block * sk_units / grid;
        u1 = (long long)(block + 1) * sk_units / grid;
    }
    // unit = CTA (1-CTA tiles) or CTA pair (2-CTA tiles): both CTAs of a pair walk the same schedule
    __host__ __device__ SegIter(const Tc3xArgs& p, int unit, int units) : SegIter(p.chains_per_tile, p.full_waves, p.sk_units, unit, units) {}
    // c1 != cpt => contributor segment (parked); c1 == cpt && c0 != 0 => finisher of a split tile; sk_tile = tile index inside phase 2
    __host__ __device__ bool next(int& tile, int& c0, int& c1, int& sk_tile) {
        if (wave < full_waves) {
            tile = wave * grid + block;
            c0 = 0;
            c1 = cpt;
            sk_tile = -1;
            ++wave;
            return true;
        }
        if (u >= u1) return false;
        // Phase 2 is walked from the END of the range backwards, segment by segment (chains inside a segment stay ascending):
        // the segment a CTA has to PARK (the head of a tile that continues in the next CTA) comes first, and the segment it
        // FINISHES (the tail of a tile begun by preceding CTAs, where it waits for their parts) comes last.  With the forward
        // order every finisher would wait right away for a part its predecessor only parks at the very end of its range -- the
        // CTAs would run one after the other.
        const long long last = u1 - 1;
        sk_tile = (int)(last / cpt);
        tile = sk_tile0 + sk_tile;
        const long long tile_start = (long long)sk_tile * cpt;
        const long long seg_start = u > tile_start ? u : tile_start;
        c0 = (int)(seg_start - tile_start);
        c1 = (int)(last - tile_start) + 1;
        u1 = seg_start;
        return true;
    }
};

// Stream-K fix-up protocol, shared by the kernel and the host-side replay (b200mm_tc3x_schedule_replay): the finisher
// of phase-2 tile `sk_tile` (CTA `block`, whose segment starts at chain c0 != 0) adds the parked parts of CTAs
// [first, block) -- those whose unit range reaches into the tile -- skipping CTAs whose range is empty.
__host__ __device__ inline long long tc3x_unit_start(int cta, long long sk_units, int grid) { return (long long)cta * sk_units / grid; }
__host__ __device__ inline int tc3x_first_contributor(int sk_tile, int cpt, long long sk_units, int block, int grid) {
    const long long tile_start = (long long)sk_tile * cpt;
    int j = block;
    while (j > 0 && tc3x_unit_start(j, sk_units, grid) > tile_start) --j;
    return j;
}
__host__ __device__ inline bool tc3x_cta_is_empty(int cta, long long sk_units, int grid) {
    return tc3x_unit_start(cta, sk_units, grid) == tc3x_unit_start(cta + 1, sk_units, grid);
}

// Host-side choice of the schedule (used by setup_tc3x and by the device-free introspection entry points).
struct Tc3xSchedule {
    int chains_per_tile, full_waves, grid, k_split;  // k_split > 0: uniform split of every tile (tiles < SMs)
    long long tiles, sk_units;
};
// `sms` = number of schedulable units: SMs for 128-row tiles, SM PAIRS for the 256-row tiles of the 2-CTA kernel (tile_m = 256);
// sc.grid counts units.
inline Tc3xSchedule tc3x_make_schedule(size_t M, size_t N, size_t K, int bn, int bk, int sms, bool pure_stream_k, int tile_m = 128) {
    Tc3xSchedule sc{};
    const size_t chain = 256 / bk, num_kb = (K + bk - 1) / bk;
    sc.chains_per_tile = (int)((num_kb + chain - 1) / chain);
    sc.tiles = (long long)((M + tile_m - 1) / tile_m) * (long long)((N + bn - 1) / bn);
    // hybrid schedule: whole-tile waves while there is >= one tile per SM, stream-K over the remainder
    long long grid = sc.tiles * sc.chains_per_tile < sms ? sc.tiles * sc.chains_per_tile : sms;
    // hi and lo parts of both operands against the 126 MB L2 (with room for C and the workspace)
    const bool fits_l2 = 8.0 * ((double)M * (double)K + (double)K * (double)N) <= 100e6;
    if (sc.tiles < sms && !pure_stream_k) {
        // Fewer tiles than SMs (skinny M, and the row panels of the pipelined host-buffer path): plain stream-K gives every CTA a
        // k-range that starts somewhere else, so CTAs that share an A row panel or a B column panel are never at the same k and
        // nothing is reused out of L2 -- ncu at 1024 x 4096 x 4096: 966 MB of DRAM reads for 176 MB of operands, DRAM-bound.
        // Instead split every tile into the same S k-slices (S | chains per tile, tiles x S <= SMs): one segment per CTA, all
        // CTAs walk k in lock-step, and CTA b and b + S read the same k-slice of neighbouring tiles at the same time.
        int S = 1;
        for (int d = 1; d <= sc.chains_per_tile; ++d)
            if (sc.chains_per_tile % d == 0 && sc.tiles * d <= sms) S = d;
        grid = sc.tiles * S;
        sc.k_split = S;
        // ... unless that leaves more than a fifth of the SMs idle (no admissible divisor: 1792^3 has 49 pair tiles of 7 chains for 74
        // SM pairs; 768 x 4096 x 4096 has 96 tiles for 148 SMs): then plain stream-K over all SMs -- idle SMs cost more than the
        // de-synchronised k positions, also beyond L2 (1792^3 68.3 -> 61.0 us, 768 x 4096 x 4096 153 -> 129 us, 1024 x 3072 x 8192
        // 284 -> 245 us).  Needs at least two chains of work per CTA to be worth the fix-ups.
        if (grid * 5 < (long long)sms * 4 && sc.tiles * sc.chains_per_tile >= 2LL * sms) {
            grid = sms;
            sc.k_split = 0;
        }
    }
    sc.grid = (int)grid;
    sc.full_waves = pure_stream_k ? 0 : (int)(sc.tiles / grid);
    // Whole waves exist to keep the CTAs in lock-step so that they share operand panels out of L2.  When the operands FIT in L2
    // that is worth nothing, and a remainder wave only costs: 2304^3 has 81 pair tiles for 74 SM pairs -- 7 tiles cut into ten
    // pieces each, every finisher adding ten parked parts one after the other.  L2-resident problems whose tile count is not a
    // multiple of the grid therefore run as plain stream-K: between one and two tile boundaries per CTA, at most two contributors per
    // tile (2304^3 127.3 -> 105 us, 2432 x 2432 x 1024 74.5 -> 65.6 us, 2560^3 150 -> 142 us).
    if (!pure_stream_k && !sc.k_split && fits_l2 && sc.tiles % grid != 0) sc.full_waves = 0;
    sc.sk_units = (sc.tiles - (long long)sc.full_waves * grid) * sc.chains_per_tile;
    return sc;
}

// CHAIN: number of k-blocks accumulated inside TMEM before the partial sum is folded into fp32 registers.
// Measured on B200 (tools/debug_tc3x.py): the tensor core adds into its fp32 accumulator with truncation, so a
// single chain over K = 4096 (1536 accumulate steps) is biased by ~1e-4 -- 10x worse than a sequential fp32 loop.
// Chains of 8 k-blocks (256 k, 96 steps) folded with round-to-nearest FADDs bring the error back to the fp32 level.
// CTA2: the tile is 256 x BN and belongs to a PAIR of CTAs on one TPC (cluster of 2, tcgen05 cta_group::2): each CTA stages its
// own 128 rows of A and HALF of the B tile, the leader issues one MMA for both tensor cores, each CTA drains its own
// 128 x BN accumulator.  Per SM and k-step that is 2 x (A 128 x BK + B BK x BN/2) instead of 2 x (A + B BK x BN): a third less
// L2 -> shared-memory traffic and a third less operand reads per MMA -- energy, on a kernel that runs into the power cap.
// TMA_STORE: the epilogue hands each 32 x 32 chunk to the TMA (cp.async.bulk.tensor store, SASS UTMASTG) from a SWIZZLE_128B staging
// buffer, double-buffered per warp, instead of eight st.global.v4 per lane: the LSU work of the store phase disappears from the warps
// that have to be back in time to drain the next chain, and ragged edges are clipped by the tensor map.
// SPLIT (1: B, 2: A and B): the lo tiles are not loaded but COMPUTED in shared memory: the epilogue warps, idle while a chain is being
// multiplied, turn every landed hi tile into its lo tile (same swizzled position, so no layout knowledge is needed), and the issuer
// waits for them.  The operand is then read from HBM / L2 once, and the pre-pass over it (a read and a write of the whole matrix per
// launch) disappears; with SPLIT = 2 the GEMM is a single launch with no workspace for lo parts at all.
template <int BN_, int STAGES_, bool ONE_PASS_, int BK_ = 32, int CHAIN_K_ = 256, bool CTA2_ = false, bool TMA_STORE_ = false, int SPLIT_ = 0>
struct Tc3xCfg {
    static constexpr bool TMA_STORE = TMA_STORE_;
    static constexpr bool SPLIT_B = SPLIT_ >= 1, SPLIT_A = SPLIT_ >= 2;
    static_assert(SPLIT_ == 0 || !ONE_PASS_, "the single-pass kernel has no lo operands");
    static constexpr int EPI_BUFS = TMA_STORE_ ? 2 : 1;
    static constexpr uint32_t BAR_AREA = TMA_STORE_ ? 1024 : 512;  // barriers; the TMA staging behind it must be 1024-byte aligned
    static constexpr int BM = 128, BN = BN_, BK = BK_, STAGES = STAGES_, CHAIN = CHAIN_K_ / BK_;
    static constexpr bool CTA2 = CTA2_;
    static constexpr int TILE_M = CTA2 ? 256 : 128;   // rows of C per scheduled tile
    static constexpr int BN_CTA = CTA2 ? BN / 2 : BN;  // columns of the B tile staged by one CTA
    static_assert(!CTA2 || !ONE_PASS_, "the 2-CTA kernel is the 3xTF32 one");
    static_assert(BK == 32 || BK == 16, "A tile rows are 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B)");
    static constexpr uint32_t A_LAYOUT = BK == 32 ? kLayoutSw128 : kLayoutSw64;
    static constexpr uint32_t A_SBO = 8 * BK * 4;               // 8 rows of BK floats
    static constexpr bool ONE_PASS = ONE_PASS_;
    static constexpr int EPI_WARPS = 8;                        // two warpgroups, each owns half of the BN columns
    static constexpr int THREADS = 128 + EPI_WARPS * 32;       // warps 0-3: TMA / MMA / TMEM alloc / idle
    static constexpr int COLS_PER_WG = BN / 2;
    static constexpr uint32_t A_BYTES = BM * BK * 4;           // 128 rows x BK floats, K-major
    static constexpr uint32_t B_BYTES = BK * BN_CTA * 4;       // BN_CTA/32 atoms x (BK k-rows x 128 B), SWIZZLE_128B_BASE32B, MN-major
    static constexpr uint32_t STAGE_BYTES = (ONE_PASS ? 1 : 2) * (A_BYTES + B_BYTES);
    static constexpr uint32_t TMEM_COLS = 2 * BN;              // two chain accumulators (ping-pong)
    static constexpr uint32_t EPI_STAGE_LD = 32;                                   // floats per staged row; 16 B chunks XOR-swizzled by row
    static constexpr uint32_t EPI_STAGE_BYTES = EPI_WARPS * EPI_BUFS * 32 * EPI_STAGE_LD * 4;  // 32 x 32 chunks per epilogue warp
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_AREA + EPI_STAGE_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM cols: power of 2");
    static_assert(COLS_PER_WG % 32 == 0, "epilogue reads 32 columns per tcgen05.ld");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
sgemm_tc3x_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ Tc3xArgs p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, CHAIN = Cfg::CHAIN;
    constexpr bool ONE_PASS = Cfg::ONE_PASS, CTA2 = Cfg::CTA2, SPLIT_B = Cfg::SPLIT_B, SPLIT_A = Cfg::SPLIT_A;
    constexpr int TILE_M = Cfg::TILE_M, BN_CTA = Cfg::BN_CTA;
    constexpr int GROUP_M = kTc3xBandRows / TILE_M;  // tile rows per rasterisation band (= per split band of A)
    constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    // scheduling unit: this CTA, or the CTA pair it belongs to (2-CTA tiles); cta_rank = position inside the pair
    uint32_t cta_rank = 0;
    if constexpr (CTA2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const int unit = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int units = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms are 1024 B aligned
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    // tiles-stored counters, one per epilogue warp: epilogue warps -> replicator warp (fused multi-GPU all-gather)
    auto tiles_done_ptr = [&]() { return reinterpret_cast<volatile unsigned int*>(smem_raw + (tmem_slot + 8 - smem_u32(smem_raw))); };
    // SPLIT: bfull = this CTA's B tile (SPLIT_A: and A tile) has landed (TMA -> epilogue warps); split = every epilogue warp (of both
    // CTAs of a pair) has written its share of the lo tile(s) (epilogue warps -> issuer)
    auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 9 + s); };
    auto split_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 9 + s); };
    static_assert(8u * (4 * STAGES + 9) <= Cfg::BAR_AREA, "barrier area");
    auto sA_hi = [&](int s) { return smem_base + s * STAGE_BYTES; };
    auto sA_lo = [&](int s) { return smem_base + s * STAGE_BYTES + A_BYTES; };
    auto sB_hi = [&](int s) { return smem_base + s * STAGE_BYTES + (ONE_PASS ? 1 : 2) * A_BYTES; };
    auto sB_lo = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_BYTES + B_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmAh);
        ptx::prefetch_tmap(&tmBh);
        if (!ONE_PASS) {
            if (!SPLIT_A) ptx::prefetch_tmap(&tmAl);
            if (!SPLIT_B) ptx::prefetch_tmap(&tmBl);
        }
        if (Cfg::TMA_STORE) ptx::prefetch_tmap(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        if constexpr (SPLIT_B) {
            for (int s = 0; s < STAGES; ++s) {
                ptx::mbar_init(bfull_bar(s), 1);
                ptx::mbar_init(split_bar(s), (CTA2 ? 2 : 1) * Cfg::EPI_WARPS);
            }
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), (CTA2 ? 2 : 1) * Cfg::EPI_WARPS);  // one arrive per epilogue warp (of both CTAs of a pair)
        }
        ptx::fence_barrier_init();
        for (int i = 0; i < Cfg::EPI_WARPS; ++i) tiles_done_ptr()[i] = 0;
    }
    if (warp == 2) {
        if constexpr (CTA2) {
            ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);  // one warp of EACH CTA of the pair, same slot offset
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();  // barrier inits and the TMEM slot are visible CTA-wide (also what compute-sanitizer racecheck understands)
    if constexpr (CTA2) ptx::cluster_sync_all();  // ... and the peer's, before anything is signalled remotely
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // Everything above is independent of earlier work on the stream; from here on the roles read the operands (and A_lo / B_lo from
    // the split pass that precedes this launch) and write C and the workspace.  No-op unless launched with programmatic
    // stream serialization (the host does that for the non-cooperative launches).
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tile order: grouped rasterisation.  Tiles are walked m-fastest inside bands of GROUP_M tile rows, band after
    // band, so the ~148 tiles in flight form a 16 x ~9 block and share 16 A panels + ~9 B panels in L2.  (Plain
    // m-fastest order streams ALL of A through L2 every wave once tiles_m >= 128: measured 193 vs 265 TFLOP/s
    // for 16384^3 vs 4096^3 on one GPU.)
    auto tile_coords = [&](int t, int& tm, int& tn) {
        const int band_tiles = GROUP_M * p.tiles_n;
        const int band = t / band_tiles;
        const int first_m = band * GROUP_M;
        const int rows = min(GROUP_M, p.tiles_m - first_m);
        const int r = t - band * band_tiles;
        tm = first_m + r % rows;
        tn = r / rows;
    };

    if (warp < 4) {
        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                SegIter it(p, unit, units);
                int t, c0, c1, skt;
                int ready_band = p.prebands - 1;  // bands are completed in order: one high-water mark suffices
                while (it.next(t, c0, c1, skt)) {
                    int tm, tn;
                    tile_coords(t, tm, tn);
                    if (!ONE_PASS && !SPLIT_A) {
                        const int band = tm / GROUP_M;
                        if (band > ready_band) {
                            // A_lo of this band is being produced by the splitter warps of ALL CTAs of this launch
                            const unsigned int target = p.epoch * gridDim.x;
                            unsigned int seen;
                            do {
                                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.band_cnt + band) : "memory");
                            } while ((int)(seen - target) < 0);
                            asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy writes -> TMA (async proxy) reads
                            ready_band = band;
                        }
                    }
                    const int kb_end = min(c1 * CHAIN, num_kb);
                    for (int kb = c0 * CHAIN; kb < kb_end; ++kb) {
                        ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                        if constexpr (SPLIT_A) {
                            // both hi tiles go to this CTA's own barrier; its epilogue warps derive the lo tiles and tell the issuer
                            ptx::mbar_arrive_expect_tx(bfull_bar(stage), A_BYTES + B_BYTES);
                            const int arow = tm * TILE_M + (int)cta_rank * BM, bcol = tn * (BN / 32) + (int)cta_rank * (BN_CTA / 32);
                            ptx::tma_load_2d(sA_hi(stage), &tmAh, bfull_bar(stage), kb * BK, arow);
                            ptx::tma_load_3d(sB_hi(stage), &tmBh, bfull_bar(stage), 0, kb * BK, bcol);
                        } else if constexpr (CTA2 && SPLIT_B) {
                            // the B tile goes to this CTA's own barrier (its epilogue warps derive B_lo from it), A and A_lo to the leader's
                            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * 2 * A_BYTES);
                            ptx::mbar_arrive_expect_tx(bfull_bar(stage), B_BYTES);
                            const int arow = tm * TILE_M + (int)cta_rank * BM, bcol = tn * (BN / 32) + (int)cta_rank * (BN_CTA / 32);
                            ptx::tma_load_3d(sB_hi(stage), &tmBh, bfull_bar(stage), 0, kb * BK, bcol);
                            ptx::tma_load_2d_pair(sA_hi(stage), &tmAh, full_bar(stage), kb * BK, arow);
                            ptx::tma_load_2d_pair(sA_lo(stage), &tmAl, full_bar(stage), kb * BK, arow);
                        } else if constexpr (SPLIT_B) {
                            ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * A_BYTES);
                            ptx::mbar_arrive_expect_tx(bfull_bar(stage), B_BYTES);
                            ptx::tma_load_3d(sB_hi(stage), &tmBh, bfull_bar(stage), 0, kb * BK, tn * (BN / 32));
                            ptx::tma_load_2d(sA_hi(stage), &tmAh, full_bar(stage), kb * BK, tm * BM);
                            ptx::tma_load_2d(sA_lo(stage), &tmAl, full_bar(stage), kb * BK, tm * BM);
                        } else if constexpr (CTA2) {
                            // both CTAs load their own 128 rows of A and their half of the B tile into their OWN shared memory;
                            // all transaction bytes are counted on the leader's barrier, which the leader's MMA issuer waits on
                            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                            const int arow = tm * TILE_M + (int)cta_rank * BM, bcol = tn * (BN / 32) + (int)cta_rank * (BN_CTA / 32);
                            ptx::tma_load_2d_pair(sA_hi(stage), &tmAh, full_bar(stage), kb * BK, arow);
                            ptx::tma_load_3d_pair(sB_hi(stage), &tmBh, full_bar(stage), 0, kb * BK, bcol);
                            ptx::tma_load_2d_pair(sA_lo(stage), &tmAl, full_bar(stage), kb * BK, arow);
                            ptx::tma_load_3d_pair(sB_lo(stage), &tmBl, full_bar(stage), 0, kb * BK, bcol);
                        } else {
                            ptx::mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
                            ptx::tma_load_2d(sA_hi(stage), &tmAh, full_bar(stage), kb * BK, tm * BM);
                            ptx::tma_load_3d(sB_hi(stage), &tmBh, full_bar(stage), 0, kb * BK, tn * (BN / 32));
                            if (!ONE_PASS) {
                                ptx::tma_load_2d(sA_lo(stage), &tmAl, full_bar(stage), kb * BK, tm * BM);
                                ptx::tma_load_3d(sB_lo(stage), &tmBl, full_bar(stage), 0, kb * BK, tn * (BN / 32));
                            }
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0 && cta_rank == 0) {  // 2-CTA tiles: the leader issues for both tensor cores
                constexpr uint32_t idesc = make_idesc_tf32(TILE_M, BN, /*A MN-major*/ false, /*B MN-major*/ true);
                int stage = 0;
                uint32_t phase = 0;
                uint32_t chain = 0;  // running chain index: TMEM buffer = chain & 1
                SegIter it(p, unit, units);
                int t, c0, c1, skt;
                while (it.next(t, c0, c1, skt)) {
                    for (int kb0 = c0 * CHAIN; kb0 < min(c1 * CHAIN, num_kb); kb0 += CHAIN, ++chain) {
                        const uint32_t as = chain & 1, aphase = (chain >> 1) & 1;
                        ptx::mbar_wait(tempty_bar(as), aphase ^ 1);  // epilogue has drained this accumulator
                        ptx::tc_fence_after();
                        const uint32_t d_tmem = tmem_base + as * BN;
                        const int kb1 = min(kb0 + CHAIN, num_kb);
                        for (int kb = kb0; kb < kb1; ++kb) {
                            if constexpr (!SPLIT_A) ptx::mbar_wait(full_bar(stage), phase);
                            if constexpr (SPLIT_B) ptx::mbar_wait(split_bar(stage), phase);  // hi landed AND lo written
                            ptx::tc_fence_after();
#pragma unroll
                            for (int j = 0; j < BK / 8; ++j) {
                                // A (K-major, SWIZZLE_128B): 8 tf32 = 32 B along the swizzled row; SBO = 8 rows x 128 B
                                const uint64_t a_hi = make_smem_desc(sA_hi(stage) + j * 32, 16, Cfg::A_SBO, Cfg::A_LAYOUT);
                                // B (MN-major, SWIZZLE_128B_BASE32B): 8 k-rows = two 4-row atoms (SBO = 512 B) per MMA;
                                // LBO = stride between 32-column atoms = BK rows x 128 B
                                const uint64_t b_hi = make_smem_desc(sB_hi(stage) + j * 1024, BK * 128, 512, kLayoutSw128Base32B);
                                const uint32_t acc0 = (kb > kb0 || j > 0) ? 1u : 0u;  // first MMA of a chain overwrites
                                if constexpr (CTA2) {
                                    // descriptors name the leader's shared memory; the peer's operands sit at the same offsets
                                    const uint64_t a_lo = make_smem_desc(sA_lo(stage) + j * 32, 16, Cfg::A_SBO, Cfg::A_LAYOUT);
                                    const uint64_t b_lo = make_smem_desc(sB_lo(stage) + j * 1024, BK * 128, 512, kLayoutSw128Base32B);
                                    ptx::mma_tf32_ss_pair(d_tmem, a_lo, b_hi, idesc, acc0);  // small terms first
                                    ptx::mma_tf32_ss_pair(d_tmem, a_hi, b_lo, idesc, 1u);
                                    ptx::mma_tf32_ss_pair(d_tmem, a_hi, b_hi, idesc, 1u);
                                } else if (!ONE_PASS) {
                                    const uint64_t a_lo = make_smem_desc(sA_lo(stage) + j * 32, 16, Cfg::A_SBO, Cfg::A_LAYOUT);
                                    const uint64_t b_lo = make_smem_desc(sB_lo(stage) + j * 1024, BK * 128, 512, kLayoutSw128Base32B);
                                    ptx::mma_tf32_ss(d_tmem, a_lo, b_hi, idesc, acc0);  // small terms first
                                    ptx::mma_tf32_ss(d_tmem, a_hi, b_lo, idesc, 1u);
                                    ptx::mma_tf32_ss(d_tmem, a_hi, b_hi, idesc, 1u);
                                } else {
                                    ptx::mma_tf32_ss(d_tmem, a_hi, b_hi, idesc, acc0);
                                }
                            }
                            // smem stage reusable once these MMAs retire (2-CTA: in both CTAs)
                            if constexpr (CTA2) ptx::mma_commit_pair(empty_bar(stage)); else ptx::mma_commit(empty_bar(stage));
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        // chain complete -> epilogue (of both CTAs)
                        if constexpr (CTA2) ptx::mma_commit_pair(tfull_bar(as)); else ptx::mma_commit(tfull_bar(as));
                    }
                }
            }
        } else if (warp == 2) {
            // ===================== splitter: A_lo of the later row bands =====================
            if (!ONE_PASS && !SPLIT_A) {
                constexpr int U = 16;  // independent 16-byte loads in flight per lane
                const int bands = (p.M + kTc3xBandRows - 1) / kTc3xBandRows;
                for (int band = p.prebands; band < bands; ++band) {
                    const size_t row0 = (size_t)band * kTc3xBandRows;
                    const size_t rows = min((size_t)kTc3xBandRows, (size_t)p.M - row0);
                    const size_t n4 = rows * (size_t)p.K / 4;  // K % 4 == 0
                    const float4* src = reinterpret_cast<const float4*>(p.A + row0 * p.K);
                    float4* dst = reinterpret_cast<float4*>(p.A_lo + row0 * p.K);
                    for (size_t base = (size_t)blockIdx.x * (32 * U); base < n4; base += (size_t)gridDim.x * (32 * U)) {
                        float4 v[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const size_t i = base + (size_t)u * 32 + lane;
                            if (i < n4) v[u] = __ldcs(src + i);  // streaming: read once here (the GEMM re-reads A through TMA)
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const size_t i = base + (size_t)u * 32 + lane;
                            if (i < n4) dst[i] = split_lo4(v[u]);
                        }
                    }
                    __threadfence();
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.band_cnt + band) : "memory");
                }
            }
        } else if (warp == 3 && p.peers.world > 1) {
            // ===================== replicator (fused all-gather) =====================
            // The epilogue warps store a finished tile only to the LOCAL copy of C and bump tiles_done; this warp then
            // streams the tile (L2-hot) to the same position of C on every peer over NVLink.  The NVLink back-pressure
            // therefore never reaches the warps that drain TMEM, and the MMA pipeline keeps running (measured at 8 GPUs:
            // stores issued by the epilogue warps themselves cost 16 % of the kernel).
            SegIter it(p, unit, units);
            int t, c0, c1, skt;
            unsigned int stored = 0;
            const float* src_base = p.peers.c[p.peers.rank];
            volatile unsigned int* tiles_done = tiles_done_ptr();
            while (it.next(t, c0, c1, skt)) {
                if (c1 != p.chains_per_tile) continue;  // contributor segments do not store C
                ++stored;
                // every epilogue warp counts its own stored tiles: the tile is complete when ALL eight have reached `stored`
                // (one aggregate counter would let seven fast warps that are already a tile ahead stand in for a slow one)
                while (!__all_sync(0xffffffffu, tiles_done[lane & 7] >= stored)) __nanosleep(200);
                __threadfence();
                int tm, tn;
                tile_coords(t, tm, tn);
                const int trow0 = tm * TILE_M + (int)cta_rank * BM;  // this CTA's 128 rows of the tile
                const int rows = max(0, min(BM, p.M - trow0));
                const int cols4 = min(BN, p.N - tn * BN) / 4;  // float4 per tile row
                const size_t off0 = (size_t)trow0 * p.peers.ldc + p.peers.col0 + (size_t)tn * BN;
                for (int r = 0; r < rows; r += 8) {
                    // 8 rows x up to 64 float4 per row are read ONCE (16 independent 16-byte loads in flight per lane) and
                    // fanned out from registers to every peer
                    float4 v[8][2];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c4 = lane + 32 * h;
                            if (r + rr < rows && c4 < cols4)
                                v[rr][h] = __ldcg(reinterpret_cast<const float4*>(src_base + off0 + (size_t)(r + rr) * p.peers.ldc) + c4);
                        }
#pragma unroll 1
                    for (int d = 0; d < p.peers.world; ++d) {
                        if (d == p.peers.rank) continue;
                        float* dst_base = p.peers.c[d];
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int c4 = lane + 32 * h;
                                if (r + rr < rows && c4 < cols4)
                                    *(reinterpret_cast<float4*>(dst_base + off0 + (size_t)(r + rr) * p.peers.ldc) + c4) = v[rr][h];
                            }
                    }
                }
            }
        }
    } else {
        // ===================== epilogue: 2 warpgroups x 4 warps =====================
        // warp % 4 selects the TMEM lane quadrant (hardware restriction), the warpgroup selects the column half.
        constexpr int COLS = Cfg::COLS_PER_WG;
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint32_t chain = 0;
        const int etid = threadIdx.x - 128;  // 0..255 inside the epilogue group
        // ---- in-kernel B split (SPLIT_B): a second, cooperative job of these warps.  Every k-step of this CTA, in ring order, warp w
        // turns its eighth of the landed B tile into the same eighth of the B_lo tile and arrives on split_bar.  split_try() does at
        // most one k-step and never blocks; it is called wherever these warps would otherwise wait (for a chain, for a parked
        // part) and between the chunks of the C store, so the splits stay ahead of the issuer by up to STAGES - 1 k-steps.
        [[maybe_unused]] int sp_left = 0, sp_stage = 0;
        [[maybe_unused]] uint32_t sp_phase = 0;
        if constexpr (SPLIT_B) {
            SegIter it2(p, unit, units);
            int t2, a0, a1, s2;
            while (it2.next(t2, a0, a1, s2)) sp_left += min(a1 * CHAIN, num_kb) - a0 * CHAIN;
        }
        auto split_try = [&]() -> bool {
            if constexpr (SPLIT_B) {
                if (sp_left == 0) return false;
                if (!__all_sync(0xffffffffu, ptx::mbar_test_wait(bfull_bar(sp_stage), sp_phase))) return false;
                constexpr uint32_t PER_WARP = B_BYTES / Cfg::EPI_WARPS, IT = PER_WARP / 512;
                static_assert(PER_WARP % 512 == 0, "32 lanes x 16 bytes per sweep");
                const uint32_t off = (uint32_t)(warp - 4) * PER_WARP + (uint32_t)lane * 16;
                constexpr uint32_t PER_WARP_A = A_BYTES / Cfg::EPI_WARPS, IT_A = SPLIT_A ? PER_WARP_A / 512 : 0;
                static_assert(PER_WARP_A % 512 == 0, "32 lanes x 16 bytes per sweep");
                const uint32_t off_a = (uint32_t)(warp - 4) * PER_WARP_A + (uint32_t)lane * 16;
                float4 v[IT], va[IT_A ? IT_A : 1];
#pragma unroll
                for (uint32_t i = 0; i < IT; ++i) v[i] = ptx::lds128(sB_hi(sp_stage) + off + i * 512);
                if constexpr (SPLIT_A) {
#pragma unroll
                    for (uint32_t i = 0; i < IT_A; ++i) va[i] = ptx::lds128(sA_hi(sp_stage) + off_a + i * 512);
                }
#pragma unroll
                for (uint32_t i = 0; i < IT; ++i) ptx::sts128(sB_lo(sp_stage) + off + i * 512, split_lo4(v[i]));
                if constexpr (SPLIT_A) {
#pragma unroll
                    for (uint32_t i = 0; i < IT_A; ++i) ptx::sts128(sA_lo(sp_stage) + off_a + i * 512, split_lo4(va[i]));
                }
                ptx::fence_proxy_async();  // generic-proxy writes -> tensor-core (async proxy) reads
                __syncwarp();
                if (lane == 0) {
                    // default (.release.cta) arrives: a cluster-scope release costs ~1 us per arrive (measured: 4096^3 1085 -> 653 us).
                    // The data sits in this SM's shared memory, which has a single point of coherence; the proxy fence above is
                    // what the tensor core's reads need.
                    if (CTA2 && cta_rank != 0)
                        ptx::mbar_arrive_leader(split_bar(sp_stage));
                    else
                        ptx::mbar_arrive(split_bar(sp_stage));
                }
                --sp_left;
                if (++sp_stage == STAGES) {
                    sp_stage = 0;
                    sp_phase ^= 1;
                }
                return true;
            }
            return false;
        };
        SegIter it(p, unit, units);
        int t, c0, c1, skt;
        while (it.next(t, c0, c1, skt)) {
            int tm, tn;
            tile_coords(t, tm, tn);
            float acc[COLS];
#pragma unroll
            for (int i = 0; i < COLS; ++i) acc[i] = 0.f;
            for (int kb0 = c0 * CHAIN; kb0 < min(c1 * CHAIN, num_kb); kb0 += CHAIN, ++chain) {
                const uint32_t as = chain & 1, aphase = (chain >> 1) & 1;
                if constexpr (SPLIT_B) {
                    while (!__all_sync(0xffffffffu, ptx::mbar_test_wait(tfull_bar(as), aphase)))
                        if (!split_try()) __nanosleep(32);
                } else {
                    ptx::mbar_wait(tfull_bar(as), aphase);
                }
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + as * BN + half * COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int c = 0; c < COLS / 16; ++c) {
                    float v[16];
                    ptx::tmem_ld_32x32b_x16(taddr + c * 16, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c * 16 + j] = __fadd_rn(acc[c * 16 + j], v[j]);  // round-to-nearest fold
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CTA2 && cta_rank != 0)
                        ptx::mbar_arrive_leader(tempty_bar(as));  // the leader's issuer waits for the epilogues of both CTAs
                    else
                        ptx::mbar_arrive(tempty_bar(as));
                }
            }
            if (c1 != p.chains_per_tile) {
                // contributor: this range ends inside the tile -> park the partial sums for the finisher (a higher-numbered CTA)
                float4* slot = p.partial + (size_t)blockIdx.x * (COLS / 4) * 256;
#pragma unroll
                for (int j = 0; j < COLS / 4; ++j) {
                    slot[j * 256 + etid] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    if (j % 8 == 7) split_try();  // the next segment's MMAs are already running
                }
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");  // all 8 epilogue warps have published their part
                if (etid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.flags + blockIdx.x), "r"(p.epoch) : "memory");
                continue;
            }
            if (c0 != 0) {
                // finisher of a tile that began in preceding CTAs: add their parts in CTA order.  Only lower-numbered,
                // non-empty CTAs are waited on (see Tc3xArgs).
                const int j0 = tc3x_first_contributor(skt, p.chains_per_tile, p.sk_units, unit, units);
                for (int j = j0; j < unit; ++j) {
                    if (tc3x_cta_is_empty(j, p.sk_units, units)) continue;
                    const int jc = CTA2 ? 2 * j + (int)cta_rank : j;  // the CTA of unit j that holds the same 128 rows
                    unsigned int seen;
                    if constexpr (SPLIT_B) {
                        for (;;) {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + jc) : "memory");
                            if (__all_sync(0xffffffffu, seen == p.epoch)) break;
                            if (!split_try()) __nanosleep(64);
                        }
                    } else {
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + jc) : "memory");
                        } while (seen != p.epoch);
                    }
                    const float4* slot = p.partial + (size_t)jc * (COLS / 4) * 256;
#pragma unroll
                    for (int jj = 0; jj < COLS / 4; ++jj) {
                        const float4 w = __ldcg(slot + jj * 256 + etid);
                        acc[4 * jj] = __fadd_rn(acc[4 * jj], w.x);
                        acc[4 * jj + 1] = __fadd_rn(acc[4 * jj + 1], w.y);
                        acc[4 * jj + 2] = __fadd_rn(acc[4 * jj + 2], w.z);
                        acc[4 * jj + 3] = __fadd_rn(acc[4 * jj + 3], w.w);
                        if (jj % 8 == 7) split_try();
                    }
                }
            }
            // tile finished: registers -> C.  Each thread owns one row x COLS columns; a direct store would write 16-byte
            // pieces of 32 different rows per instruction.  Instead every warp transposes 32 x 32 chunks through its own
            // 4 KB of shared memory (XOR-swizzled, conflict-free) so that each st.global.v4 covers 4 rows x 128 contiguous bytes -- full 128 B lines for
            // HBM and for NVLink when the tile also goes to the peers (fused all-gather).
            float* stage = reinterpret_cast<float*>(smem_raw + (bar_base + Cfg::BAR_AREA - smem_u32(smem_raw))) + (warp - 4) * Cfg::EPI_BUFS * 32 * Cfg::EPI_STAGE_LD;
            const int row0 = tm * TILE_M + (int)cta_rank * BM + q * 32;
            const int col0 = tn * BN + half * COLS;
            if constexpr (Cfg::TMA_STORE) {
                // The staging layout (16-byte chunk j of row r at position j ^ (r & 7)) IS the TMA's SWIZZLE_128B pattern, so a chunk
                // goes out as one 32 x 32 box; tmC addresses this rank's panel of the (local) C, out-of-range rows / columns are clipped.
#pragma unroll  // must stay unrolled: acc[] is indexed with c and has to live in registers
                for (int c = 0; c < COLS / 32; ++c) {
                    float* buf = stage + (c & 1) * 32 * Cfg::EPI_STAGE_LD;
                    if (lane == 0) ptx::bulk_wait_read<1>();  // the store that used this buffer two chunks ago has read it
                    __syncwarp();
                    split_try();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(buf + lane * Cfg::EPI_STAGE_LD + 4 * (j ^ (lane & 7))) =
                            make_float4(acc[c * 32 + 4 * j], acc[c * 32 + 4 * j + 1], acc[c * 32 + 4 * j + 2], acc[c * 32 + 4 * j + 3]);
                    ptx::fence_proxy_async();  // generic-proxy writes -> async-proxy (TMA) reads
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tmC, smem_u32(buf), col0 + c * 32, row0);
                        ptx::bulk_commit();
                    }
                }
                if (p.peers.world > 1) {
                    if (lane == 0) ptx::bulk_wait<0>();  // the tile is in (local) global memory before the replicator is told
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) {
                        volatile unsigned int* td = tiles_done_ptr() + (warp - 4);
                        *td = *td + 1u;  // single writer per counter
                    }
                }
                continue;
            }
            float* const base = p.peers.world == 0 ? p.C : p.peers.c[p.peers.rank];  // fused mode: local copy only, see replicator
            const size_t ld = p.peers.world == 0 ? (size_t)p.ldc : p.peers.ldc;
            const size_t coff = p.peers.world == 0 ? 0 : p.peers.col0;
#pragma unroll  // must stay unrolled: acc[] is indexed with c and has to live in registers
            for (int c = 0; c < COLS / 32; ++c) {
                __syncwarp();
                split_try();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + lane * Cfg::EPI_STAGE_LD + 4 * (j ^ (lane & 7))) =
                        make_float4(acc[c * 32 + 4 * j], acc[c * 32 + 4 * j + 1], acc[c * 32 + 4 * j + 2], acc[c * 32 + 4 * j + 3]);
                __syncwarp();
                const int cc = col0 + c * 32 + (lane & 7) * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = 4 * i + (lane >> 3);
                    const float4 v = *reinterpret_cast<const float4*>(stage + rr * Cfg::EPI_STAGE_LD + 4 * ((lane & 7) ^ (rr & 7)));
                    const int r = row0 + rr;
                    if (r < p.M && cc < p.N) *reinterpret_cast<float4*>(base + (size_t)r * ld + coff + cc) = v;
                }
            }
            if (p.peers.world > 1) {
                __threadfence();  // the tile is in (local) global memory before the replicator is told
                __syncwarp();
                if (lane == 0) {
                    volatile unsigned int* td = tiles_done_ptr() + (warp - 4);
                    *td = *td + 1u;  // single writer per counter
                }
            }
        }
        if constexpr (SPLIT_B) {
            while (sp_left)
                if (!split_try()) __nanosleep(32);  // (every chain this CTA waited for needed all its splits: nothing is left here)
        }
    }

    if constexpr (Cfg::TMA_STORE) {
        if (warp >= 4 && lane == 0) ptx::bulk_wait_read<0>();  // shared memory must outlive the TMA stores that read it
    }
    ptx::tc_fence_before();
    if constexpr (CTA2)
        ptx::cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or remote arrives are in flight
    else
        __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        if constexpr (CTA2) ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace b200mm
