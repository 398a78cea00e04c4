/*
 * b200mm.h -- C ABI of the B200-native replacement for wgpu-mm's GPU hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one of the
 * five places where the reference's harness touches wgpu (all paths relative to the reference
 * crate root):
 *
 *   gpu_handle                         src/harness.rs:87-101   -> b200mm_ctx_create / _destroy
 *   create_buffer_init                 src/harness.rs:135,158  -> b200mm_buffer_create_init
 *   create_shader_module_unchecked +
 *   create_compute_pipeline            src/harness.rs:179-191  -> b200mm_kernel_get
 *   mm (encode + bind + dispatch)      src/harness.rs:250-287  -> b200mm_launch
 *   to_cpu (DownloadBuffer + poll)     src/harness.rs:289-302  -> b200mm_buffer_read
 *
 * Conventions
 *   - plain C: opaque handles, raw pointers and sizes; no C++/torch types cross this boundary.
 *   - every function returns 0 (B200MM_OK) on success, a negative b200mm_status otherwise;
 *     b200mm_last_error() gives the message.  The reference panics on every error
 *     (SURVEY 5.3); wrappers turn a non-zero status into a panic / exception.
 *   - all matrices are row-major f32 with natural leading dimension, C = A(MxK) * B(KxN),
 *     alpha = 1, beta = 0, C is OVERWRITTEN (the reference pre-fills C with noise,
 *     src/harness.rs:55).  Quantised B is K*N/4 u32 words, 4 x int8 little-endian along N
 *     (src/quant.rs:20-26).
 *   - a ctx owns one CUDA stream; launches on a ctx are asynchronous and in-order, like
 *     command buffers submitted to one wgpu queue (src/harness.rs:212-237).  A ctx is not
 *     thread-safe; any number of ctxs may coexist (one per test thread, as `cargo test` does).
 *   - there is NO CPU fallback: without a CUDA device every call that needs one fails.
 */
#ifndef B200MM_H
#define B200MM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200MM_API
#else
#define B200MM_API __attribute__((visibility("default")))
#endif

typedef enum b200mm_status {
    B200MM_OK = 0,
    B200MM_ERR_INVALID = -1,     /* bad argument / shape the kernel cannot take                 */
    B200MM_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                         */
    B200MM_ERR_NO_DEVICE = -3,   /* "No GPU found given preference" (src/harness.rs:96)         */
    B200MM_ERR_LIMITS = -4,      /* "Compute limits exceeded" (src/workload.rs:60)              */
    B200MM_ERR_UNSUPPORTED = -5, /* kernel id not built / device is not sm_100                  */
    B200MM_ERR_TOLERANCE = -6    /* "MAE too high" (src/harness.rs:83), host harness only       */
} b200mm_status;

/* Kernel ids.  1..15 are faithful CUDA ports of the WGSL shaders (same work split per invocation,
 * same accumulation order, launched with the Workload's grid/block); 32.. are the B200-native
 * kernels that the hot path actually uses. */
typedef enum b200mm_kernel_id {
    B200MM_K_GEMM_1 = 1,      /* shaders/gemm/gemm_1.wgsl   + src/gemm.rs:16-32   */
    B200MM_K_GEMM_1V = 2,     /* shaders/gemm/gemm_1v.wgsl  + src/gemm.rs:34-50   */
    B200MM_K_GEMM_2 = 3,      /* shaders/gemm/gemm_2.wgsl   + src/gemm.rs:52-67   */
    B200MM_K_GEMM_3 = 4,      /* shaders/gemm/gemm_3.wgsl   + src/gemm.rs:69-90   */
    B200MM_K_GEMM_4 = 5,      /* shaders/gemm/gemm_4.wgsl   + src/gemm.rs:92-119  */
    B200MM_K_GEMM_5 = 6,      /* shaders/gemm/gemm_5.wgsl   + src/gemm.rs:121-150 */
    B200MM_K_GEMM_WONNX = 7,  /* shaders/gemm.wgsl + shaders/gemm_macro.wgsl (orphan)  */
    B200MM_K_BRAM = 8,        /* shaders/bram.wgsl (orphan)                             */
    B200MM_K_BRAM8X8 = 9,     /* shaders/bram8x8.wgsl (orphan)                          */
    B200MM_K_GEMM3 = 10,      /* shaders/gemm3.wgsl (orphan)                            */
    B200MM_K_QGEMV_1 = 11,    /* shaders/gemv/qgemv_1.wgsl + src/gemv.rs:17-33          */

    B200MM_K_SGEMM_SIMT = 32, /* warp-tiled FP32 FMA-pipe SGEMM (gemm_5's idea, B200 sized)   */
    B200MM_K_SGEMM_TC3X = 33, /* TMA + tcgen05/TMEM SGEMM, 3xTF32 split (FP32-accurate)       */
    B200MM_K_GEMV_F32 = 34,   /* HBM-streaming fp32 GEMV; M <= 16 rows of x (1, 2, 4, 8 share one pass over W; other counts run in such chunks) */
    B200MM_K_QGEMV_SINT8 = 35 /* HBM-streaming sint8 GEMV, in-register dequant; M <= 16 (chunks of 4, 2, 1 rows)                        */
} b200mm_kernel_id;

/* Replaces what the reference bakes into the WGSL text through Tera (src/gemm.rs:24-29,
 * src/gemv.rs:26-31).  Zero-initialise, then set what applies. */
typedef struct b200mm_kernel_params {
    uint32_t workgroup_size[3]; /* blockDim for the faithful ports (Workload::size); 0 = kernel default      */
    float absmax;               /* dequant scale for the quantised kernels ("absmax", src/gemv.rs:30)        */
    uint32_t batch;             /* qgemv: number of (x, W, y) problems along global_id.y (qgemv_1.wgsl:12-14) */
    uint32_t flags;             /* B200MM_F_* below                                                            */
    uint32_t tune[4];           /* kernel-specific tuning knobs; 0 = default.  Results never depend on them beyond the summation order.
                                 * SGEMM_TC3X: [0] 128 / 256 = tile columns, 512 / 513 = force CTA pairs / single CTAs; [1] 1 = pure stream-K;
                                 *   [2] 32 = BK 32, 6 = st.global epilogue on pairs; [3] where the tf32 lo operands come from: 1 / 4 = B_lo
                                 *   computed in shared memory (default for M <= 256), 5 = A_lo and B_lo, 2 / 3 = split pre-pass (default otherwise).
                                 * GEMV_F32 / QGEMV_SINT8: [0] instantiation, [1] K-splits, [2] 1 = no programmatic dependent launch (W written
                                 *   by the preceding kernel), [3] 1 = no cluster reduction, >= 16 = number of column panels.           */
    uint32_t group_k;           /* qgemv_sint8 only: rows per quantisation group: 32, 64 or a multiple of 128; 0 = the reference's one
                                 * global absmax (src/quant.rs:17).  When > 0, B holds the K*N int8 weights followed by
                                 * ceil(K/group_k)*N f32 scales (per group and column) and `absmax` is ignored: SURVEY 8f rank 3. */
} b200mm_kernel_params;

#define B200MM_F_NONE 0u
#define B200MM_F_TC3X_1X 0x1u       /* SGEMM_TC3X: single-pass TF32 (fails the reference gate; for ncu/accuracy tables only) */
#define B200MM_F_SEQUENTIAL_K 0x4u   /* SGEMM_SIMT: never split K across CTAs: every output is one k-sequential fma chain (gemm_5.wgsl order, bit-exact vs its restatement) */
#define B200MM_F_AUTOTUNE 0x8u      /* GEMV_F32 / QGEMV_SINT8 (M == 1, tune[0] == tune[1] == 0): b200mm_kernel_get times the candidate
                                     * (geometry, K-split count) pairs once on scratch weights larger than L2 and keeps the fastest.  The
                                     * shape is baked into the kernel object as in the reference (src/gemm.rs:5-7), so this is the analogue
                                     * of picking the WGSL tile constants per shape; the result is deterministic per kernel OBJECT. */
#define B200MM_F_CONST_B 0x10u      /* SGEMM_TC3X: B is a constant operand (weights): its tf32 lo part is computed on the first launch with a given
                                     * B pointer and reused while the pointer stays the same (the per-launch split pass then covers A only).
                                     * The caller promises not to change B's contents in place; pass a different buffer (or re-create the
                                     * kernel object) when the weights change. */
#define B200MM_F_PEER_STORE 0x2u    /* SGEMM_*: epilogue also stores the C panel to the peers set by b200mm_kernel_set_peers   */

typedef struct b200mm_ctx b200mm_ctx;
typedef struct b200mm_buffer b200mm_buffer;
typedef struct b200mm_kernel b200mm_kernel;

/* ---- library -------------------------------------------------------------------------------- */
B200MM_API const char* b200mm_version(void);
/* Number of visible CUDA devices (0 if none); never fails. */
B200MM_API int b200mm_device_count(void);
/* Message of the last failure on this thread (ctx may be NULL for ctx-less calls). */
B200MM_API const char* b200mm_last_error(const b200mm_ctx* ctx);

/* ---- device bring-up: gpu_handle, src/harness.rs:87-101 -------------------------------------- */
B200MM_API int b200mm_ctx_create(int device_ordinal, b200mm_ctx** out);
B200MM_API int b200mm_ctx_destroy(b200mm_ctx* ctx);
/* Device facts the host harness prints / uses for rooflines. */
B200MM_API int b200mm_ctx_device_info(const b200mm_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                                      size_t* global_mem_bytes, char* name, size_t name_len);
/* Run this ctx on a caller-owned cudaStream_t (e.g. the torch current stream); NULL = own stream. */
B200MM_API int b200mm_ctx_set_stream(b200mm_ctx* ctx, void* cuda_stream);
B200MM_API void* b200mm_ctx_stream(const b200mm_ctx* ctx);
/* Blocks until everything submitted on the ctx has finished (device.poll(Wait), src/harness.rs:300). */
B200MM_API int b200mm_sync(b200mm_ctx* ctx);
/* Count of kernels this library has launched on the ctx since creation (for bench.py's gpu_launches). */
B200MM_API uint64_t b200mm_ctx_launch_count(const b200mm_ctx* ctx);

/* ---- buffers: create_buffer_init src/harness.rs:135,158; to_cpu :289-302 ---------------------- */
B200MM_API int b200mm_buffer_create(b200mm_ctx* ctx, size_t bytes, b200mm_buffer** out);
B200MM_API int b200mm_buffer_create_init(b200mm_ctx* ctx, const void* host, size_t bytes, b200mm_buffer** out);
/* Non-owning view of device memory allocated elsewhere (torch tensor, peer mapping). */
B200MM_API int b200mm_buffer_wrap(b200mm_ctx* ctx, void* device_ptr, size_t bytes, b200mm_buffer** out);
B200MM_API int b200mm_buffer_free(b200mm_ctx* ctx, b200mm_buffer* buf);
B200MM_API void* b200mm_buffer_device_ptr(const b200mm_buffer* buf);
B200MM_API size_t b200mm_buffer_bytes(const b200mm_buffer* buf);
/* Asynchronous host->device copy on the ctx stream (host memory should be pinned to overlap). */
B200MM_API int b200mm_buffer_write(b200mm_ctx* ctx, b200mm_buffer* buf, size_t offset, const void* host, size_t bytes);
/* Blocking device->host read of the whole range (to_cpu). */
B200MM_API int b200mm_buffer_read(b200mm_ctx* ctx, const b200mm_buffer* buf, size_t offset, void* host, size_t bytes);
/* Blocking read of a column panel: `rows` rows of width_bytes, row r taken from offset + r*src_pitch of the buffer and
 * stored at host + r*dst_pitch (an N-sharded rank reads back only its own panel of the row-major C). */
B200MM_API int b200mm_buffer_read_2d(b200mm_ctx* ctx, const b200mm_buffer* buf, size_t offset, size_t src_pitch, void* host,
                                     size_t dst_pitch, size_t width_bytes, size_t rows);
/* Pinned host memory helpers (the reference's staging buffers are wgpu-internal). */
B200MM_API int b200mm_host_alloc(size_t bytes, void** out);
B200MM_API int b200mm_host_free(void* p);
/* Device-side synthetic data, bit-identical to oracle_generate_weight_data_at (U[-10,10)/50,
 * src/harness.rs:103-121 with a seed added): fills n f32 starting at stream position `offset`. */
B200MM_API int b200mm_buffer_fill_weights(b200mm_ctx* ctx, b200mm_buffer* buf, uint64_t seed, uint64_t offset, size_t n);

/* Same stream, for a column panel of a row-major matrix with leading dimension src_ld: element (r, c) of the
 * rows x cols buffer receives stream position offset + r*src_ld + src_col0 + c (N-sharded B panels). */
B200MM_API int b200mm_buffer_fill_weights_2d(b200mm_ctx* ctx, b200mm_buffer* buf, uint64_t seed, uint64_t offset, size_t rows,
                                             size_t cols, size_t src_ld, size_t src_col0);

/* ---- kernels: shader module + pipeline, src/harness.rs:179-191 -------------------------------- */
/* Shapes are fixed per kernel object exactly as they are baked into the reference's WGSL
 * (src/gemm.rs:5-7).  May allocate device workspace (split operands, split-K partials). */
B200MM_API int b200mm_kernel_get(b200mm_ctx* ctx, int kernel_id, size_t M, size_t N, size_t K,
                                 const b200mm_kernel_params* params, b200mm_kernel** out);
B200MM_API int b200mm_kernel_free(b200mm_ctx* ctx, b200mm_kernel* kern);
B200MM_API const char* b200mm_kernel_name(int kernel_id);
/* Launch geometry the library would choose for this kernel (grid[3], block[3]). */
B200MM_API int b200mm_kernel_geometry(const b200mm_kernel* kern, uint32_t grid[3], uint32_t block[3]);
/* Device workspace held by the kernel object, in bytes. */
B200MM_API size_t b200mm_kernel_workspace_bytes(const b200mm_kernel* kern);

/* ---- launch: mm, src/harness.rs:250-287 ------------------------------------------------------- */
/* Asynchronous, in order on the ctx stream.  `grid` is the Workload's WorkgroupCount: the faithful
 * ports (ids < 32) use it as gridDim (NULL = library default); the B200-native kernels derive their
 * own launch configuration and treat it as advisory. */
B200MM_API int b200mm_launch(b200mm_ctx* ctx, b200mm_kernel* kern, const b200mm_buffer* A, const b200mm_buffer* B,
                             b200mm_buffer* C, const uint32_t grid[3]);
/* Same with raw device pointers (for callers that own their memory). */
B200MM_API int b200mm_launch_ptr(b200mm_ctx* ctx, b200mm_kernel* kern, const void* A, const void* B, void* C,
                                 const uint32_t grid[3]);
/* End-to-end call with HOST buffers: H2D of A and B, launch, D2H of C, blocking.  hostA/B/C should be
 * pinned.  dA/dB/dC are caller-provided device staging buffers of matching size.  For the SGEMM kernels
 * the copies are pipelined with the compute over row panels of A / C (same tolerances as b200mm_launch). */
B200MM_API int b200mm_mm_host(b200mm_ctx* ctx, b200mm_kernel* kern, const void* hostA, size_t bytesA, const void* hostB,
                              size_t bytesB, void* hostC, size_t bytesC, b200mm_buffer* dA, b200mm_buffer* dB,
                              b200mm_buffer* dC);

/* ---- device-free introspection of the SGEMM_TC3X work schedule (no GPU needed; used by the CPU tests) --------------
 * out = {grid, full_waves, chains_per_tile, k_split, tiles, stream_k_units} for a bn x 128 tile, bk-deep stages, `sms` SMs. */
B200MM_API int b200mm_tc3x_schedule(size_t M, size_t N, size_t K, int bn, int bk, int sms, int pure_stream_k, int out[6]);
/* What b200mm_kernel_get(SGEMM_TC3X, M, N, K, {tune, flags}) picks on a device with `sms` SMs (tune may be NULL = all defaults):
 * out = {tile columns 128 / 256, BK, 1 = CTA pairs (256 x 256 tiles), 1 = TMA-store epilogue, lo operands computed in shared
 * memory (0 none, 1 B, 2 A and B), 1 = all of A split by the pre-pass, grid in CTAs, whole-tile waves, k-slices per tile (0 = none)}.
 * N or K not a multiple of 4: the plan of the zero-padded shape. */
B200MM_API int b200mm_tc3x_plan(size_t M, size_t N, size_t K, int sms, const uint32_t tune[4], uint32_t flags, int out[9]);
/* Runs the kernel's own segment iterator for every CTA on the host and counts how often each (tile, chain) unit is
 * visited: cover[tile * chains_per_tile + chain] += 1 (caller zero-fills; cover_len >= tiles * chains_per_tile). */
B200MM_API int b200mm_tc3x_schedule_cover(size_t M, size_t N, size_t K, int bn, int bk, int sms, int pure_stream_k, uint16_t* cover,
                                          size_t cover_len, int* max_segments_per_cta, int* max_chains_per_cta);

/* Replays the stream-K fix-up protocol (who waits on whom) on the host: *violations == 0 means every finisher CTA waits only
 * on lower-numbered CTAs that do publish a part of the same tile, and the parts cover the tile exactly. */
B200MM_API int b200mm_tc3x_schedule_replay(size_t M, size_t N, size_t K, int bn, int bk, int sms, int pure_stream_k, int* violations,
                                           int* max_wait_list);

/* ---- timing: CUDA events on the ctx stream (the reference uses Instant::now, src/harness.rs:225) */
B200MM_API int b200mm_timer_begin(b200mm_ctx* ctx);
B200MM_API int b200mm_timer_end(b200mm_ctx* ctx, float* elapsed_ms); /* records, synchronises, returns ms */
/* Per-launch timing of the kernel object's DOMINANT device kernel only (e.g. the tcgen05 GEMM without the
 * operand-split pass): when enabled, every launch records a CUDA-event pair around that kernel on the ctx
 * stream (ring of 256).  profile_read synchronises and returns the durations recorded since the last
 * read / enable, oldest first. */
B200MM_API int b200mm_kernel_profile_enable(b200mm_ctx* ctx, b200mm_kernel* kern, int enable);
B200MM_API int b200mm_kernel_profile_read(b200mm_ctx* ctx, b200mm_kernel* kern, float* ms_out, int max_n, int* n_out);
/* Measurement tool: FP32 FMA-pipe ceiling of this device in TFLOP/s from a register-only microbenchmark (packed != 0: FFMA2,
 * else scalar FFMA); the measured denominator of the SIMT SGEMM's roofline. */
B200MM_API int b200mm_measure_fma_peak(b200mm_ctx* ctx, int packed, int iters, int reps, double* tflops_out);
/* Overwrites a >L2-sized scratch buffer so the next launch starts with a cold L2. */
B200MM_API int b200mm_flush_l2(b200mm_ctx* ctx);

/* ---- multi-GPU (new work): N-sharded panels, one process per GPU -------------------- */
/* Export / import a CUDA IPC handle (64 bytes) of a library-owned buffer so that another rank can
 * map it; the handles travel over the caller's control plane (torch.distributed / MPI). */
B200MM_API int b200mm_ipc_export(b200mm_ctx* ctx, const b200mm_buffer* buf, void* handle64);
B200MM_API int b200mm_ipc_import(b200mm_ctx* ctx, const void* handle64, size_t bytes, b200mm_buffer** out);
/* Tell a SGEMM kernel object where the full row-major C (M x ldc) lives on every rank and which column
 * offset this rank's panel starts at.  The epilogue then stores each finished tile to all `world`
 * destinations over NVLink instead of a later all-gather.  For the GEMV kernels peer_c[] are the full
 * y vectors (ldc is ignored) and col_offset the first output of this rank's slice. */
B200MM_API int b200mm_kernel_set_peers(b200mm_kernel* kern, int rank, int world, void* const* peer_c, size_t ldc,
                                       size_t col_offset);
/* In-kernel cross-rank completion for the N-sharded GEMV kernels (after b200mm_kernel_set_peers, world >= 2): peer_flags[r] is
 * rank r's mapping of a library-allocated, ZERO-INITIALISED flag array of >= world + 1 u32 (own entry = local).  Every launch
 * then ends with the last CTA publishing a per-launch epoch to all ranks and waiting for theirs: when the kernel completes on
 * the ctx stream, every rank's slice of this step has landed in the local y -- no barrier launch, no collective.  All ranks
 * must launch the same number of times.  pingpong_stride (floats; 0 = off): successive launches alternate between two y
 * buffers that distance apart (launch e writes y + (e & 1) * stride on every rank; e = b200mm_kernel_peer_epoch after the
 * launch), so a fast rank's step e + 1 never overwrites the y a slower rank's consumer of step e is still reading.
 * deferred != 0: a launch only PUBLISHES its epoch; the wait for all ranks moves to the start of the NEXT launch of the same
 * kernel object (after the previous grid has drained, before x is read), so the NVLink flag latency (~3 us one way, measured
 * with tools/peer_latency.py) hides behind the next launch's ramp-up and weight prefetch -- the form a decode chain wants.
 * b200mm_kernel_peer_wait then closes the chain: a one-warp kernel on the ctx stream that returns when every rank's last
 * launch has landed here.
 * peer_flags == NULL switches back to caller-side synchronisation. */
B200MM_API int b200mm_kernel_set_peer_flags(b200mm_kernel* kern, void* const* peer_flags, size_t pingpong_stride, int deferred);
B200MM_API int b200mm_kernel_peer_wait(b200mm_ctx* ctx, b200mm_kernel* kern);
B200MM_API unsigned int b200mm_kernel_peer_epoch(const b200mm_kernel* kern);
/* Stream-ordered barrier across the ranks of one box without a collective library: `local_flags` is a
 * library-allocated, zero-initialised buffer of >= world u32 on every rank, `peer_flags[r]` its mapping on
 * rank r (b200mm_ipc_import; own entry = local).  The kernel stores a per-call epoch into slot `rank` of every
 * peer's flags and spins until all `world` local slots carry it: everything enqueued on the ctx stream before
 * the call on ANY rank (e.g. the peer stores of a fused GEMM) has completed when it returns on the stream.
 * Every rank must call it the same number of times. */
B200MM_API int b200mm_peer_barrier(b200mm_ctx* ctx, b200mm_buffer* local_flags, void* const* peer_flags, int rank, int world);
/* After an NCCL all-gather of column panels (layout [world][M][N/world]) interleave them into
 * row-major C (M x N).  `gathered` and `C` are device pointers. */
B200MM_API int b200mm_unshard_columns(b200mm_ctx* ctx, const void* gathered, void* C, size_t M, size_t N, int world);

#ifdef __cplusplus
}
#endif
#endif /* B200MM_H */
