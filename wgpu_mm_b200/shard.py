"""N-sharded SGEMM / GEMV across the GPUs of one box (new work defined by north_star).

One process per GPU.  A (M x K) is replicated, B (K x N) and C are cut into `world` column panels; rank r
computes C[:, r*N/g : (r+1)*N/g] and every rank ends the step holding the full row-major C.  Two ways
to get the panels everywhere:

  fused  the SGEMM epilogue stores each finished tile straight into the final (row, column) position of
         the full C on EVERY rank through CUDA-IPC peer mappings over NVLink, tile by tile while later
         tiles are still computing -- no collective, no interleave pass;
  nccl   panel GEMM into a contiguous [M][N/g] buffer, one NCCL all-gather (layout [g][M][N/g]), then
         b200mm_unshard_columns interleaves the panels into row-major C.  This is the baseline.

torch.distributed is used for rendezvous, handle exchange, the NCCL all-gather and barriers only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ShardPlan:
    """Column partition of an N-wide matrix over `world` ranks (contiguous, equal panels)."""
    N: int
    world: int
    rank: int

    def __post_init__(self):
        if self.world <= 0 or not (0 <= self.rank < self.world):
            raise ValueError(f"bad rank/world {self.rank}/{self.world}")
        if self.N % (4 * self.world) != 0:
            raise ValueError(f"N={self.N} must be a multiple of 4*world={4 * self.world} (float4 panels)")

    @property
    def cols(self) -> int:
        return self.N // self.world

    @property
    def col0(self) -> int:
        return self.rank * self.cols

    def panel_of(self, r: int):
        return (r * self.cols, (r + 1) * self.cols)

    def gathered_to_row_major(self, gathered: np.ndarray, M: int) -> np.ndarray:
        """Host restatement of b200mm_unshard_columns: [world][M][cols] -> [M][N] (tests / documentation)."""
        g = np.asarray(gathered).reshape(self.world, M, self.cols)
        return np.ascontiguousarray(np.transpose(g, (1, 0, 2)).reshape(M, self.N))


class PeerBarrier:
    """Stream-ordered cross-rank barrier over CUDA-IPC mapped flag words (b200mm_peer_barrier): ~5 us instead of a
    collective's launch + latency.  torch.distributed is used once, to exchange the IPC handles."""

    def __init__(self, ctx, plan: ShardPlan):
        import torch.distributed as dist
        self.ctx, self.plan = ctx, plan
        self.flags = ctx.buffer(64)
        self.flags.write(np.zeros(16, dtype=np.uint32))
        ctx.sync()
        handles = [None] * plan.world
        dist.all_gather_object(handles, self.flags.ipc_export())
        self.imports, self.ptrs = [], []
        for r in range(plan.world):
            if r == plan.rank:
                self.ptrs.append(self.flags.ptr)
            else:
                b = ctx.ipc_import(handles[r], 64)
                self.imports.append(b)
                self.ptrs.append(b.ptr)
        dist.barrier()

    def __call__(self):
        self.ctx.peer_barrier(self.flags, self.ptrs, self.plan.rank, self.plan.world)

    def close(self):
        for b in self.imports:
            b.free()
        self.flags.free()


class ShardedSgemm:
    """One 16384^3-style job on this rank.  Requires torch.distributed (NCCL) to be initialised."""

    def __init__(self, ctx, M: int, N: int, K: int, plan: ShardPlan, mode: str = "fused", kernel_id=None, seed: int = 100, tc_bn: int = 0):
        import torch
        import torch.distributed as dist
        import wgpu_mm_b200 as w

        self.torch, self.dist, self.w = torch, dist, w
        self.ctx, self.M, self.N, self.K, self.plan, self.mode = ctx, M, N, K, plan, mode
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        ctx.set_stream(self.stream.cuda_stream)  # our kernels and NCCL's waits share one stream
        Np = plan.cols
        kid = kernel_id if kernel_id is not None else w.KernelId.SGEMM_TC3X
        # operands: A replicated (same stream positions on every rank), B panel = columns [col0, col0+Np) of the full B
        self.A = ctx.buffer(M * K * 4)
        self.A.fill_weights(seed + 1, M * K)
        self.Bp = ctx.buffer(K * Np * 4)
        self.Bp.fill_weights_2d(seed + 2, K, Np, N, plan.col0)
        self.C = ctx.buffer(M * N * 4)  # full result, cudaMalloc'd by the library so it can be IPC-exported
        self.peers = []
        if mode == "fused":
            self.kern = ctx.kernel(kid, M, Np, K, w.KernelParams(flags=int(w.Flags.PEER_STORE), tune=(tc_bn, 0, 0, 0)))
            handles = [None] * plan.world
            dist.all_gather_object(handles, self.C.ipc_export())
            ptrs = []
            for r in range(plan.world):
                if r == plan.rank:
                    ptrs.append(self.C.ptr)
                else:
                    pb = ctx.ipc_import(handles[r], M * N * 4)
                    self.peers.append(pb)
                    ptrs.append(pb.ptr)
            self.kern.set_peers(plan.rank, plan.world, ptrs, N, plan.col0)
            self.Cp_t = None
            self.pbar = PeerBarrier(ctx, plan)
        elif mode == "nccl":
            self.kern = ctx.kernel(kid, M, Np, K, w.KernelParams(tune=(tc_bn, 0, 0, 0)))
            self.Cp_t = torch.empty(M * Np, dtype=torch.float32, device="cuda")
            self.G_t = torch.empty(plan.world * M * Np, dtype=torch.float32, device="cuda")
            self.Cp = ctx.wrap(self.Cp_t.data_ptr(), M * Np * 4)
        else:
            raise ValueError(mode)
        self.kern.profile(True)
        ctx.sync()
        dist.barrier()

    def step(self, A=None):
        """One full C = A*B on all ranks: every rank holds the complete row-major C when its stream drains."""
        A = A if A is not None else self.A
        if self.mode == "fused":
            self.ctx.launch(self.kern, A, self.Bp, self.C)
            # peer-flag barrier on the same stream: when it completes every rank's tiles of this step have landed here
            self.pbar()
        else:
            self.ctx.launch(self.kern, A, self.Bp, self.Cp)
            self.dist.all_gather_into_tensor(self.G_t, self.Cp_t)
            self.ctx.unshard_columns(self.G_t.data_ptr(), self.C.ptr, self.M, self.N, self.plan.world)

    def consumed(self):
        """Call (stream-ordered) after this rank has finished READING C of the last step and before the next step() when the
        operands change between steps: in fused mode a rank that is ahead would otherwise start storing the next step's tiles
        into a peer's C while that peer still reads the previous result (write-after-read across ranks).  One peer-flag
        barrier; the NCCL mode needs nothing (the gather writes C only after every rank has entered the collective)."""
        if self.mode == "fused":
            self.pbar()

    def barrier(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()
        self.dist.barrier()

    def kernel_times(self):
        return self.kern.profile_read(256)

    def read_rows(self, rows):
        """Rows of the local copy of the full C (for the out-of-band verification in bench.py / tests)."""
        out = np.empty((len(rows), self.N), dtype=np.float32)
        for i, r in enumerate(rows):
            self.C.read_into(out[i], offset=int(r) * self.N * 4)
        return out

    def checksum(self, rows=(0, 1, 4095)):
        """Sum of a few rows of the local copy of C (consistency across ranks is checked by the caller)."""
        out = np.empty(self.N, dtype=np.float32)
        acc = 0.0
        for r in rows:
            if r < self.M:
                self.C.read_into(out, offset=r * self.N * 4)
                acc += float(out.astype(np.float64).sum())
        return acc

    def e2e(self, steps: int = 2):
        """The same step through HOST buffers without redundant PCIe traffic: per step this rank uploads only its 1/world
        row slice of A and its B panel from pinned memory, the slices of A are all-gathered over NVLink (NCCL, in place),
        the sharded GEMM runs, and the rank reads back only its own column panel of C (the union over the ranks is the
        full C in host memory, each byte crossing PCIe once).  Returns (seconds per step (max over ranks), H2D bytes,
        D2H bytes) -- bytes per rank."""
        import ctypes as C
        import time
        w, torch, dist = self.w, self.torch, self.dist
        M, N, K, Np, g, r = self.M, self.N, self.K, self.plan.cols, self.plan.world, self.plan.rank
        if M % g:
            raise ValueError("e2e: M must be a multiple of the world size")
        Ms = M // g
        hs, arrs = [], []
        for nfloat in (Ms * K, K * Np, M * Np):
            h = C.c_void_p()
            w._lib.check(w.lib().b200mm_host_alloc(nfloat * 4, C.byref(h)))
            hs.append(h)
            arrs.append(np.ctypeslib.as_array((C.c_float * nfloat).from_address(h.value)))
        hA, hB, hC = arrs
        self.A.read_into(hA, offset=r * Ms * K * 4)
        self.Bp.read_into(hB)
        A_t = torch.empty(M * K, dtype=torch.float32, device="cuda")  # gathered A (NCCL needs a torch tensor)
        A_e = self.ctx.wrap(A_t.data_ptr(), M * K * 4)
        mine = A_t[r * Ms * K:(r + 1) * Ms * K]
        times = []
        for i in range(steps + 1):
            self.barrier()
            t0 = time.perf_counter()
            A_e.write(hA, offset=r * Ms * K * 4)
            self.Bp.write(hB)
            dist.all_gather_into_tensor(A_t, mine)  # in place: rank r's slice already sits at its offset
            self.step(A=A_e)
            # blocking, ordered after the step's trailing cross-rank barrier on the same stream
            self.C.read_2d_into(hC, offset=self.plan.col0 * 4, src_pitch=N * 4, width_bytes=Np * 4, rows=M)
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(dt)
        self.e2e_panel_checksum = float(hC[:Np].astype(np.float64).sum())
        t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self.barrier()
        A_e.free()
        del A_t
        for h in hs:
            w.lib().b200mm_host_free(h)
        return float(t.item()), (Ms * K + K * Np) * 4, M * Np * 4

    def close(self):
        self.barrier()
        if getattr(self, "pbar", None):
            self.pbar.close()
        for b in self.peers:
            b.free()
        self.kern.free()
        for b in (self.A, self.Bp, self.C):
            b.free()


class ShardedGemv:
    """y[1 x N] = x[1 x K] * W[K x N] with W (fp32 or sint8 words) cut into `world` column panels (SURVEY 8e:
    "weight rows" in the LLM out x in convention are columns of the reference's K x N matrix, Q3).  x is replicated,
    every rank ends the step holding the full y.  fused: the kernel's final store writes the slice into y on every
    rank through CUDA-IPC peer mappings, and the kernel's last CTA publishes a per-step epoch to every rank and waits
    for theirs (b200mm_kernel_set_peer_flags) -- ONE launch per step, no barrier kernel, no collective.  By default the
    wait is DEFERRED to the start of the next step's launch (a decode chain: x of step i+1 derives from y of step i), so
    the NVLink flag latency hides behind that launch's ramp-up; finish() closes the chain.  y is
    double-buffered by step parity so a fast rank's next step cannot overwrite a y a slower rank still reads.
    nccl: local slice + all_gather_into_tensor (slices are contiguous)."""

    def __init__(self, ctx, K: int, N: int, plan: ShardPlan, quant: bool = False, mode: str = "fused", seed: int = 300,
                 x_host=None, panel_host=None, absmax: float = 2.0, nsets: int = 1, deferred: bool = True, autotune: bool = True, group_k: int = 0):
        import torch
        import torch.distributed as dist
        import wgpu_mm_b200 as w

        self.torch, self.dist, self.w = torch, dist, w
        self.ctx, self.K, self.N, self.plan, self.mode, self.quant = ctx, K, N, plan, mode, quant
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        ctx.set_stream(self.stream.cuda_stream)
        Np = plan.cols
        if quant and Np % 16:
            raise ValueError("sint8 panels must be a multiple of 16 columns")
        self.x = ctx.buffer_from(x_host) if x_host is not None else ctx.buffer(K * 4)
        if x_host is None:
            self.x.fill_weights(seed + 1, K)
        # nsets > 1: that many weight panels, rotated step by step, so that a benchmark loop streams from HBM instead of L2
        self.Ws = []
        for i in range(max(1, nsets)):
            if panel_host is not None:
                Wb = ctx.buffer_from(panel_host)
            elif quant:
                Wb = ctx.buffer(K * Np)
                Wb.fill_weights(seed + 2 + plan.rank + 1000 * i, K * Np // 4)  # arbitrary bytes: bandwidth measurements only
            else:
                Wb = ctx.buffer(K * Np * 4)
                Wb.fill_weights_2d(seed + 2 + 1000 * i, K, Np, N, plan.col0)
            self.Ws.append(Wb)
        self.W, self.steps_done = self.Ws[0], 0
        self.y = ctx.buffer(2 * N * 4)  # two y buffers, alternated by step parity (fused mode)
        self.peers = []
        kid = w.KernelId.QGEMV_SINT8 if quant else w.KernelId.GEMV_F32
        # the per-rank panel is a small matrix: let the library measure the launch geometry once (B200MM_F_AUTOTUNE)
        self.kern = ctx.kernel(kid, 1, Np, K, w.KernelParams(absmax=0.0 if group_k else absmax, batch=1, group_k=group_k, flags=int(w.Flags.AUTOTUNE) if autotune else 0))
        if mode == "fused":
            self.flags = ctx.buffer(64)
            self.flags.write(np.zeros(16, dtype=np.uint32))
            ctx.sync()
            handles = [None] * plan.world
            dist.all_gather_object(handles, (self.y.ipc_export(), self.flags.ipc_export()))
            ptrs, fptrs = [], []
            for r in range(plan.world):
                if r == plan.rank:
                    ptrs.append(self.y.ptr)
                    fptrs.append(self.flags.ptr)
                else:
                    pb = ctx.ipc_import(handles[r][0], 2 * N * 4)
                    fb = ctx.ipc_import(handles[r][1], 64)
                    self.peers += [pb, fb]
                    ptrs.append(pb.ptr)
                    fptrs.append(fb.ptr)
            self.kern.set_peers(plan.rank, plan.world, ptrs, N, plan.col0)
            self.deferred = deferred and plan.world > 1
            if plan.world > 1:
                self.kern.set_peer_flags(fptrs, pingpong_stride=N, deferred=deferred)
        elif mode == "nccl":
            self.ys_t = torch.empty(Np, dtype=torch.float32, device="cuda")
            self.yg_t = torch.empty(N, dtype=torch.float32, device="cuda")
            self.ys = ctx.wrap(self.ys_t.data_ptr(), Np * 4)
            self.yg = ctx.wrap(self.yg_t.data_ptr(), N * 4)
        else:
            raise ValueError(mode)
        ctx.sync()
        dist.barrier()

    def step(self):
        self.W = self.Ws[self.steps_done % len(self.Ws)]
        self.steps_done += 1
        if self.mode == "fused":
            # one launch: when it completes on the stream every rank's y slice of this step has landed here
            self.ctx.launch(self.kern, self.x, self.W, self.y)
        else:
            self.ctx.launch(self.kern, self.x, self.W, self.ys)
            self.dist.all_gather_into_tensor(self.yg_t, self.ys_t)

    def finish(self):
        """Closes a chain of steps on the stream: afterwards the full y of the last step is complete on this rank."""
        if self.mode == "fused" and getattr(self, "deferred", False):
            self.kern.peer_wait()

    def result(self) -> np.ndarray:
        self.finish()
        self.barrier()
        if self.mode != "fused":
            return self.yg.read(np.float32, count=self.N)
        parity = self.kern.peer_epoch & 1 if self.plan.world > 1 else 0
        return self.y.read(np.float32, count=self.N, offset=parity * self.N * 4)

    def barrier(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()
        self.dist.barrier()

    def bytes_per_rank(self) -> int:
        Np = self.plan.cols
        return (self.K * Np if self.quant else 4 * self.K * Np) + 4 * self.K + 4 * Np

    def close(self):
        self.barrier()
        for b in self.peers:
            b.free()
        self.kern.free()
        for b in [self.x, self.y] + self.Ws:
            b.free()
        if getattr(self, "flags", None):
            self.flags.free()
