"""Multi-GPU parity check (run under torchrun, one rank per GPU; uses the oracle, hence lives in tests/):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py

For both gather modes (fused peer stores / NCCL all-gather + interleave) and both SGEMM kernels it checks that
every rank ends up with the full row-major C = A*B: sampled rows against an FP64 GEMM (<= 5e-6 relative) and the
reference gate (<= 1e-3 abs vs mm_ref)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import oracle
    import wgpu_mm_b200 as w
    from wgpu_mm_b200 import shard

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = w.Context(local)
    size = int(os.environ.get("CHECK_SIZE", "2048"))
    M = N = K = size
    A = oracle.generate_weight_data(101, M, K)
    B = oracle.generate_weight_data(102, K, N)
    rows = np.array(sorted({0, 1, 127, 128, M // 2 + 3, M - 1}))
    ref64 = oracle.mm_f64_rows(A, B, rows)
    ref32 = oracle.mm_ref(A[rows], B)
    failures = 0
    for mode in ("fused", "nccl"):
        for kid in (w.KernelId.SGEMM_TC3X, w.KernelId.SGEMM_SIMT):
            plan = shard.ShardPlan(N, world, rank)
            job = shard.ShardedSgemm(ctx, M, N, K, plan, mode=mode, kernel_id=kid, seed=100)
            job.C.write(np.full(M * N, 123.25, dtype=np.float32))
            job.barrier()
            job.step()
            job.barrier()
            got = np.stack([job.C.read(np.float32, count=N, offset=int(r) * N * 4) for r in rows])
            e, m = oracle.err_vs_f64(got, ref64)
            mae = oracle.max_abs_err(got, ref32)
            unwritten = int((got == 123.25).sum())
            ok = (e / m <= 5e-6) and (mae <= 1e-3) and unwritten == 0
            print(f"rank {rank}/{world} mode={mode} kernel={w.KernelId(kid).name}: rel_f64={e / m:.3e} max_abs_vs_mm_ref={mae:.3e} "
                  f"unwritten={unwritten} {'OK' if ok else 'FAIL'}", flush=True)
            failures += 0 if ok else 1
            job.close()
    # ---- the 2-CTA pair kernel in fused mode: a per-rank panel with >= 74 pair tiles (10 x 8), whatever CHECK_SIZE is ----
    Mp, Kp, Np_total = 2560, 512, 2048 * world
    Ap = oracle.generate_weight_data(101, Mp, Kp)
    Bp_full = oracle.generate_weight_data(102, Kp, Np_total)
    prow = np.array(sorted({0, 127, 128, 255, 256, Mp // 2 + 3, Mp - 1}))
    pref64 = oracle.mm_f64_rows(Ap, Bp_full, prow)
    plan = shard.ShardPlan(Np_total, world, rank)
    job = shard.ShardedSgemm(ctx, Mp, Np_total, Kp, plan, mode="fused", kernel_id=w.KernelId.SGEMM_TC3X, seed=100)
    grid, _ = job.kern.geometry()
    job.C.write(np.full(Mp * Np_total, 123.25, dtype=np.float32))
    job.barrier()
    for _ in range(2):
        job.step()
    job.barrier()
    got = np.stack([job.C.read(np.float32, count=Np_total, offset=int(r) * Np_total * 4) for r in prow])
    e, m = oracle.err_vs_f64(got, pref64)
    unwritten = int((job.read_rows(range(0, Mp, 97)) == 123.25).sum())
    ok = (e / m <= 5e-6) and unwritten == 0 and grid[0] % 2 == 0
    print(f"rank {rank}/{world} mode=fused kernel=SGEMM_TC3X pair tiles (grid {grid[0]}): rel_f64={e / m:.3e} unwritten={unwritten} {'OK' if ok else 'FAIL'}", flush=True)
    failures += 0 if ok else 1
    job.close()
    # ---- skinny M in fused mode: each rank's 128 x 4096 x 4096 panel takes the 128 x 128-tile kernel that derives B_lo in shared
    # memory (epilogue warps split while the replicator warp streams finished tiles to the peers) ----
    Ms, Ks, Ns_total = 128, 4096, 4096 * world
    plan = shard.ShardPlan(Ns_total, world, rank)
    job = shard.ShardedSgemm(ctx, Ms, Ns_total, Ks, plan, mode="fused", kernel_id=w.KernelId.SGEMM_TC3X, seed=110)
    job.C.write(np.full(Ms * Ns_total, 123.25, dtype=np.float32))
    job.barrier()
    for _ in range(2):
        job.step()
    job.barrier()
    got = job.read_rows(range(Ms))
    unwritten = int((got == 123.25).sum())
    # own column panel against FP64 from the operands as they sit in this rank's HBM; the other panels through the cross-rank checksum
    A_dev = job.A.read(np.float32).reshape(Ms, Ks)
    B_dev = job.Bp.read(np.float32).reshape(Ks, plan.cols)
    own = got[:, plan.col0:plan.col0 + plan.cols]
    e, m = oracle.err_vs_f64(own, oracle.mm_f64(A_dev, B_dev))
    cs = torch.from_numpy(np.ascontiguousarray(got.astype(np.float64).sum(axis=0))).cuda()
    ref_cs = cs.clone()
    dist.broadcast(ref_cs, src=0)  # all ranks must hold the same full C: compare the column sums with rank 0's
    same = bool(torch.equal(cs, ref_cs))
    ok = (e / m <= 5e-6) and unwritten == 0 and same
    print(f"rank {rank}/{world} mode=fused kernel=SGEMM_TC3X skinny M=128 (grid {job.kern.geometry()[0][0]}): rel_f64={e / m:.3e} unwritten={unwritten} "
          f"same_on_all_ranks={same} {'OK' if ok else 'FAIL'}", flush=True)
    failures += 0 if ok else 1
    job.close()
    # ---- N-sharded GEMV (fp32 and sint8 in the quant.rs format) ----
    Kv, Nv = 2048, 4096
    x = oracle.generate_weight_data(301, 1, Kv)
    Wf = oracle.generate_weight_data(302, Kv, Nv)
    words, _ = oracle.sint8_quantize(Wf, Kv, Nv)
    Wq = words.reshape(Kv, Nv // 4)
    for mode in ("fused", "nccl"):
        for quant in (False, True):
            plan = shard.ShardPlan(Nv, world, rank)
            if quant:
                panel = np.ascontiguousarray(Wq[:, plan.col0 // 4:(plan.col0 + plan.cols) // 4])
                want = oracle.qgemv_ref(x, words, 1, Nv, Kv, 2.0)
                f64 = oracle.qgemv_f64(x, words, 1, Nv, Kv, 2.0)
            else:
                panel = np.ascontiguousarray(Wf[:, plan.col0:plan.col0 + plan.cols])
                want = oracle.mm_ref(x, Wf)
                f64 = oracle.mm_f64(x, Wf)
            job = shard.ShardedGemv(ctx, Kv, Nv, plan, quant=quant, mode=mode, x_host=x, panel_host=panel)
            for _ in range(3):  # several steps: epochs advance, y ping-pongs (fused mode completes across ranks in-kernel)
                job.step()
            got = job.result().reshape(1, Nv)
            e, m = oracle.err_vs_f64(got, f64)
            mae = oracle.max_abs_err(got, want)
            ok = (e / m <= 5e-6) and (mae <= 1e-3)
            print(f"rank {rank}/{world} gemv mode={mode} quant={quant}: rel_f64={e / m:.3e} max_abs={mae:.3e} {'OK' if ok else 'FAIL'}", flush=True)
            failures += 0 if ok else 1
            job.close()
    # ---- N-sharded sint8 GEMV with per-group scales (group_k = 128): every rank's panel carries its own columns' scales ----
    from wgpu_mm_b200.quant import sint8_quantize_grouped, split_grouped
    G = 128
    plan = shard.ShardPlan(Nv, world, rank)
    packed = sint8_quantize_grouped(Wf, Kv, Nv, G)
    gwords, gscales = split_grouped(packed, Kv, Nv, G)
    f64 = oracle.qgemv_grouped_f64(x, gwords, gscales, 1, Nv, Kv, G)
    wpanel = np.ascontiguousarray(gwords.reshape(Kv, Nv // 4)[:, plan.col0 // 4:(plan.col0 + plan.cols) // 4])
    spanel = np.ascontiguousarray(gscales.reshape(-1, Nv)[:, plan.col0:plan.col0 + plan.cols])
    panel = np.concatenate([wpanel.reshape(-1).view(np.uint32), spanel.reshape(-1).view(np.uint32)])
    job = shard.ShardedGemv(ctx, Kv, Nv, plan, quant=True, mode="fused", x_host=x, panel_host=panel, group_k=G)
    for _ in range(3):
        job.step()
    got = job.result().reshape(1, Nv)
    e, m = oracle.err_vs_f64(got, f64)
    ok = e / m <= 5e-6
    print(f"rank {rank}/{world} gemv mode=fused quant=grouped({G}): rel_f64={e / m:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    failures += 0 if ok else 1
    job.close()
    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    ctx.close()
    dist.destroy_process_group()
    if int(t.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("multi_gpu_check: all ranks OK", flush=True)


if __name__ == "__main__":
    main()
