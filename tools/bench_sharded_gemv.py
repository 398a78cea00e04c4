"""N-sharded GEMV step time under torchrun: end-of-kernel cross-rank wait vs deferred wait, fp32 and sint8, at the
BASELINE shapes and at the per-rank panel an 8-GPU run would have (emulated with N scaled down).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/bench_sharded_gemv.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import wgpu_mm_b200 as w
    from wgpu_mm_b200 import shard
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = w.Context(local)
    for quant, K, N, note in ((False, 4096, 16384, "cfg3"), (True, 4096, 14336, "cfg4"),
                              (False, 4096, 2048 * world, "cfg3 per-rank panel of an 8-GPU run"), (True, 4096, 1792 * world, "cfg4 per-rank panel of an 8-GPU run")):
        plan = shard.ShardPlan(N, world, rank)
        pbytes = K * plan.cols * (1 if quant else 4)
        nsets = int(min(32, max(2, -(-(160 << 20) // pbytes))))
        for mode, deferred in (("fused", False), ("fused", True), ("fused-nosync", False), ("nccl", False)):
            gj = shard.ShardedGemv(ctx, K, N, plan, quant=quant, mode=mode.split("-")[0], nsets=nsets, deferred=deferred)
            if mode == "fused-nosync":  # peer stores only, no cross-rank completion: NOT a valid step, isolates the cost of the stores
                gj.kern.set_peer_flags(None)
                gj.deferred = False
            for _ in range(20):
                gj.step()
            gj.finish(); gj.barrier()
            best = 1e9
            for rep in range(3):
                ctx.timer_begin()
                for _ in range(200):
                    gj.step()
                gj.finish()
                ms = ctx.timer_end()
                gj.barrier()
                best = min(best, ms / 200)
            t = torch.tensor([best], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            yl = ctx.buffer(plan.cols * 4)
            kl = ctx.kernel(w.KernelId.QGEMV_SINT8 if quant else w.KernelId.GEMV_F32, 1, plan.cols, K, w.KernelParams(absmax=2.0, batch=1, flags=int(w.Flags.AUTOTUNE)))
            for _ in range(20):
                ctx.launch(kl, gj.x, gj.Ws[0], yl)
            ctx.sync(); ctx.timer_begin()
            for i in range(200):
                ctx.launch(kl, gj.x, gj.Ws[i % nsets], yl)
            kms = ctx.timer_end() / 200
            kl.free(); yl.free()
            if rank == 0:
                print(f"{'sint8' if quant else 'fp32 '} K={K} N={N} ({note}) x{world} {mode}{' deferred' if deferred else ''}: step {float(t) * 1e3:7.2f} us   panel-only kernel {kms * 1e3:6.2f} us  geometry {gj.kern.geometry()}", flush=True)
            gj.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
