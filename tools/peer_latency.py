"""Measurement tool: NVLink flag round trip between two B200s, inside one kernel per rank (b200mm_debug_peer_pingpong).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/peer_latency.py

Prints ns per ROUND TRIP (rank 0 -> rank 1 -> rank 0) for: release/acquire flags, relaxed flags, and data + fence + flag
with 2 K / 16 K floats of payload.  The one-way figure (half) is the floor of any cross-rank completion signal, e.g. the
in-kernel completion of the N-sharded GEMV (b200mm_kernel_set_peer_flags)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import wgpu_mm_b200 as w
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = w.Context(local)
    flags = ctx.buffer(256)
    flags.write(np.zeros(64, dtype=np.uint32))
    pay = ctx.buffer(1 << 20)
    ctx.sync()
    hs = [None, None]
    dist.all_gather_object(hs, (flags.ipc_export(), pay.ipc_export()))
    pf = ctx.ipc_import(hs[1 - rank][0], 256)
    pp = ctx.ipc_import(hs[1 - rank][1], 1 << 20)
    lib = w.lib()
    lib.b200mm_debug_peer_pingpong.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_double)]
    base = 0
    for mode, payload, name in ((0, 0, "release/acquire flag only"), (1, 0, "relaxed flag only"), (2, 0, "fence.sys + relaxed flag"),
                                (2, 2048, "8 KB payload + fence.sys + flag"), (2, 16384, "64 KB payload + fence.sys + flag"),
                                (0, 2048, "8 KB payload + release flag (thread 0 only orders its own stores)")):
        for rep in range(2):
            dist.barrier()
            ns = C.c_double()
            iters = 2000
            w._lib.check(lib.b200mm_debug_peer_pingpong(ctx.handle, C.c_void_p(flags.ptr), C.c_void_p(pf.ptr), C.c_void_p(pp.ptr), payload, rank, iters, mode, base, C.byref(ns)), ctx.handle)
            base += iters
        if rank == 0:
            print(f"{name:70s} round trip {ns.value:8.0f} ns   one way {ns.value / 2:7.0f} ns", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
