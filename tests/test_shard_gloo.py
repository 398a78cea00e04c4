"""Host-side logic of the N>1 path on CPU: world_size-2 `gloo` processes exercise ShardPlan (column
partition, panel stream offsets) and the all-gather layout + interleave that b200mm_unshard_columns
implements on the device.  No GPU, no libb200mm compute calls."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, N, K, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from wgpu_mm_b200.shard import ShardPlan

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = ShardPlan(N, world, rank)
    # every rank regenerates the replicated A and ITS column panel of B from the shared counter-based stream
    A = oracle.generate_weight_data(101, M, K)
    Bp = np.stack([oracle.generate_weight_data(102, 1, plan.cols, offset=k * N + plan.col0).reshape(-1) for k in range(K)])
    Cp = oracle.mm_ref(A, Bp)  # (M, N/world) panel, CPU checker standing in for the panel GEMM
    gathered = [torch.empty(M * plan.cols) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(Cp.reshape(-1).copy()))
    C = plan.gathered_to_row_major(torch.cat(gathered).numpy(), M)
    q.put((rank, C, plan.col0, plan.cols))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_n_shard_allgather_layout_gloo(oracle, world):
    import torch.multiprocessing as mp
    M, N, K = 24, 64, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, N, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A = oracle.generate_weight_data(101, M, K)
    B = oracle.generate_weight_data(102, K, N)
    want = oracle.mm_ref(A, B)
    for rank, C, col0, cols in results:
        assert (col0, cols) == (rank * N // world, N // world)
        assert np.array_equal(C, want), f"rank {rank}: gathered C differs from the unsharded product"


def test_shard_plan_validation():
    from wgpu_mm_b200.shard import ShardPlan
    p = ShardPlan(16384, 8, 3)
    assert (p.cols, p.col0, p.panel_of(7)) == (2048, 6144, (14336, 16384))
    with pytest.raises(ValueError):
        ShardPlan(100, 8, 0)
    with pytest.raises(ValueError):
        ShardPlan(64, 2, 2)
    # interleave is the inverse of cutting column panels
    M, N, g = 5, 16, 4
    full = np.arange(M * N, dtype=np.float32).reshape(M, N)
    gathered = np.concatenate([full[:, r * 4:(r + 1) * 4].reshape(-1) for r in range(g)])
    assert np.array_equal(ShardPlan(N, g, 0).gathered_to_row_major(gathered, M), full)


def _e2e_worker(rank, world, port, M, N, K, q):
    """The host-side data movement of ShardedSgemm.e2e (round 2) with numpy / gloo standing in for the device: every rank holds
    only its 1/world ROW slice of A and its COLUMN panel of B, the A slices are all-gathered IN PLACE into the full A (the rank's
    slice already sits at its offset), and each rank produces only its own column panel of C."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from wgpu_mm_b200.shard import ShardPlan

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = ShardPlan(N, world, rank)
    Ms = M // world
    A_full = torch.full((M * K,), float("nan"))
    mine = A_full[rank * Ms * K:(rank + 1) * Ms * K]
    mine.copy_(torch.from_numpy(oracle.generate_weight_data(101, Ms, K, offset=rank * Ms * K).reshape(-1)))  # "H2D" of the row slice
    dist.all_gather_into_tensor(A_full, mine)  # in place, exactly as e2e() calls it
    Bp = np.stack([oracle.generate_weight_data(102, 1, plan.cols, offset=k * N + plan.col0).reshape(-1) for k in range(K)])
    Cp = oracle.mm_ref(A_full.numpy().reshape(M, K), Bp)
    q.put((rank, A_full.numpy().copy(), Cp, plan.col0))
    dist.barrier()
    dist.destroy_process_group()


def test_e2e_row_slices_gather_in_place_gloo(oracle):
    """Every byte crosses the host link once: rank r contributes rows [r*M/g, (r+1)*M/g) of A and gets back only its panel of C;
    the union of the panels is the full product."""
    import torch.multiprocessing as mp
    world, M, N, K = 2, 16, 32, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_e2e_worker, args=(r, world, port, M, N, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A = oracle.generate_weight_data(101, M, K)
    B = oracle.generate_weight_data(102, K, N)
    want = oracle.mm_ref(A, B)
    C = np.empty((M, N), dtype=np.float32)
    for rank, A_full, Cp, col0 in results:
        assert np.array_equal(A_full.reshape(M, K), A), f"rank {rank}: gathered A differs"
        C[:, col0:col0 + Cp.shape[1]] = Cp
    assert np.array_equal(C, want)
