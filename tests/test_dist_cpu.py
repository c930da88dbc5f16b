"""world_size-2 gloo test of the multi-GPU host logic (batch sharding: scatter labels, gather codes)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_local, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jukebox_b200.utils.dist_sampling import scatter_rows, gather_rows, seed_per_rank, world as w
    assert w() == (rank, world)
    y_all = torch.arange(world * n_local * 5, dtype=torch.long).view(world * n_local, 5)
    mine = scatter_rows(y_all if rank == 0 else torch.zeros_like(y_all), n_local, "cpu")
    assert torch.equal(mine, y_all[rank * n_local:(rank + 1) * n_local])
    z = mine[:, :3] * 10 + rank                      # stands in for the sampled codes of this replica
    out = gather_rows(z)
    if rank == 0:
        exp = torch.cat([y_all[r * n_local:(r + 1) * n_local, :3] * 10 + r for r in range(world)])
        assert torch.equal(out, exp)
    seed_per_rank(123)
    q.put((rank, float(torch.rand(1))))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_gather_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    draws = dict(q.get(timeout=10) for _ in range(2))
    assert draws[0] != draws[1], "replicas must not share an RNG stream"


def test_single_process_passthrough():
    from jukebox_b200.utils.dist_sampling import scatter_rows, gather_rows
    x = torch.arange(12).view(4, 3)
    assert torch.equal(scatter_rows(x, 4, "cpu"), x)
    assert torch.equal(gather_rows(x), x)
