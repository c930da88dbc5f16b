"""Batch-sharded sampling across the GPUs of one box.

Samples are independent, so every rank runs its own replica of the model on its slice of the batch;
the only traffic is one broadcast of the conditioning (labels `y`, optionally upper-level codes) from
rank 0 before a window and one gather of the int64 codes after it.  Nothing is exchanged inside the
token loop.  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch as t
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def scatter_rows(x_all, n_local, device, src=0):
    """rank `src` holds x_all [world*n_local, ...] (any device, may be pinned host memory); every rank
    returns its rows [n_local, ...] on `device`.  Other ranks pass a tensor of the same shape/dtype
    (contents ignored) or None together with shape/dtype of x_all via `like`."""
    rank, ws = world()
    if ws == 1:
        return x_all.to(device, non_blocking=True)
    buf = x_all.to(device, non_blocking=True) if rank == src else t.empty(x_all.shape, dtype=x_all.dtype, device=device)
    dist.broadcast(buf, src)
    assert buf.shape[0] == ws * n_local, f"expected {ws * n_local} rows, got {buf.shape[0]}"
    return buf[rank * n_local:(rank + 1) * n_local].contiguous()


def gather_rows(z, dst=0):
    """inverse of scatter_rows: rank `dst` returns the concatenation over ranks, others return z"""
    rank, ws = world()
    if ws == 1:
        return z
    out = [t.empty_like(z) for _ in range(ws)] if rank == dst else None
    dist.gather(z.contiguous(), out, dst)
    return t.cat(out, dim=0) if rank == dst else z


def seed_per_rank(seed):
    """the reference never seeds per rank (every replica would draw the same tokens under a fixed seed);
    offset the torch RNG so that replicas are independent"""
    rank, _ = world()
    t.manual_seed(seed + 7919 * rank)
    if t.cuda.is_available():
        t.cuda.manual_seed(seed + 7919 * rank)
