"""Rank helpers that tolerate a missing process group (the reference's adapter,
jukebox/utils/dist_adapter.py, raises on torch >= 2 when no group exists - SURVEY.md section 8c)."""
import torch.distributed as dist


def _ready():
    return dist.is_available() and dist.is_initialized()


def get_rank():
    return dist.get_rank() if _ready() else 0


def get_world_size():
    return dist.get_world_size() if _ready() else 1


def barrier():
    if _ready():
        dist.barrier()


def broadcast(tensor, src):
    if _ready():
        dist.broadcast(tensor, src)


def all_gather(tensor_list, tensor):
    if _ready():
        dist.all_gather(tensor_list, tensor)
    else:
        tensor_list[0] = tensor


def print_once(msg):
    if get_rank() == 0:
        print(msg)
