"""Thin torch.distributed helpers with the reference's names (unidistill/utils/torch_dist.py:5-64)."""
from torch import distributed as dist


def _ready():
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if _ready() else 0


def get_world_size() -> int:
    return dist.get_world_size() if _ready() else 1


def synchronize():
    if _ready() and dist.get_world_size() > 1:
        dist.barrier()


def reduce_sum(tensor):
    if get_world_size() < 2:
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def reduce_mean(tensor):
    return reduce_sum(tensor) / float(get_world_size())


def all_gather_object(obj):
    world = get_world_size()
    if world < 2:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def is_distributed() -> bool:
    return _ready()


def is_available() -> bool:
    return dist.is_available()
