"""World / rank helpers (reference vidgen/utils/comm.py:21-79): thin wrappers over torch.distributed."""
import torch.distributed as dist


def _ready():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if _ready() else 1


def get_rank():
    return dist.get_rank() if _ready() else 0


def is_main_process():
    return get_rank() == 0


def synchronize():
    if _ready() and dist.get_world_size() > 1:
        dist.barrier()


def all_reduce_sum_(t):
    """In-place sum over ranks (NCCL for CUDA tensors, gloo for CPU tensors); no-op for a single process."""
    if _ready() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t
