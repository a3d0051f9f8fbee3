"""Multi-GPU plumbing: how products are spread over ranks and how timings are reduced.
(torch.distributed is plumbing only; backend nccl on GPUs, gloo in the CPU tests.)"""
import torch
import torch.distributed as dist


def shard_units(total, rank, world):
    """contiguous share [lo, hi) of `total` independent units (products, rows, columns) for `rank`"""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """device-time reduction used by bench.py: the job is as slow as its slowest rank"""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
