"""Host-side helpers for running the environment on several GPUs of one node.

Batched instances are independent (the reference's only parallelism is the batch dimension
of one tensor, carle/env.py:46-48), so they shard embarrassingly: rank ``r`` of ``G`` owns a
contiguous range of instances and steps it with its own ``CARLE`` — no collective on the
data path.  The only cross-rank traffic is measurement plumbing (a max over ranks of the
CUDA-event time) and, for the single giant grid, the halo exchange in ``bigrid.py``.
Everything here is backend-agnostic ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests)."""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous, balanced ``[start, stop)`` of ``total`` instances owned by ``rank``."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu", group=None):
    """Max of a python float over all ranks (the time that bounds a multi-GPU step)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, device="cpu", group=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
