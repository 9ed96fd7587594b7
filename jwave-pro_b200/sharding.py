"""Partition of a batch of independent signals over ranks / devices.

Every signal is independent in all three transforms, so the multi-GPU path is a partition by signal into contiguous
blocks -- no data-path collective.  The only communication is the control plane of a measurement: a barrier and a
max-reduction of the per-rank device times (bench.py), done with torch.distributed (NCCL on GPUs, gloo in the CPU
tests)."""


def shard_signals(total, world, rank):
    """Contiguous block [start, start + count) of `total` signals owned by `rank` of `world` (same rule as the C
    layer's host fan-out in jwc_api.cu: start = total * rank / world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    start = total * rank // world
    end = total * (rank + 1) // world
    return start, end - start


def reduce_max(values, device=None):
    """Max over ranks of a list of floats (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def reduce_sum(values, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
