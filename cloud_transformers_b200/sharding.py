"""Multi-GPU plumbing for the hot path.  Splat / Slice shard over independent (batch, head) units with NO
data-path collective: every (b, h) owns its grid slab (layers/cloud_transform.py:164-178).
One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to agree on the
timing: whole-job throughput = units processed by all ranks / max over ranks of the device time.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (strong scaling over clouds)."""
    assert 0 <= rank < world
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(units_local, ms_local, device=None):
    """(total units over all ranks, max ms over ranks).  Works without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(units_local), float(ms_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()), float(t.item())
