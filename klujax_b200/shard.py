"""Sharding of the n_lhs batch over the GPUs of one box.

Systems never interact (the reference's loop body touches slice i only,
/root/reference/klujax.cpp:300-322), so the batch is cut into contiguous blocks,
one per rank, with NO collective on the data path.  The symbolic handle is
replicated: every rank runs the (deterministic) host analysis itself.  Result
placement is the caller's choice: leave x sharded, or all-gather it.
"""
from __future__ import annotations


def shard_range(n_lhs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of systems owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n_lhs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_sharded(x_local, n_lhs: int, group=None):
    """Result placement: all-gather the per-rank blocks of x (torch.distributed;
    NCCL on GPUs, gloo on CPU).  Blocks may be ragged (n_lhs % world != 0)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_lhs, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    if pad.is_complex():
        buf = [torch.zeros_like(torch.view_as_real(pad)) for _ in range(world)]
        dist.all_gather(buf, torch.view_as_real(pad).contiguous(), group=group)
        buf = [torch.view_as_complex(t) for t in buf]
    else:
        buf = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(buf, pad, group=group)
    del rank
    return torch.cat([buf[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
