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


# --------------------------------------------------------------------------- one process, several GPUs
class ShardedSymbolic:
    """One symbolic handle per device (the analysis is deterministic, so it is simply repeated:
    the schedules of a handle live on ONE device).  The analogue of placing klujax's operands with
    `jax.sharding.NamedSharding(mesh, P('lhs'))`: the n_lhs axis is cut into contiguous blocks, the
    symbolic object is replicated, nothing is exchanged on the data path
    (/root/reference/tests.py:472-492 is the reference's only multi-device test)."""

    def __init__(self, Ai, Aj, n_col: int, devices=None):
        import torch
        from . import api
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("klujax_b200 needs a CUDA device (no CPU fallback)")
        self.devices = [int(d) for d in devices]
        self.n_col = int(n_col)
        self.handles = []
        for d in self.devices:
            with torch.cuda.device(d):
                self.handles.append(api.analyze(Ai, Aj, n_col))

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def analyze_sharded(Ai, Aj, n_col: int, devices=None) -> ShardedSymbolic:
    return ShardedSymbolic(Ai, Aj, n_col, devices)


def solve_with_symbol_sharded(Ai, Aj, Ax, b, symbolic: ShardedSymbolic, transpose: bool = False,
                              check: bool = True, gather: bool = True):
    """Batched fused factor + solve with the n_lhs batch sharded over the devices of `symbolic`
    (one process, one stream per device; the devices run concurrently, the host only enqueues).
    Ax [n_lhs, n_nz], b [n_lhs, n_col, n_rhs] may live on the host (pinned for overlap) or on any
    device.  Returns x on the host if `gather` (result placement = one concatenation), else the
    list of per-device blocks, still sharded."""
    import torch
    from . import api
    Ax_t, b_t = api._as_tensor(Ax), api._as_tensor(b)
    if Ax_t.ndim != 2 or b_t.ndim != 3 or Ax_t.shape[0] != b_t.shape[0]:
        raise ValueError("sharded solve takes Ax [n_lhs, n_nz] and b [n_lhs, n_col, n_rhs]")
    L, W = Ax_t.shape[0], len(symbolic.devices)
    fn = api.tsolve_with_symbol if transpose else api.solve_with_symbol
    blocks, stats = [], []
    for r, (d, h) in enumerate(zip(symbolic.devices, symbolic.handles)):
        lo, hi = shard_range(L, r, W)
        if lo == hi:
            blocks.append(None)
            stats.append(None)
            continue
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            Ax_d = Ax_t[lo:hi].to(dev, non_blocking=True)
            b_d = b_t[lo:hi].to(dev, non_blocking=True)
            blocks.append(fn(Ai, Aj, Ax_d, b_d, h, check=False))   # enqueue only: devices overlap
            stats.append(api.last_status())
    if check:  # the reference aborts on the first failing system (klujax.cpp:311-314)
        for r, (d, h) in enumerate(zip(symbolic.devices, symbolic.handles)):
            if blocks[r] is None:
                continue
            lo, hi = shard_range(L, r, W)
            with torch.cuda.device(d):
                st, gr = stats[r]
                if bool((st != 0).any()):   # re-seed / report inside this shard
                    dev = torch.device("cuda", d)
                    blocks[r] = fn(Ai, Aj, Ax_t[lo:hi].to(dev), b_t[lo:hi].to(dev), h, check=True)
    if not gather:
        return blocks
    out = torch.empty(b_t.shape, dtype=blocks[[q for q in range(W) if blocks[q] is not None][0]].dtype)
    for r, d in enumerate(symbolic.devices):
        if blocks[r] is None:
            continue
        lo, hi = shard_range(L, r, W)
        out[lo:hi].copy_(blocks[r], non_blocking=True)
    for d in symbolic.devices:
        torch.cuda.synchronize(d)
    return out
