"""N > 1 path on CPU: world_size-2 gloo.  Each rank owns a contiguous block of
the n_lhs batch, solves it independently (here with the oracle standing in for
the GPU call — this test checks the sharding/placement logic, not the kernels)
and the blocks are gathered; no collective on the data path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from klujax_b200.shard import gather_sharded, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_the_batch():
    for n in (1, 2, 7, 64, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_lhs, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from klujax_b200 import synth
    from oracle import oracle as O
    pat = synth.SaxPattern(64, seed=1)
    lo, hi = shard_range(n_lhs, rank, world)
    Ax = pat.values(lo, hi - lo)  # system m always uses rng(100 + m): shards are reproducible
    b = np.ones((hi - lo, 64, 1), np.complex128)
    sym = O.analyze(pat.Ai, pat.Aj, 64)  # replicated symbolic analysis
    x = O.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym)
    O.free_symbolic(sym)
    full = gather_sharded(torch.from_numpy(x), n_lhs)
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "x.npy"), full.numpy())
        np.save(os.path.join(out_dir, "count.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_solve(tmp_path, oracle):
    from klujax_b200 import synth
    n_lhs = 9  # ragged: 5 + 4
    mp.spawn(_worker, args=(2, _free_port(), n_lhs, str(tmp_path)), nprocs=2, join=True)
    x = np.load(tmp_path / "x.npy")
    assert float(np.load(tmp_path / "count.npy")[0]) == n_lhs
    pat = synth.SaxPattern(64, seed=1)
    Ax = pat.values(0, n_lhs)
    sym = oracle.analyze(pat.Ai, pat.Aj, 64)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, np.ones((n_lhs, 64, 1), np.complex128), sym)
    oracle.free_symbolic(sym)
    assert x.shape == ref.shape and np.array_equal(x, ref)
