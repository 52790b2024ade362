"""klujax_b200 — B200-native (sm_100a) twin of klujax's numeric hot path.

Batched sparse LU refactorization + triangular solves over the n_lhs batch with
one host-side symbolic analysis, behind the reference's operator interface
(klujax_b200.api mirrors /root/reference/klujax.py) and a C-ABI
(include/klujax_cuda.h).  Importing the package does not load the CUDA library;
calling any operator without it (or without a GPU) raises — no CPU fallback.
"""
__version__ = "0.1.0"

_API = ("solve", "dot", "dot_transpose_ax", "solve_with_symbol_value_and_jvp",
        "solve_with_symbol_value_and_vjp", "coalesce", "analyze", "factor", "refactor", "solve_with_symbol",
        "tsolve_with_symbol", "solve_with_numeric", "tsolve_with_numeric", "refactor_and_solve",
        "free_symbolic", "free_numeric", "KLUHandleManager", "KlujaxCudaError", "last_status")


_SHARD = ("analyze_sharded", "solve_with_symbol_sharded", "ShardedSymbolic", "shard_range", "gather_sharded")


def __getattr__(name):  # lazy: `import klujax_b200.synth` must not need torch
    if name in _API:
        from . import api
        return getattr(api, name)
    if name in _SHARD:
        from . import shard
        return getattr(shard, name)
    raise AttributeError(name)
