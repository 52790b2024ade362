"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as /root/reference/klujax.py
(`solve` :63, `dot` :99, `analyze` :297, `solve_with_symbol` :376,
`tsolve_with_symbol` :416, `factor` :451, `refactor` :479,
`solve_with_numeric` :523, `tsolve_with_numeric` :559, `refactor_and_solve` :612,
`free_symbolic` :240, `free_numeric` :267, `coalesce` :134, `KLUHandleManager`
:180), but the arrays are torch CUDA tensors and every call goes through the
C-ABI of include/klujax_cuda.h instead of an XLA custom call.  torch is only
used for device memory and streams.  There is no CPU fallback: without the CUDA
library or a GPU these functions raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Callable

import numpy as np
import torch

from . import _cabi
from ._cabi import KlujaxCudaError

__all__ = [
    "solve", "dot", "dot_transpose_ax", "solve_with_symbol_value_and_jvp",
    "solve_with_symbol_value_and_vjp", "coalesce", "analyze", "factor", "refactor", "solve_with_symbol",
    "tsolve_with_symbol", "solve_with_numeric", "tsolve_with_numeric", "refactor_and_solve",
    "free_symbolic", "free_numeric", "KLUHandleManager", "KlujaxCudaError", "last_status",
]

_last = {"status": None, "growth": None}


def last_status():
    """(status int32[n_lhs], growth float64[n_lhs]) device tensors of the most recent
    factorizing call: KLU-style per-system codes and the pivot-growth certificate."""
    return _last["status"], _last["growth"]


# --------------------------------------------------------------------------- helpers
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("klujax_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _as_tensor(a: Any) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a
    return torch.from_numpy(np.ascontiguousarray(a))


def _is_complex(*ts: torch.Tensor) -> bool:
    return any(t.is_complex() for t in ts)


def _values(t: torch.Tensor, cx: bool) -> torch.Tensor:
    """float64 / complex128, contiguous, on the device."""
    return t.to(device=_device(), dtype=torch.complex128 if cx else torch.float64,
                non_blocking=True).contiguous()


def _indices(t: Any) -> torch.Tensor:
    return _as_tensor(t).to(device=_device(), dtype=torch.int32, non_blocking=True).contiguous()


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else None)


def _handle_value(h: Any) -> np.ndarray:
    h = getattr(h, "handle", h)
    if isinstance(h, torch.Tensor):
        h = h.cpu().numpy()
    return np.ascontiguousarray(h, dtype=np.uint64)


def _raise_if_failed(status: torch.Tensor, what: str, check: bool) -> None:
    """The reference aborts the whole call on the first failing system
    (klujax.cpp:311-314); reading the status array back is what that costs here."""
    if not check:
        return
    bad = status != 0
    if bool(bad.any()):
        first = int(torch.nonzero(bad)[0])
        code = int(status[first])
        raise KlujaxCudaError(code, f"{what} failed (singular matrix?) at system {first}")


GROWTH_LIMIT = 1000.0 * (1.0 + 1e-9)  # 1 / tol of KLU's pivot rule (klu_defaults: tol = 1e-3)
PIVOT_DRIFT = 2  # status: the pivot-parity certificate failed (include/klujax_cuda.h)


def _bcast(t: torch.Tensor):
    """A batch axis that `_canonical` expanded from one entry is passed as ONE entry (the kernels
    take stride 0): no copy of a broadcast matrix / right-hand side."""
    if t.shape[0] > 1 and t.stride(0) == 0:
        return t[:1], 1
    return t, t.shape[0]


def _check_and_reseed(fn: str, cx: bool, L: int, nA: int, nb: int, n: int, r: int, nz: int,
                      Ai_d, Aj_d, Ax_d, b_d, x, st, gr) -> int:
    """`check=True` of the fused solves, below the C-ABI (klujax_cuda_reseed_failed_*): systems whose
    pivot-parity certificate failed are solved again with a pivot sequence seeded from one of them;
    raises like the reference on the first system that still fails (klujax.cpp:311-314).
    Returns how many systems were solved again."""
    f = getattr(_cabi.lib(), f"klujax_cuda_reseed_failed_{_suffix(cx)}")
    fb, nr = C.c_int32(-1), C.c_int32(0)
    rc = f(int(fn.startswith("tsolve")), L, nA, nb, n, r, nz, _ptr(Ai_d), _ptr(Aj_d), _ptr(Ax_d), _ptr(b_d),
           _ptr(x), _ptr(st), _ptr(gr), C.byref(fb), C.byref(nr), _stream())
    if rc != 0:
        msg = _cabi.lib().klujax_cuda_last_error().decode()
        raise KlujaxCudaError(rc, f"{msg} at system {fb.value}" if fb.value >= 0 else msg)
    return nr.value


def _canonical(Ax: torch.Tensor, x: torch.Tensor, x_name: str):
    """Shape normalisation of klujax.py:1748-1849 (validate_args).  Returns
    Ax[n_lhs, n_nz], x[n_lhs, n_col, n_rhs] and the caller-facing output shape."""
    if Ax.ndim == 0 or Ax.ndim > 2:
        raise ValueError("Ax should be 1D with shape (n_nz,) or 2D with shape (n_lhs, n_nz). "
                         f"Got: Ax.shape={tuple(Ax.shape)}.")
    if x.ndim == 0 or x.ndim > 3:
        raise ValueError(f"{x_name} should be 1D (n_col,), 2D (n_col, n_rhs) or 3D "
                         f"(n_lhs, n_col, n_rhs). Got: {x_name}.shape={tuple(x.shape)}.")
    shape = tuple(x.shape)
    if Ax.ndim == 1:
        Ax = Ax[None, :]
        if x.ndim == 1:
            x = x[None, :, None]
        elif x.ndim == 2:
            x = x[None, :, :]
    elif x.ndim == 1:
        x = x[None, :, None]
        shape = (Ax.shape[0], shape[0])
    elif x.ndim == 2:
        if Ax.shape[0] != x.shape[0] and Ax.shape[0] != 1 and x.shape[0] != 1:
            raise ValueError("Ax (2D) and " + x_name + " (2D) should have their first shape index "
                             f"`n_lhs` match. Got: Ax.shape={tuple(Ax.shape)}; "
                             f"{x_name}.shape={tuple(x.shape)}.")
        x = x[:, :, None]
        if x.shape[0] == 1 and Ax.shape[0] > 0:
            shape = (Ax.shape[0], *shape[1:])
    if Ax.shape[0] != x.shape[0] and Ax.shape[0] != 1 and x.shape[0] != 1:
        raise ValueError(f"Ax (2D) and {x_name} (3D) should have their first shape index `n_lhs` "
                         f"match. Got: Ax.shape={tuple(Ax.shape)}; {x_name}.shape={tuple(x.shape)}.")
    n_lhs = max(Ax.shape[0], x.shape[0])
    Ax = Ax.expand(n_lhs, Ax.shape[1])
    x = x.expand(n_lhs, x.shape[1], x.shape[2])
    if len(shape) == 3 and shape[0] != n_lhs:
        shape = (n_lhs, shape[1], shape[2])
    return Ax, x, shape


def _check_coo(Ai: torch.Tensor, Aj: torch.Tensor, Ax: torch.Tensor) -> None:
    if Ai.ndim != 1:
        raise ValueError(f"Ai should be 1D with shape (n_nz). Got: Ai.shape={tuple(Ai.shape)}.")
    if Aj.ndim != 1:
        raise ValueError(f"Aj should be 1D with shape (n_nz,). Got: Aj.shape={tuple(Aj.shape)}.")
    if Ai.shape[0] != Aj.shape[0] or Ax.shape[-1] != Ai.shape[0]:
        raise ValueError("n_nz mismatch between Ai, Aj and Ax: "
                         f"{tuple(Ai.shape)}, {tuple(Aj.shape)}, {tuple(Ax.shape)}.")


def _new_status(n_lhs: int):
    dev = _device()
    st = torch.empty(n_lhs, dtype=torch.int32, device=dev)
    gr = torch.empty(n_lhs, dtype=torch.float64, device=dev)
    _last["status"], _last["growth"] = st, gr
    return st, gr


def _suffix(cx: bool) -> str:
    return "c128" if cx else "f64"


# ----------------------------------------------------------------------- handle object
class KLUHandleManager:
    """RAII wrapper of a symbolic / numeric handle (klujax.py:180-237): frees on
    close / __exit__ / __del__ when it is the owner."""

    def __init__(self, handle: Any, free_callable: Callable | None, owner: bool = True) -> None:
        self.handle = handle
        self.free_callable = free_callable
        self._owner = owner
        self._freed = False

    def close(self) -> None:
        if self._freed:
            return
        if self._owner and self.free_callable:
            try:
                self.free_callable(self.handle)
            except Exception:  # noqa: BLE001 - mirror of contextlib.suppress in the reference
                pass
        self._freed = True

    def __enter__(self) -> "KLUHandleManager":
        return self

    def __exit__(self, *exc: Any) -> None:
        self.close()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def free_symbolic(symbolic: Any, dependency: Any = None) -> int:  # noqa: ARG001
    if isinstance(symbolic, KLUHandleManager):
        symbolic.close()
        return 0
    return _cabi.free_symbolic(int(_handle_value(symbolic).reshape(-1)[0]))


def free_numeric(numeric: Any, dependency: Any = None) -> int:  # noqa: ARG001
    if isinstance(numeric, KLUHandleManager):
        numeric.close()
        return 0
    return _cabi.free_numeric(_handle_value(numeric).reshape(-1))


# ----------------------------------------------------------------------------- API
def coalesce(Ai: Any, Aj: Any, Ax: Any):
    """Sum duplicate COO entries and sort by (row, col) (klujax.py:134-174)."""
    Ai, Aj, Ax = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax)
    n = int(max(int(Ai.max()), int(Aj.max()))) + 1 if Ai.numel() else 1
    key = Ai.to(torch.int64) * n + Aj.to(torch.int64)
    ukey, inv = torch.unique(key, return_inverse=True)
    out = torch.zeros((*Ax.shape[:-1], ukey.numel()), dtype=Ax.dtype, device=Ax.device)
    out.index_add_(-1, inv, Ax)
    return (ukey // n).to(Ai.dtype), (ukey % n).to(Aj.dtype), out


def analyze(Ai: Any, Aj: Any, n_col: int) -> KLUHandleManager:
    """Symbolic analysis (BTF + AMD) on the host; klujax.py:297-312."""
    Ai_h = _as_tensor(Ai).detach().cpu().numpy().astype(np.int32)
    Aj_h = _as_tensor(Aj).detach().cpu().numpy().astype(np.int32)
    if Ai_h.ndim != 1 or Aj_h.ndim != 1 or Ai_h.size != Aj_h.size:
        raise ValueError("Ai and Aj must be 1D arrays of equal length")
    h = _cabi.analyze(Ai_h, Aj_h, int(n_col))
    return KLUHandleManager(np.uint64(h), free_symbolic, owner=True)


_PIPELINE_MIN_SYSTEMS = 8192   # host-resident batches at least this large are streamed in chunks
_PIPELINE_CHUNKS = int(os.environ.get("KLUJAX_PIPELINE_CHUNKS", "16"))
_PIPELINE_STREAMS = int(os.environ.get("KLUJAX_PIPELINE_STREAMS", "2"))
_streams: dict = {}


def _copy_streams(dev: torch.device):
    if dev not in _streams:
        _streams[dev] = [torch.cuda.Stream(device=dev) for _ in range(max(1, _PIPELINE_STREAMS))]
    return _streams[dev]


def _solve_symbol(fn: str, Ai, Aj, Ax, b, symbolic, check: bool, out=None):
    Ai_t, Aj_t, Ax_t, b_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax), _as_tensor(b)
    _check_coo(Ai_t, Aj_t, Ax_t)
    Ax_t, b_t, shape = _canonical(Ax_t, b_t, "b")
    cx = _is_complex(Ax_t, b_t)
    Ai_d, Aj_d = _indices(Ai_t), _indices(Aj_t)
    sym = int(_handle_value(symbolic).reshape(-1)[0])
    f = getattr(_cabi.lib(), f"klujax_cuda_{fn}_{_suffix(cx)}")
    L, n, r = b_t.shape
    nz = Ai_d.numel()
    vdt = torch.complex128 if cx else torch.float64
    host_batch = (not Ax_t.is_cuda and not b_t.is_cuda and L >= _PIPELINE_MIN_SYSTEMS
                  and Ax_t.dtype == vdt and b_t.dtype == vdt
                  and Ax_t.is_contiguous() and b_t.is_contiguous())
    out_v = None
    if out is not None:
        # extension of the reference signature: a host (ideally pinned) result buffer; with a
        # host-resident batch the device->host copy of chunk k overlaps the work on chunk k+1
        if (not isinstance(out, torch.Tensor) or out.is_cuda or out.dtype != vdt
                or out.numel() != b_t.numel() or not out.is_contiguous()):
            raise ValueError("out must be a contiguous CPU tensor with the shape and dtype of the result")
        out_v = out.view(L, n, r)
    nA, nb = L, L
    if not host_batch:
        Ax_b, nA = _bcast(Ax_t)
        b_b, nb = _bcast(b_t)
        Ax_d, b_d = _values(Ax_b, cx), _values(b_b, cx)
        x = torch.empty((L, n, r), dtype=vdt, device=Ax_d.device)
        st, gr = _new_status(L)
        fex = getattr(_cabi.lib(), f"klujax_cuda_{fn}_ex_{_suffix(cx)}")
        _cabi.check(fex(C.c_uint64(sym), L, nA, nb, n, r, nz, _ptr(Ai_d), _ptr(Aj_d), _ptr(Ax_d),
                        _ptr(b_d), _ptr(x), _ptr(st), _ptr(gr), 0, None, None, _stream()))
        if out_v is not None:
            out_v.copy_(x, non_blocking=True)
    else:
        # Host-resident batch: systems are independent, so the batch is streamed in chunks on
        # two streams - the host->device copy of chunk k+1 overlaps the kernels of chunk k.
        dev = _device()
        Ax_d = torch.empty((L, nz), dtype=vdt, device=dev)
        b_d = torch.empty((L, n, r), dtype=vdt, device=dev)
        x = torch.empty_like(b_d)
        st, gr = _new_status(L)
        cur = torch.cuda.current_stream()
        streams = _copy_streams(dev)
        bounds = [L * k // _PIPELINE_CHUNKS for k in range(_PIPELINE_CHUNKS + 1)]
        for s in streams:
            s.wait_stream(cur)
        for k in range(_PIPELINE_CHUNKS):
            lo, hi = bounds[k], bounds[k + 1]
            s = streams[k % len(streams)]
            with torch.cuda.stream(s):
                Ax_d[lo:hi].copy_(Ax_t[lo:hi], non_blocking=True)
                b_d[lo:hi].copy_(b_t[lo:hi], non_blocking=True)
                _cabi.check(f(C.c_uint64(sym), hi - lo, n, r, nz, _ptr(Ai_d), _ptr(Aj_d),
                              _ptr(Ax_d[lo:hi]), _ptr(b_d[lo:hi]), _ptr(x[lo:hi]), _ptr(st[lo:hi]),
                              _ptr(gr[lo:hi]), C.c_void_p(s.cuda_stream)))
                if out_v is not None:
                    out_v[lo:hi].copy_(x[lo:hi], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
        for t in (Ax_d, b_d, x, st, gr, Ai_d, Aj_d):
            for s in streams:
                t.record_stream(s)
    if check and _check_and_reseed(fn, cx, L, nA, nb, n, r, nz, Ai_d, Aj_d, Ax_d, b_d, x, st, gr):
        if out_v is not None:
            out_v.copy_(x, non_blocking=True)
    if out_v is not None:
        torch.cuda.current_stream().synchronize()  # the caller may read `out` on return
        return out.view(shape)
    return x.reshape(shape)


def solve_with_symbol(Ai, Aj, Ax, b, symbolic, check: bool = True, out=None) -> torch.Tensor:
    """Factor + solve with a precomputed symbolic analysis (klujax.py:376-393).  `out`
    (extension): a CPU result buffer filled with pipelined device->host copies."""
    return _solve_symbol("solve_with_symbol", Ai, Aj, Ax, b, symbolic, check, out)


def tsolve_with_symbol(Ai, Aj, Ax, b, symbolic, check: bool = True, out=None) -> torch.Tensor:
    """Factor + transpose solve A^T x = b (plain transpose; klujax.py:416-440)."""
    return _solve_symbol("tsolve_with_symbol", Ai, Aj, Ax, b, symbolic, check, out)


def solve(Ai, Aj, Ax, b, check: bool = True) -> torch.Tensor:
    """analyze + factor + solve in one call (klujax.py:63-96)."""
    Ai_t, Aj_t, Ax_t, b_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax), _as_tensor(b)
    _check_coo(Ai_t, Aj_t, Ax_t)
    Ax_t, b_t, shape = _canonical(Ax_t, b_t, "b")
    cx = _is_complex(Ax_t, b_t)
    Ai_h = np.ascontiguousarray(Ai_t.detach().cpu().numpy(), dtype=np.int32)
    Aj_h = np.ascontiguousarray(Aj_t.detach().cpu().numpy(), dtype=np.int32)
    Ax_d, b_d = _values(Ax_t, cx), _values(b_t, cx)
    L, n, r = b_d.shape
    x = torch.empty_like(b_d)
    st, gr = _new_status(L)
    f = getattr(_cabi.lib(), f"klujax_cuda_solve_{_suffix(cx)}")
    _cabi.check(f(L, n, r, int(Ai_h.size), _cabi.i32p(Ai_h), _cabi.i32p(Aj_h), _ptr(Ax_d), _ptr(b_d),
                  _ptr(x), _ptr(st), _ptr(gr), _stream()))
    if check:
        _check_and_reseed("solve_with_symbol", cx, L, L, L, n, r, int(Ai_h.size), _indices(Ai_t), _indices(Aj_t),
                          Ax_d, b_d, x, st, gr)
    return x.reshape(shape)


def dot(Ai, Aj, Ax, x) -> torch.Tensor:
    """Batched COO sparse matrix times vector(s) (klujax.py:99-131)."""
    Ai_t, Aj_t, Ax_t, x_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax), _as_tensor(x)
    _check_coo(Ai_t, Aj_t, Ax_t)
    Ax_t, x_t, shape = _canonical(Ax_t, x_t, "x")
    cx = _is_complex(Ax_t, x_t)
    Ai_d, Aj_d = _indices(Ai_t), _indices(Aj_t)
    Ax_d, x_d = _values(Ax_t, cx), _values(x_t, cx)
    L, n, r = x_d.shape
    out = torch.empty_like(x_d)
    f = getattr(_cabi.lib(), f"klujax_cuda_dot_{_suffix(cx)}")
    _cabi.check(f(L, n, r, Ai_d.numel(), _ptr(Ai_d), _ptr(Aj_d), _ptr(Ax_d), _ptr(x_d), _ptr(out),
                  _stream()))
    return out.reshape(shape)


def dot_transpose_ax(Ai, Aj, ct, x) -> torch.Tensor:
    """dAx[m, e] = sum_k ct[m, Ai[e], k] * x[m, Aj[e], k]: the transpose of `dot` with respect to
    the values (klujax.py:1712-1714, pure jnp in the reference).  ct and x: (n_lhs, n_col, n_rhs)
    or lower rank as for `dot`; returns (n_lhs, n_nz)."""
    Ai_t, Aj_t, ct_t, x_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(ct), _as_tensor(x)
    cx = _is_complex(ct_t, x_t)
    Ai_d, Aj_d = _indices(Ai_t), _indices(Aj_t)
    while ct_t.ndim < 3:
        ct_t = ct_t.unsqueeze(0) if ct_t.ndim == 2 else ct_t.unsqueeze(-1)
    while x_t.ndim < 3:
        x_t = x_t.unsqueeze(0) if x_t.ndim == 2 else x_t.unsqueeze(-1)
    if ct_t.shape != x_t.shape:
        ct_t, x_t = torch.broadcast_tensors(ct_t, x_t)
    ct_d, x_d = _values(ct_t, cx), _values(x_t, cx)
    L, n, r = x_d.shape
    out = torch.empty((L, Ai_d.numel()), dtype=x_d.dtype, device=x_d.device)
    f = getattr(_cabi.lib(), f"klujax_cuda_dot_transpose_ax_{_suffix(cx)}")
    _cabi.check(f(L, n, r, Ai_d.numel(), _ptr(Ai_d), _ptr(Aj_d), _ptr(ct_d), _ptr(x_d), _ptr(out),
                  _stream()))
    return out


def solve_with_symbol_value_and_jvp(Ai, Aj, Ax, b, symbolic, t_Ax=None, t_b=None, check: bool = True):
    """x and dx of `solve_with_symbol` from ONE factorization ("next" row f2).

    Same formula as the reference's JVP rule (klujax.py:1856-1897: x = A^-1 b,
    dx = A^-1 (t_b - dot(Ai, Aj, t_Ax, x))), which binds `solve_with_symbol` twice and so
    factors the same A twice; here A is factored once and both right-hand sides are solved
    from the device-resident factors."""
    num = factor(Ai, Aj, Ax, symbolic, check=check)
    try:
        x = solve_with_numeric(num, b, symbolic)
        rhs = None if t_b is None else _as_tensor(t_b).to(x.device)
        if t_Ax is not None:
            dAx = dot(Ai, Aj, t_Ax, x)
            rhs = -dAx if rhs is None else rhs - dAx
        dx = torch.zeros_like(x) if rhs is None else solve_with_numeric(num, rhs, symbolic)
    finally:
        num.close()
    return x, dx


def solve_with_symbol_value_and_vjp(Ai, Aj, Ax, b, symbolic, ct, check: bool = True):
    """x, b_bar and Ax_bar of `solve_with_symbol` from ONE factorization ("next" row f2).

    b_bar = A^-T ct is the reference's transpose rule for b (klujax.py:1948-1957, plain
    transpose also for complex); Ax_bar[m, e] = -sum_k b_bar[m, Ai[e], k] x[m, Aj[e], k] is the
    transpose of the JVP's `dot` term (klujax.py:1886-1890 with dot_transpose :1712-1714).
    The reference factors A three times for this (primal, JVP, transpose)."""
    num = factor(Ai, Aj, Ax, symbolic, check=check)
    try:
        x = solve_with_numeric(num, b, symbolic)
        b_bar = tsolve_with_numeric(num, ct, symbolic)
    finally:
        num.close()
    Ax_bar = -dot_transpose_ax(Ai, Aj, b_bar, x)
    Ax_t = _as_tensor(Ax)
    return x, b_bar, Ax_bar.reshape(Ax_t.shape) if Ax_bar.numel() == Ax_t.numel() else Ax_bar


def factor(Ai, Aj, Ax, symbolic, check: bool = True) -> KLUHandleManager:
    """Numeric factorization kept on the device behind u64 handles (klujax.py:451-468)."""
    Ai_t, Aj_t, Ax_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax)
    _check_coo(Ai_t, Aj_t, Ax_t)
    if Ax_t.ndim == 1:
        Ax_t = Ax_t[None, :]
    if Ax_t.ndim != 2:
        raise ValueError(f"Ax should be 1D or 2D. Got: Ax.shape={tuple(Ax_t.shape)}.")
    cx = _is_complex(Ax_t)
    Ai_d, Aj_d, Ax_d = _indices(Ai_t), _indices(Aj_t), _values(Ax_t, cx)
    L = Ax_d.shape[0]
    num = np.zeros(L, np.uint64)
    st, gr = _new_status(L)
    sym = int(_handle_value(symbolic).reshape(-1)[0])
    # check=True: the reference's semantics (klu_factor pivots every system, klujax.cpp:679) -
    # systems whose pivot-parity certificate fails get their own pivot sequence in a child batch,
    # below the C-ABI; a system that is singular under its own pivoting fails the call and the
    # partial batch is freed (klujax.cpp:684-687)
    f = getattr(_cabi.lib(), f"klujax_cuda_factor_ex_{_suffix(cx)}")
    fb = C.c_int32(-1)
    rc = f(C.c_uint64(sym), L, Ai_d.numel(), _ptr(Ai_d), _ptr(Aj_d), _ptr(Ax_d),
           _cabi.u64p(num), _ptr(st), _ptr(gr), int(bool(check)), C.byref(fb), _stream())
    if rc != 0:
        msg = _cabi.lib().klujax_cuda_last_error().decode()
        raise KlujaxCudaError(rc, f"{msg} at system {fb.value}" if fb.value >= 0 else msg)
    return KLUHandleManager(num, free_numeric, owner=True)


def refactor(Ai, Aj, Ax, numeric, symbolic, check: bool = True) -> KLUHandleManager:
    """In-place numeric refactorization; returns a non-owning manager of the same
    handles (klujax.py:479-512)."""
    Ai_t, Aj_t, Ax_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax)
    _check_coo(Ai_t, Aj_t, Ax_t)
    if Ax_t.ndim == 1:
        Ax_t = Ax_t[None, :]
    cx = _is_complex(Ax_t)
    Ai_d, Aj_d, Ax_d = _indices(Ai_t), _indices(Aj_t), _values(Ax_t, cx)
    L = Ax_d.shape[0]
    nin = _handle_value(numeric).reshape(-1)
    if nin.size != L:
        raise KlujaxCudaError(_cabi.KLU_INVALID, "numeric and Ax batch size mismatch")
    nout = np.zeros_like(nin)
    st, gr = _new_status(L)
    sym = int(_handle_value(symbolic).reshape(-1)[0])
    f = getattr(_cabi.lib(), f"klujax_cuda_refactor_{_suffix(cx)}")
    _cabi.check(f(C.c_uint64(sym), L, Ai_d.numel(), _ptr(Ai_d), _ptr(Aj_d), _ptr(Ax_d),
                  _cabi.u64p(nin), _cabi.u64p(nout), _ptr(st), _ptr(gr), _stream()))
    _raise_if_failed(st, "klu_refactor", check)
    return KLUHandleManager(nout, free_numeric, owner=False)


def refactor_and_solve(Ai, Aj, Ax, b, numeric, symbolic, check: bool = True):
    """Fused refactor + solve (klujax.py:612-648).  Returns (x, non-owning numeric)."""
    Ai_t, Aj_t, Ax_t, b_t = _as_tensor(Ai), _as_tensor(Aj), _as_tensor(Ax), _as_tensor(b)
    _check_coo(Ai_t, Aj_t, Ax_t)
    Ax_t, b_t, shape = _canonical(Ax_t, b_t, "b")
    cx = _is_complex(Ax_t, b_t)
    Ai_d, Aj_d = _indices(Ai_t), _indices(Aj_t)
    Ax_d, b_d = _values(Ax_t, cx), _values(b_t, cx)
    L, n, r = b_d.shape
    nin = _handle_value(numeric).reshape(-1)
    if nin.size != L:
        raise KlujaxCudaError(_cabi.KLU_INVALID, "numeric and Ax batch size mismatch")
    nout = np.zeros_like(nin)
    x = torch.empty_like(b_d)
    st, gr = _new_status(L)
    sym = int(_handle_value(symbolic).reshape(-1)[0])
    f = getattr(_cabi.lib(), f"klujax_cuda_refactor_and_solve_{_suffix(cx)}")
    _cabi.check(f(C.c_uint64(sym), L, n, r, Ai_d.numel(), _ptr(Ai_d), _ptr(Aj_d), _ptr(Ax_d),
                  _ptr(b_d), _cabi.u64p(nin), _ptr(x), _cabi.u64p(nout), _ptr(st), _ptr(gr),
                  _stream()))
    _raise_if_failed(st, "klu_refactor", check)
    return x.reshape(shape), KLUHandleManager(nout, free_numeric, owner=False)


def _solve_numeric(fn: str, numeric, b, symbolic) -> torch.Tensor:
    b_t = _as_tensor(b)
    num = _handle_value(numeric).reshape(-1)
    n_numeric = int(num.size)
    # rank parsing of klujax.cpp:1011-1059: the 2-D case is decided by the numeric count
    if b_t.ndim == 1:
        b3 = b_t[None, :, None]
    elif b_t.ndim == 2:
        if n_numeric > 1 and b_t.shape[0] == n_numeric:
            b3 = b_t[:, :, None]
        else:
            b3 = b_t[None, :, :]
    elif b_t.ndim == 3:
        b3 = b_t
    else:
        raise KlujaxCudaError(_cabi.KLU_INVALID, "b must be 1D, 2D, or 3D")
    cx = _is_complex(b3)
    b_d = _values(b3, cx)
    Lb, n, r = b_d.shape
    if n_numeric != 1 and Lb != 1 and n_numeric != Lb:
        raise KlujaxCudaError(_cabi.KLU_INVALID, "numeric and b batch size mismatch")
    L = max(n_numeric, Lb)
    x = torch.empty((L, n, r), dtype=b_d.dtype, device=b_d.device)
    sym = int(_handle_value(symbolic).reshape(-1)[0])
    f = getattr(_cabi.lib(), f"klujax_cuda_{fn}_{_suffix(cx)}")
    _cabi.check(f(C.c_uint64(sym), n_numeric, _cabi.u64p(num), Lb, n, r, _ptr(b_d), _ptr(x),
                  _stream()))
    # output has the rank of b (klujax.cpp:1061-1064); a broadcast batch keeps its size
    if b_t.ndim == 1:
        return x.reshape(n) if L == 1 else x.reshape(L, n)
    if b_t.ndim == 2:
        if n_numeric > 1 and b_t.shape[0] == n_numeric:
            return x.reshape(L, n)
        return x.reshape(n, r) if L == 1 else x
    return x


def solve_with_numeric(numeric, b, symbolic) -> torch.Tensor:
    """Triangular solves with stored factors (klujax.py:523-546)."""
    return _solve_numeric("solve_with_numeric", numeric, b, symbolic)


def tsolve_with_numeric(numeric, b, symbolic) -> torch.Tensor:
    """Transpose solve with stored factors (klujax.py:559-586)."""
    return _solve_numeric("tsolve_with_numeric", numeric, b, symbolic)
