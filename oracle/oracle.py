"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Mirrors the reference's FFI targets (/root/reference/klujax.cpp:1386-1429) on
canonical shapes: Ai, Aj int32[nz]; Ax T[L, nz]; b T[L, n, r].  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libklu_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory)."""
    srcs = ["btf_oracle.c", "amd_oracle.c", "klu_oracle.c", "klujax_oracle.c",
            "klu_kernel.inc", "klu_oracle.h"]
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(os.path.join(_HERE, s))
        and os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(_SO)
        for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.kjo_symbolic_ptr.restype = C.c_void_p
        _lib.kjo_numeric_ptr.restype = C.c_void_p
    return _lib


class KoSymbolic(C.Structure):
    _fields_ = [("n", C.c_int), ("nz", C.c_int), ("nblocks", C.c_int),
                ("maxblock", C.c_int), ("nzoff", C.c_int),
                ("structural_rank", C.c_int),
                ("P", C.POINTER(C.c_int)), ("Q", C.POINTER(C.c_int)),
                ("R", C.POINTER(C.c_int)), ("Lnz", C.POINTER(C.c_double))]


class KoNumeric(C.Structure):
    _fields_ = [("n", C.c_int), ("nblocks", C.c_int), ("nzoff", C.c_int),
                ("is_complex", C.c_int), ("lnz", C.c_int), ("unz", C.c_int),
                ("noffdiag", C.c_int),
                ("Pnum", C.POINTER(C.c_int)), ("Pinv", C.POINTER(C.c_int)),
                ("Lp", C.POINTER(C.c_int)), ("Llen", C.POINTER(C.c_int)),
                ("Up", C.POINTER(C.c_int)), ("Ulen", C.POINTER(C.c_int)),
                ("Li", C.POINTER(C.c_int)), ("Ui", C.POINTER(C.c_int)),
                ("Lx", C.POINTER(C.c_double)), ("Ux", C.POINTER(C.c_double)),
                ("Udiag", C.POINTER(C.c_double)), ("Rs", C.POINTER(C.c_double)),
                ("Offp", C.POINTER(C.c_int)), ("Offi", C.POINTER(C.c_int)),
                ("Offx", C.POINTER(C.c_double)), ("Xwork", C.POINTER(C.c_double)),
                ("lcap", C.c_int64), ("ucap", C.c_int64)]


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _vals(a):
    a = np.asarray(a)
    cplx = np.iscomplexobj(a)
    a = np.ascontiguousarray(a, dtype=np.complex128 if cplx else np.float64)
    return a, int(cplx)


class OracleError(RuntimeError):
    pass


def coo_to_csc(n_col, Ai, Aj):
    Ai, Aj = _i32(Ai), _i32(Aj)
    nz = Ai.size
    Bi = np.zeros(nz, np.int32)
    Bp = np.zeros(n_col + 1, np.int32)
    Bk = np.zeros(nz, np.int32)
    lib().kjo_coo_to_csc(n_col, nz, _ip(Ai), _ip(Aj), _ip(Bi), _ip(Bp), _ip(Bk))
    return Bp, Bi, Bk


def analyze(Ai, Aj, n_col) -> int:
    Ai, Aj = _i32(Ai), _i32(Aj)
    h = C.c_uint64(0)
    rc = lib().kjo_analyze(int(n_col), Ai.size, _ip(Ai), _ip(Aj), C.byref(h))
    if rc:
        raise OracleError(f"analyze failed rc={rc}")
    return h.value


def free_symbolic(h) -> int:
    return lib().kjo_free_symbolic(C.c_uint64(int(h)))


def free_numeric(hs) -> int:
    hs = np.ascontiguousarray(np.atleast_1d(hs), dtype=np.uint64)
    return lib().kjo_free_numeric(hs.size, _u64p(hs))


def symbolic_info(h) -> dict:
    s = C.cast(lib().kjo_symbolic_ptr(C.c_uint64(int(h))), C.POINTER(KoSymbolic)).contents
    n = s.n
    return dict(n=n, nz=s.nz, nblocks=s.nblocks, maxblock=s.maxblock, nzoff=s.nzoff,
                structural_rank=s.structural_rank,
                P=np.ctypeslib.as_array(s.P, (n,)).copy(),
                Q=np.ctypeslib.as_array(s.Q, (n,)).copy(),
                R=np.ctypeslib.as_array(s.R, (s.nblocks + 1,)).copy())


def numeric_info(h) -> dict:
    """Pnum and the L/U patterns per column (sorted), in pivot order."""
    N = C.cast(lib().kjo_numeric_ptr(C.c_uint64(int(h))), C.POINTER(KoNumeric)).contents
    n = N.n
    es = 2 if N.is_complex else 1
    Lp = np.ctypeslib.as_array(N.Lp, (n,)).copy()
    Ll = np.ctypeslib.as_array(N.Llen, (n,)).copy()
    Up = np.ctypeslib.as_array(N.Up, (n,)).copy()
    Ul = np.ctypeslib.as_array(N.Ulen, (n,)).copy()
    lt, ut = int(Ll.sum()), int(Ul.sum())
    lend = int((Lp + Ll).max()) if n else 0
    uend = int((Up + Ul).max()) if n else 0
    Li = np.ctypeslib.as_array(N.Li, (max(lend, 1),)).copy()
    Ui = np.ctypeslib.as_array(N.Ui, (max(uend, 1),)).copy()
    Lx = np.ctypeslib.as_array(N.Lx, (max(lend, 1) * es,)).copy()
    Ux = np.ctypeslib.as_array(N.Ux, (max(uend, 1) * es,)).copy()
    Ud = np.ctypeslib.as_array(N.Udiag, (n * es,)).copy()
    if es == 2:
        Lx, Ux, Ud = Lx.view(np.complex128), Ux.view(np.complex128), Ud.view(np.complex128)
    return dict(n=n, lnz=N.lnz, unz=N.unz, noffdiag=N.noffdiag, nzoff=N.nzoff,
                Pnum=np.ctypeslib.as_array(N.Pnum, (n,)).copy(),
                Lp=Lp, Llen=Ll, Up=Up, Ulen=Ul, Li=Li, Ui=Ui, Lx=Lx, Ux=Ux,
                Udiag=Ud, Rs=np.ctypeslib.as_array(N.Rs, (n,)).copy(),
                Offp=np.ctypeslib.as_array(N.Offp, (n + 1,)).copy(),
                Offi=np.ctypeslib.as_array(N.Offi, (max(N.nzoff, 1),)).copy()[:N.nzoff],
                nnz_l_off=lt, nnz_u_off=ut)


def _canon(Ax, b):
    Ax, c1 = _vals(Ax)
    b, c2 = _vals(b)
    if c1 != c2:
        Ax = Ax.astype(np.complex128)
        b = b.astype(np.complex128)
    assert Ax.ndim == 2 and b.ndim == 3 and Ax.shape[0] == b.shape[0]
    return Ax, b, int(c1 or c2)


def solve(Ai, Aj, Ax, b):
    Ai, Aj = _i32(Ai), _i32(Aj)
    Ax, b, cx = _canon(Ax, b)
    L, n, r = b.shape
    x = np.empty_like(b)
    rc = lib().kjo_solve(cx, L, n, r, Ai.size, _ip(Ai), _ip(Aj),
                         _dp(Ax.view(np.float64)), _dp(b.view(np.float64)),
                         _dp(x.view(np.float64)))
    if rc:
        raise OracleError(f"solve failed rc={rc} (singular matrix?)")
    return x


def solve_with_symbol(Ai, Aj, Ax, b, sym, transpose=False, lo=None, hi=None, out=None):
    Ai, Aj = _i32(Ai), _i32(Aj)
    Ax, b, cx = _canon(Ax, b)
    L, n, r = b.shape
    x = np.empty_like(b) if out is None else out
    lo = 0 if lo is None else lo
    hi = L if hi is None else hi
    rc = lib().kjo_solve_with_symbol_range(
        cx, int(transpose), lo, hi, n, r, Ai.size, _ip(Ai), _ip(Aj),
        _dp(Ax.view(np.float64)), _dp(b.view(np.float64)), C.c_uint64(int(sym)),
        _dp(x.view(np.float64)))
    if rc:
        raise OracleError(f"solve_with_symbol failed rc={rc} (singular matrix?)")
    return x


def factor(Ai, Aj, Ax, sym):
    Ai, Aj = _i32(Ai), _i32(Aj)
    Ax, cx = _vals(Ax)
    L = Ax.shape[0]
    num = np.zeros(L, np.uint64)
    rc = lib().kjo_factor(cx, L, Ai.size, _ip(Ai), _ip(Aj), _dp(Ax.view(np.float64)),
                          C.c_uint64(int(sym)), _u64p(num))
    if rc:
        raise OracleError(f"factor failed rc={rc} (singular matrix?)")
    return num


def refactor(Ai, Aj, Ax, sym, num):
    Ai, Aj = _i32(Ai), _i32(Aj)
    Ax, cx = _vals(Ax)
    num = np.ascontiguousarray(num, dtype=np.uint64)
    out = np.zeros_like(num)
    rc = lib().kjo_refactor(cx, Ax.shape[0], Ai.size, _ip(Ai), _ip(Aj),
                            _dp(Ax.view(np.float64)), C.c_uint64(int(sym)),
                            _u64p(num), _u64p(out))
    if rc:
        raise OracleError(f"refactor failed rc={rc}")
    return out


def solve_with_numeric(sym, num, b, transpose=False):
    """b canonical [L_b, n, r]; numeric u64[L or 1] (broadcast rules of
    /root/reference/klujax.cpp:1044-1059)."""
    num = np.ascontiguousarray(np.atleast_1d(num), dtype=np.uint64)
    b, cx = _vals(b)
    Lb, n, r = b.shape
    L = max(Lb, num.size)
    x = np.empty((L, n, r), b.dtype)
    rc = lib().kjo_solve_with_numeric(cx, int(transpose), num.size, Lb, n, r,
                                      C.c_uint64(int(sym)), _u64p(num),
                                      _dp(b.view(np.float64)), _dp(x.view(np.float64)))
    if rc:
        raise OracleError(f"solve_with_numeric failed rc={rc}")
    return x


def refactor_and_solve(Ai, Aj, Ax, b, sym, num):
    Ai, Aj = _i32(Ai), _i32(Aj)
    Ax, b, cx = _canon(Ax, b)
    L, n, r = b.shape
    num = np.ascontiguousarray(num, dtype=np.uint64)
    out = np.zeros_like(num)
    x = np.empty_like(b)
    rc = lib().kjo_refactor_and_solve(cx, L, n, r, Ai.size, _ip(Ai), _ip(Aj),
                                      _dp(Ax.view(np.float64)), _dp(b.view(np.float64)),
                                      C.c_uint64(int(sym)), _u64p(num),
                                      _dp(x.view(np.float64)), _u64p(out))
    if rc:
        raise OracleError(f"refactor_and_solve failed rc={rc}")
    return x, out
