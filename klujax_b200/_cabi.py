"""ctypes binding of include/klujax_cuda.h (the drop-in boundary).

The library is loaded lazily; a missing or unloadable library is a hard error —
there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libklujax_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "klujax_cuda.h")

_lib = None

KLU_OK, KLU_SINGULAR, KLU_OUT_OF_MEMORY, KLU_INVALID, KLU_TOO_LARGE = 0, 1, -2, -3, -4


class KlujaxCudaError(RuntimeError):
    """Raised where the reference raises XlaRuntimeError from an ffi::Error."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{msg} [status {code}]")
        self.code = code


def declared_symbols() -> list[str]:
    """Every function name include/klujax_cuda.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(klujax_cuda_\w+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m klujax_b200.build` "
                "(nvcc, sm_100a). klujax_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.klujax_cuda_last_error.restype = C.c_char_p
        for name in declared_symbols():
            getattr(_lib, name)  # fail loudly on a missing export
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise KlujaxCudaError(rc, lib().klujax_cuda_last_error().decode())


def i32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def u64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def vp(ptr: int | None):
    return C.c_void_p(int(ptr) if ptr else None)


# ---- host-only entry points -------------------------------------------------
def analyze(Ai: np.ndarray, Aj: np.ndarray, n_col: int) -> int:
    Ai = np.ascontiguousarray(Ai, dtype=np.int32)
    Aj = np.ascontiguousarray(Aj, dtype=np.int32)
    h = C.c_uint64(0)
    check(lib().klujax_cuda_analyze(int(n_col), int(Ai.size), i32p(Ai), i32p(Aj), C.byref(h)))
    return h.value


def free_symbolic(h: int) -> int:
    out = C.c_int32(0)
    check(lib().klujax_cuda_free_symbolic(C.c_uint64(int(h)), C.byref(out)))
    return out.value


def free_numeric(hs) -> int:
    hs = np.ascontiguousarray(np.atleast_1d(hs), dtype=np.uint64)
    out = C.c_int32(0)
    check(lib().klujax_cuda_free_numeric(int(hs.size), u64p(hs), C.byref(out)))
    return out.value


def seed_host(h: int, Ax0: np.ndarray) -> None:
    cx = int(np.iscomplexobj(Ax0))
    Ax0 = np.ascontiguousarray(Ax0, dtype=np.complex128 if cx else np.float64)
    check(lib().klujax_cuda_seed_host(C.c_uint64(int(h)), cx,
                                      Ax0.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))))


_INFO_KEYS = ("n", "nblocks", "maxblock", "nzoff", "structural_rank", "nnz_l", "nnz_u", "seeded",
              "regime", "levels", "mac", "noffdiag", "nbatch", "stream_words")


def symbolic_info(h: int, is_complex: bool) -> dict:
    out = np.zeros(14, np.int64)
    check(lib().klujax_cuda_symbolic_info(C.c_uint64(int(h)), int(is_complex),
                                          out.ctypes.data_as(C.POINTER(C.c_int64))))
    return dict(zip(_INFO_KEYS, (int(v) for v in out)))


def symbolic_perm(h: int) -> dict:
    info = symbolic_info(h, False)
    n, nb = info["n"], info["nblocks"]
    P, Q, R = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(nb + 1, np.int32)
    check(lib().klujax_cuda_symbolic_perm(C.c_uint64(int(h)), i32p(P), i32p(Q), i32p(R)))
    return dict(P=P, Q=Q, R=R, **info)


def symbolic_pattern(h: int, is_complex: bool) -> dict:
    info = symbolic_info(h, is_complex)
    n = info["n"]
    Pnum = np.zeros(n, np.int32)
    Lp, Up = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    Li, Ui = np.zeros(max(info["nnz_l"], 1), np.int32), np.zeros(max(info["nnz_u"], 1), np.int32)
    check(lib().klujax_cuda_symbolic_pattern(C.c_uint64(int(h)), int(is_complex), i32p(Pnum),
                                             i32p(Lp), i32p(Li), i32p(Up), i32p(Ui)))
    return dict(Pnum=Pnum, Lp=Lp, Li=Li[:info["nnz_l"]], Up=Up, Ui=Ui[:info["nnz_u"]], **info)
