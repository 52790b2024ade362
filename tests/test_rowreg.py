"""Register-resident regime (klujax_b200/csrc/rowreg.*): the planner's instruction list is run by
its host interpreter (klujax_cuda_rowreg_selftest, test infrastructure) and checked against dense
solves and the oracle; the generated PTX must assemble for sm_100a.  No GPU needed."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from klujax_b200 import _cabi, synth
from util import dense, relerr


def selftest(h, cx, Ax, b, transpose=0, ptx=None, env=None):
    lib = _cabi.lib()
    dt = np.complex128 if cx else np.float64
    Ax = np.ascontiguousarray(Ax, dt)
    b = np.ascontiguousarray(b, dt)
    x = np.zeros_like(b)
    st, g, stats = C.c_int32(0), C.c_double(0), np.zeros(16, np.int64)
    f = lambda a: a.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.klujax_cuda_rowreg_selftest(C.c_uint64(h), int(cx), int(transpose), int(b.shape[1]), f(Ax), f(b), f(x),
                                         C.byref(st), C.byref(g), stats.ctypes.data_as(C.POINTER(C.c_int64)),
                                         ptx.encode() if ptx else None)
    return rc, x, st.value, g.value, stats


@pytest.mark.parametrize("n,seed,r", [(64, 1, 1), (64, 1, 4), (128, 2, 1)])
@pytest.mark.parametrize("transpose", [0, 1])
def test_sax_plan_matches_dense(n, seed, r, transpose, monkeypatch, tmp_path):
    monkeypatch.setenv("KLUJAX_CUDA_RR_MAXREGS", "96")
    pat = synth.SaxPattern(n=n, seed=seed)
    Ax = pat.values(0, 2)
    h = _cabi.analyze(pat.Ai, pat.Aj, n)
    _cabi.seed_host(h, Ax[0])
    rng = np.random.default_rng(5)
    b = rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r))
    ptx = str(tmp_path / "k.ptx")
    rc, x, st, g, stats = selftest(h, True, Ax[1], b, transpose, ptx=ptx)
    assert rc == 0 and st == 0 and g <= 1000.0
    A = dense(pat.Ai, pat.Aj, Ax[1], n)
    assert relerr(x, np.linalg.solve(A.T if transpose else A, b)) < 1e-12
    assert stats[2] >= stats[1]          # every multiply-add instruction does useful work
    if shutil.which("ptxas") and n == 64 and not transpose:
        out = subprocess.run(["ptxas", "-arch=sm_100a", "-v", "-o", str(tmp_path / "k.cubin"), ptx],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "0 bytes spill stores" in out.stderr
    _cabi.free_symbolic(h)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("shape", [(3, 8, 1), (5, 15, 2), (20, 60, 3), (40, 160, 4)])
def test_reference_style_cases(dtype, shape, oracle):
    n, nnz, r = shape
    Ai, Aj, Ax, b = synth.ref_style_case(n, nnz, 3, r, dtype)
    cx = dtype is np.complex128
    h = _cabi.analyze(Ai, Aj, n)
    _cabi.seed_host(h, Ax[0])
    so = oracle.analyze(Ai, Aj, n)
    ref = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
    reft = oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)
    oracle.free_symbolic(so)
    for m in range(3):
        rc, x, st, g, _ = selftest(h, cx, Ax[m], b[m])
        assert rc == 0 and st == 0
        assert relerr(x, ref[m]) < 1e-10
        rc, xt, st, g, _ = selftest(h, cx, Ax[m], b[m], transpose=1)
        assert rc == 0 and st == 0
        assert relerr(xt, reft[m]) < 1e-10
    _cabi.free_symbolic(h)


def test_btf_blocks_and_offdiagonal_entries(oracle):
    """Block triangular form with off-diagonal blocks (klu_solve's block back substitution)."""
    pat = synth.MnaPattern(n=150, seed=3)
    n = pat.n
    Ax = pat.values(0, 2)
    h = _cabi.analyze(pat.Ai, pat.Aj, n)
    _cabi.seed_host(h, Ax[0])
    info = _cabi.symbolic_info(h, False)
    assert info["nblocks"] > 1 and info["nzoff"] > 0
    rng = np.random.default_rng(2)
    b = rng.standard_normal((n, 2))
    A = dense(pat.Ai, pat.Aj, Ax[1], n)
    rc, x, st, g, stats = selftest(h, False, Ax[1], b, env=None)
    if rc == -4:
        pytest.skip("pattern needs more value registers than the kernel has")
    assert rc == 0 and st == 0
    assert relerr(x, np.linalg.solve(A, b)) < 1e-10
    rc, xt, st, g, stats = selftest(h, False, Ax[1], b, transpose=1)
    assert rc == 0 and st == 0
    assert relerr(xt, np.linalg.solve(A.T, b)) < 1e-10
    _cabi.free_symbolic(h)


def test_singular_pivot_is_reported():
    n = 6
    A = np.eye(n) * 3.0
    A[0, 1] = A[1, 0] = 1.0
    A[2, 3] = A[3, 2] = 1.0
    Ai, Aj = np.nonzero(A)
    Ai, Aj = Ai.astype(np.int32), Aj.astype(np.int32)
    Ax = A[Ai, Aj]
    h = _cabi.analyze(Ai, Aj, n)
    _cabi.seed_host(h, Ax)
    bad = Ax.copy()
    bad[[k for k, (i, j) in enumerate(zip(Ai, Aj)) if i == 4 and j == 4][0]] = 0.0
    rc, x, st, g, _ = selftest(h, False, bad, np.ones((n, 1)))
    assert rc == 0 and st == 1
    _cabi.free_symbolic(h)
