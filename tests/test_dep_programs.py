"""Dependency-driven regime (plan.hpp: ColProgram, PanelProgram; kernels_dep.cuh) and the row-blocked
multi-right-hand-side solve (plan.hpp: GroupProgram; kernels_level.cuh: lv_solve_grouped).

CPU part: tools/dep_check.cpp interprets both programs the way the kernels do on one system of
several patterns (BTF blocks with off-diagonal entries, chains, dense fill, rows long enough to
be split into partial sums) and checks the order property of the column program and the backward
error of x.  GPU part: the panel solve and the opt-in column factor against the oracle
(restated klu_factor + klu_solve / klu_tsolve per system, /root/reference/klujax.cpp:431-451,
1081-1092, 1213-1224) at 1e-10, both value types, several right-hand sides."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "klujax_b200", "csrc")
X_TOL = 1e-10


def _patterns():
    from klujax_b200 import synth
    rng = np.random.default_rng(5)
    out = {}
    p = synth.SaxPattern(64, seed=2)
    out["sax64"] = (p.Ai, p.Aj, p.values(0, 1)[0])
    Ai, Aj, Ax = synth.random_dd(n=300, n_off=1500, seed=1)
    out["dd300"] = (Ai, Aj, np.asarray(Ax).reshape(-1)[:len(Ai)])
    m = synth.MnaPattern(n=1500, seed=4)
    out["mna1500"] = (m.Ai, m.Aj, m.values(0, 1)[0])
    m2 = synth.MnaPattern(n=800, seed=9, symmetric_hide=False)  # independent permutations: real BTF work
    out["mna800i"] = (m2.Ai, m2.Aj, m2.values(0, 1)[0])
    n = 700                                                     # arrow + chain: one very long row and column
    rows = [np.arange(n), np.arange(1, n), np.full(n - 1, n - 1), np.arange(n - 1)]
    cols = [np.arange(n), np.arange(n - 1), np.arange(n - 1), np.full(n - 1, n - 1)]
    Ai = np.concatenate(rows); Aj = np.concatenate(cols)
    key = np.unique(Ai.astype(np.int64) * n + Aj)
    Ai, Aj = (key // n).astype(np.int32), (key % n).astype(np.int32)
    Ax = rng.uniform(-1, 1, Ai.size) + 0j
    Ax[Ai == Aj] = 6.0
    out["arrow700"] = (Ai, Aj, Ax)
    return out


@pytest.fixture(scope="module")
def dep_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("dep") / "dep_check")
    src = [os.path.join(ROOT, "tools", "dep_check.cpp")] + [os.path.join(CSRC, f) for f in ("ordering.cpp", "seed_lu.cpp", "plan.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-I", CSRC, "-o", exe] + src)
    return exe


@pytest.mark.parametrize("name", ["sax64", "dd300", "mna1500", "mna800i", "arrow700"])
def test_programs_on_the_host(dep_check, tmp_path, name):
    Ai, Aj, Ax = _patterns()[name]
    n = int(max(Ai.max(), Aj.max())) + 1
    f = tmp_path / (name + ".txt")
    with open(f, "w") as fh:
        fh.write(f"{n} {len(Ai)}\n")
        for i, j, v in zip(Ai, Aj, Ax):
            fh.write(f"{int(i)} {int(j)} {float(np.real(v))!r} {float(np.imag(v))!r}\n")
    res = subprocess.run([dep_check, str(f)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "order violations 0" in res.stdout and res.stdout.strip().endswith("OK"), res.stdout
    assert res.stdout.count("group program") == 2 and "level snapshot" in res.stdout, res.stdout


# --------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def K():
    import torch
    assert torch.cuda.is_available()
    import klujax_b200
    from klujax_b200 import _cabi
    _cabi.lib()
    return klujax_b200


def _relerr(x, ref):
    return float(np.abs(x - ref).max() / np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name,cplx,r", [("mna1500", False, 1), ("mna1500", True, 3), ("dd300", True, 2), ("arrow700", True, 1),
                                         ("sax64", True, 4)])
def test_panel_solve_and_column_factor_vs_oracle(K, oracle, monkeypatch, name, cplx, r):
    import torch
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", "1")
    Ai, Aj, Ax0 = _patterns()[name]
    n = int(max(Ai.max(), Aj.max())) + 1
    rng = np.random.default_rng(11)
    L = 5
    scale = 1.0 + 0.05 * rng.standard_normal((L, Ax0.size))
    Ax = (np.asarray(Ax0)[None, :] * scale).astype(np.complex128 if cplx else np.float64)
    if cplx:
        Ax = Ax * np.exp(1j * 0.1 * rng.standard_normal((L, 1)))
    b = rng.standard_normal((L, n, r)) + (1j * rng.standard_normal((L, n, r)) if cplx else 0)
    b = b.astype(Ax.dtype)
    so = oracle.analyze(Ai, Aj, n)
    ref = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
    tref = oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)
    tAi, tAj = torch.from_numpy(np.asarray(Ai, np.int32)).cuda(), torch.from_numpy(np.asarray(Aj, np.int32)).cuda()
    tAx, tb = torch.from_numpy(Ax).cuda(), torch.from_numpy(b).cuda()
    for dep, dep_factor in (("0", "0"), ("1", "0"), ("1", "1")):
        monkeypatch.setenv("KLUJAX_CUDA_DEP", dep)
        monkeypatch.setenv("KLUJAX_CUDA_DEP_FACTOR", dep_factor)
        sym = K.analyze(Ai, Aj, n)
        x = K.solve_with_symbol(tAi, tAj, tAx, tb, sym).cpu().numpy()
        assert _relerr(x, ref) < X_TOL, (dep, dep_factor)
        xt = K.tsolve_with_symbol(tAi, tAj, tAx, tb, sym).cpu().numpy()
        assert _relerr(xt, tref) < X_TOL, (dep, dep_factor)
        num = K.factor(tAi, tAj, tAx, sym)
        x2 = K.solve_with_numeric(num, tb, sym).cpu().numpy()
        xt2 = K.tsolve_with_numeric(num, tb, sym).cpu().numpy()
        assert _relerr(x2, ref) < X_TOL and _relerr(xt2, tref) < X_TOL, (dep, dep_factor)
        K.free_numeric(num)


@pytest.mark.gpu
def test_panel_solve_reports_singular_and_keeps_status(K, monkeypatch):
    """A zero pivot is KLU_SINGULAR through the dependency-driven path too (no CPU fallback)."""
    import torch
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", "1")
    monkeypatch.setenv("KLUJAX_CUDA_DEP_FACTOR", "1")
    Ai, Aj, Ax0 = _patterns()["dd300"]
    n = int(max(Ai.max(), Aj.max())) + 1
    Ax = np.tile(np.asarray(Ax0, np.float64), (3, 1))
    Ax[1, :] = 0.0
    b = np.ones((3, n, 1))
    sym = K.analyze(Ai, Aj, n)
    t = lambda a: torch.from_numpy(a).cuda()
    with pytest.raises(Exception):
        K.solve_with_symbol(t(np.asarray(Ai, np.int32)), t(np.asarray(Aj, np.int32)), t(Ax), t(b), sym)


@pytest.mark.gpu
@pytest.mark.parametrize("cplx,g,L,r", [(False, 36, 8, 256), (True, 36, 8, 256), (False, 20, 48, 48), (True, 16, 64, 33),
                                        (False, 16, 200, 16)])  # the last one: more systems than SMs (batch-interleaved values)
def test_row_blocked_multi_rhs_solve_vs_oracle(K, oracle, monkeypatch, cplx, g, L, r):
    """lv_solve_grouped (plan.hpp: GroupProgram): many right-hand sides per system, rows of a
    supernode blocked in registers; solve and transpose solve against the oracle and against the plain
    level kernel."""
    import torch
    from klujax_b200 import synth
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", "1")
    # r = 48 and r = 33: the last warp of a system has idle lanes (and one column per lane for odd r)
    Ai, Aj, Ax = synth.poisson2d(g, L)
    n = g * g
    rng = np.random.default_rng(3)
    Ax = Ax * (1.0 + 0.05 * rng.standard_normal(Ax.shape))
    if cplx:
        Ax = Ax * np.exp(1j * 0.2 * rng.standard_normal((L, 1)))
    b = rng.standard_normal((L, n, r)) + (1j * rng.standard_normal((L, n, r)) if cplx else 0)
    b = b.astype(Ax.dtype)
    so = oracle.analyze(Ai, Aj, n)
    ns = 2                                   # the oracle on two systems is enough (seconds)
    ref = oracle.solve_with_symbol(Ai, Aj, Ax[:ns], b[:ns], so)
    tref = oracle.solve_with_symbol(Ai, Aj, Ax[:ns], b[:ns], so, transpose=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tAi, tAj, tAx, tb = t(np.asarray(Ai, np.int32)), t(np.asarray(Aj, np.int32)), t(Ax), t(b)
    out = {}
    for grouped in ("1", "0"):
        monkeypatch.setenv("KLUJAX_CUDA_GROUPED", grouped)
        sym = K.analyze(Ai, Aj, n)
        num = K.factor(tAi, tAj, tAx, sym)
        x = K.solve_with_numeric(num, tb, sym).cpu().numpy()
        xt = K.tsolve_with_numeric(num, tb, sym).cpu().numpy()
        xs = K.solve_with_symbol(tAi, tAj, tAx, tb, sym).cpu().numpy()
        assert _relerr(x[:ns], ref) < X_TOL and _relerr(xt[:ns], tref) < X_TOL and _relerr(xs[:ns], ref) < X_TOL, grouped
        out[grouped] = (x, xt)
        K.free_numeric(num)
    assert _relerr(out["1"][0], out["0"][0]) < X_TOL and _relerr(out["1"][1], out["0"][1]) < X_TOL


@pytest.mark.gpu
def test_row_blocked_solve_on_fuzzed_structures(K, monkeypatch):
    """lv_solve_grouped and panel_solve over structures with hidden diagonals, arrow heads, BTF chains with
    off-diagonal blocks, dense blocks and bands (seeded), 32 right-hand sides per system: x and the
    transpose solve against dense LAPACK."""
    import torch
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", "1")
    rng = np.random.default_rng(4321)

    def pattern(kind, n):
        A = np.zeros((n, n))
        if kind == "random":
            A = rng.standard_normal((n, n)) * (rng.random((n, n)) < min(0.5, 4.0 / n))
        elif kind == "arrow":
            A[0, :] = rng.standard_normal(n); A[:, 0] = rng.standard_normal(n); A[-1, :] = rng.standard_normal(n)
        elif kind == "btf":
            k = 0
            while k < n:
                m = int(rng.integers(1, 9)); e = min(n, k + m)
                A[k:e, k:e] = rng.standard_normal((e - k, e - k))
                if e < n:
                    A[k:e, e:] = rng.standard_normal((e - k, n - e)) * (rng.random((e - k, n - e)) < 0.1)
                k = e
        elif kind == "dense":
            A = rng.standard_normal((n, n))
        elif kind == "band":
            for d in range(-3, 4):
                A += np.diag(rng.standard_normal(n - abs(d)), d)
        A[np.arange(n), np.arange(n)] = 0.0
        A += np.diag(2.0 + np.abs(A).sum(1))
        return A[rng.permutation(n)]

    for kind, n, cplx in (("random", 90, False), ("arrow", 48, True), ("btf", 130, False), ("dense", 40, True), ("band", 120, False)):
        A = pattern(kind, n)
        Ai, Aj = (v.astype(np.int32) for v in np.nonzero(A))
        for L, r in ((64, 32), (3, 2)):   # row-blocked solve / panel solve
            Ax = np.tile(A[Ai, Aj], (L, 1)) * (1.0 + 1e-4 * rng.standard_normal((L, Ai.size)))  # same pivots for every system
            b = rng.standard_normal((L, n, r))
            if cplx:
                Ax = Ax * (1.0 + 0.1j)
                b = b + 1j * rng.standard_normal((L, n, r))
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            res = {}
            for new_kernels in ("1", "0"):  # default kernels, then the plain level kernels on the same factors
                monkeypatch.setenv("KLUJAX_CUDA_DEP", new_kernels)
                monkeypatch.setenv("KLUJAX_CUDA_GROUPED", new_kernels)
                sym = K.analyze(Ai, Aj, n)
                num = K.factor(t(Ai), t(Aj), t(Ax), sym)
                res[new_kernels] = (K.solve_with_numeric(num, t(b), sym).cpu().numpy(),
                                    K.tsolve_with_numeric(num, t(b), sym).cpu().numpy())
                K.free_numeric(num)
            x, xt = res["1"]
            assert _relerr(x, res["0"][0]) < 1e-9 and _relerr(xt, res["0"][1]) < 1e-9, (kind, n, L, r)
            # the transversal need not recover the dominant diagonal of a row-permuted matrix, so KLU's
            # tol = 1e-3 pivoting may run on small pivots: a loose bar against dense LAPACK
            for m in range(0, L, max(1, L // 4)):
                D = np.zeros((n, n), dtype=Ax.dtype)
                D[Ai, Aj] = Ax[m]
                assert _relerr(x[m], np.linalg.solve(D, b[m])) < 1e-6, (kind, n, L, r, m)
                assert _relerr(xt[m], np.linalg.solve(D.T, b[m])) < 1e-6, (kind, n, L, r, m)
