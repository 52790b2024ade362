"""Pins the CPU oracle (oracle/) before anything is compared against it:
the reference's published known answers, dense LAPACK and SuperLU on the
reference's own test shapes (/root/reference/tests.py:299-358), and the
size-independent properties of the path.  No GPU."""
import numpy as np
import pytest

from klujax_b200 import synth
from util import dense, golden, relerr, residual

DTYPES = [np.float64, np.complex128]


def test_readme_known_answer(oracle):
    g = golden()["readme_5x5"]
    A = np.array(g["A_dense"], float)
    Ai, Aj = np.nonzero(A)
    x = oracle.solve(Ai, Aj, A[Ai, Aj][None], np.array(g["b"], float)[None, :, None])
    assert np.abs(x.ravel() - np.array(g["x"])).max() < 1e-12
    # zeros on the diagonal: needs the maximum transversal; BTF blocks [1, 3, 1]
    sym = oracle.analyze(Ai, Aj, 5)
    info = oracle.symbolic_info(sym)
    assert info["nblocks"] == 3 and sorted(np.diff(info["R"])) == [1, 1, 3]
    oracle.free_symbolic(sym)


def test_docs_batched_known_answers(oracle):
    g = golden()
    k = g["batched_multi_rhs"]
    x = oracle.solve(k["Ai"], k["Aj"], np.array(k["Ax"])[None], np.array(k["b"])[None])
    assert np.allclose(x[0], np.array(k["x"]), atol=1e-14)
    k = g["batched_multi_matrix"]
    x = oracle.solve(k["Ai"], k["Aj"], np.array(k["Ax"]), np.array(k["b"])[:, :, None])
    assert np.allclose(x[:, :, 0], np.array(k["x"]), atol=1e-14)
    k = g["batched_broadcast_b"]
    b = np.broadcast_to(np.array(k["b"])[None, :, None], (2, 3, 1))
    x = oracle.solve(k["Ai"], k["Aj"], np.array(k["Ax"]), b)
    assert np.allclose(x[:, :, 0], np.array(k["x"]), atol=1e-14)


def test_golden_oracle_vectors(oracle):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) == 4
    for nm in names:
        Ai, Aj, Ax, b = (z[f"{nm}/{k}"] for k in ("Ai", "Aj", "Ax", "b"))
        n = b.shape[1]
        sym = oracle.analyze(Ai, Aj, n)
        assert relerr(oracle.solve_with_symbol(Ai, Aj, Ax, b, sym), z[f"{nm}/x"]) < 1e-14
        assert relerr(oracle.solve_with_symbol(Ai, Aj, Ax, b, sym, transpose=True), z[f"{nm}/xt"]) < 1e-14
        oracle.free_symbolic(sym)
        for m in range(b.shape[0]):  # and the stored answers are right
            A = dense(Ai, Aj, Ax[m], n)
            assert relerr(z[f"{nm}/x"][m], np.linalg.solve(A, b[m])) < 1e-12
            assert relerr(z[f"{nm}/xt"][m], np.linalg.solve(A.T, b[m])) < 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(3, 8, 1, 1), (5, 15, 1, 2), (5, 15, 4, 1), (5, 15, 3, 2), (40, 160, 2, 3)])
def test_all_entry_points_vs_dense(oracle, dtype, shape):
    n, nz, L, r = shape
    Ai, Aj, Ax, b = synth.ref_style_case(n, nz, L, r, dtype)
    x = oracle.solve(Ai, Aj, Ax, b)
    sym = oracle.analyze(Ai, Aj, n)
    xs = oracle.solve_with_symbol(Ai, Aj, Ax, b, sym)
    xt = oracle.solve_with_symbol(Ai, Aj, Ax, b, sym, transpose=True)
    num = oracle.factor(Ai, Aj, Ax, sym)
    xn = oracle.solve_with_numeric(sym, num, b)
    xnt = oracle.solve_with_numeric(sym, num, b, transpose=True)
    Ax2 = Ax * 2.0 + 0.5 * Ax.conj()
    x2, num2 = oracle.refactor_and_solve(Ai, Aj, Ax2, b, sym, num)
    assert np.array_equal(num2, num)  # same handles (klujax.cpp:922-923)
    num3 = oracle.refactor(Ai, Aj, Ax, sym, num)
    x3 = oracle.solve_with_numeric(sym, num3, b)
    for m in range(L):
        A = dense(Ai, Aj, Ax[m], n)
        ref = np.linalg.solve(A, b[m])
        for got in (x[m], xs[m], xn[m], x3[m]):
            assert relerr(got, ref) < 1e-12
        reft = np.linalg.solve(A.T, b[m])
        for got in (xt[m], xnt[m]):
            assert relerr(got, reft) < 1e-12
        assert relerr(x2[m], np.linalg.solve(dense(Ai, Aj, Ax2[m], n), b[m])) < 1e-12
    # numeric broadcast: one factorization against a batch of b (klujax.cpp:1044-1059)
    xb = oracle.solve_with_numeric(sym, num[:1], b)
    A0 = dense(Ai, Aj, Ax[0], n)
    for m in range(L):
        assert relerr(xb[m], np.linalg.solve(A0, b[m])) < 1e-12
    oracle.free_numeric(num)
    oracle.free_symbolic(sym)


def test_singular_is_reported(oracle):
    Ai = np.array([0, 1, 0, 1], np.int32)
    Aj = np.array([0, 0, 1, 1], np.int32)
    Ax = np.array([[1.0, 2.0, 2.0, 4.0]])  # rank 1
    with pytest.raises(oracle.OracleError):
        oracle.solve(Ai, Aj, Ax, np.ones((1, 2, 1)))
    # structurally singular: empty column
    with pytest.raises(oracle.OracleError):
        oracle.solve(np.array([0, 1], np.int32), np.array([0, 0], np.int32), np.ones((1, 2)), np.ones((1, 2, 1)))


def test_duplicates_rejected_like_klu_scale(oracle):
    Ai = np.array([0, 0, 1], np.int32)
    Aj = np.array([0, 0, 1], np.int32)
    with pytest.raises(oracle.OracleError):
        oracle.solve(Ai, Aj, np.ones((1, 3)), np.ones((1, 2, 1)))


def test_coo_to_csc_is_the_stable_bucket_sort(oracle):
    r = np.random.default_rng(0)
    n, nz = 7, 40
    Ai, Aj = r.integers(0, n, nz), r.integers(0, n, nz)
    Bp, Bi, Bk = oracle.coo_to_csc(n, Ai, Aj)
    order = np.argsort(Aj, kind="stable")
    assert np.array_equal(Bk, order) and np.array_equal(Bi, Ai[order])
    assert np.array_equal(Bp, np.concatenate([[0], np.cumsum(np.bincount(Aj, minlength=n))]))


def test_btf_partition_matches_scipy(oracle):
    """Block *sets* are unique (strong components of the matched graph)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components, maximum_bipartite_matching
    pat = synth.MnaPattern(n=600, seed=5, symmetric_hide=False)
    n = pat.n
    sym = oracle.analyze(pat.Ai, pat.Aj, n)
    info = oracle.symbolic_info(sym)
    oracle.free_symbolic(sym)
    A = sp.csr_matrix((np.ones(pat.nnz), (pat.Ai, pat.Aj)), shape=(n, n))
    match = maximum_bipartite_matching(A, perm_type="column")  # match[i] = column of row i
    assert (match >= 0).all() and info["structural_rank"] == n
    G = A[:, match]  # zero-free diagonal
    ncomp, lab = connected_components(G, directed=True, connection="strong")
    assert ncomp == info["nblocks"]
    assert sorted(np.bincount(lab).tolist()) == sorted(np.diff(info["R"]).tolist())
    assert sorted(info["P"].tolist()) == list(range(n)) and sorted(info["Q"].tolist()) == list(range(n))


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg5", "cfg4_small", "cfg3_small"])
def test_configs_against_superlu(oracle, cfg):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    r = np.random.default_rng(7)
    if cfg == "cfg1":
        Ai, Aj, Ax = synth.random_dd()
        n = 1000
    elif cfg in ("cfg2", "cfg5"):
        n = 64 if cfg == "cfg2" else 128
        pat = synth.SaxPattern(n, seed=1 if cfg == "cfg2" else 2)
        Ai, Aj, Ax = pat.Ai, pat.Aj, pat.values(0, 2)
    elif cfg == "cfg4_small":
        pat = synth.MnaPattern(n=2000)
        Ai, Aj, Ax, n = pat.Ai, pat.Aj, pat.values(0, 2), pat.n
    else:
        Ai, Aj, Ax = synth.poisson2d(48, 2)
        n = 48 * 48
    L = Ax.shape[0]
    b = r.standard_normal((L, n, 2)).astype(Ax.dtype)
    sym = oracle.analyze(Ai, Aj, n)
    x = oracle.solve_with_symbol(Ai, Aj, Ax, b, sym)
    xt = oracle.solve_with_symbol(Ai, Aj, Ax, b, sym, transpose=True)
    oracle.free_symbolic(sym)
    for m in range(L):
        A = sp.csc_matrix((Ax[m], (Ai, Aj)), shape=(n, n))
        lu = spl.splu(A)
        assert relerr(x[m], lu.solve(b[m])) < 1e-9
        assert relerr(xt[m], lu.solve(b[m], trans="T")) < 1e-9
        assert residual(Ai, Aj, Ax[m], x[m], b[m], n) < 1e-12
        assert residual(Ai, Aj, Ax[m], xt[m], b[m], n, transpose=True) < 1e-12


def test_linearity_property(oracle):
    """solve is linear in b: A^-1(2 b1 + b2) = 2 x1 + x2 (size-independent check)."""
    pat = synth.SaxPattern(64, seed=1)
    Ax = pat.values(3, 1)
    r = np.random.default_rng(1)
    b1 = r.standard_normal((1, 64, 1)) + 1j * r.standard_normal((1, 64, 1))
    b2 = r.standard_normal((1, 64, 1)) + 0j
    sym = oracle.analyze(pat.Ai, pat.Aj, 64)
    x1 = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b1, sym)
    x2 = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b2, sym)
    x12 = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, 2 * b1 + b2, sym)
    oracle.free_symbolic(sym)
    assert relerr(x12, 2 * x1 + x2) < 1e-13


@pytest.mark.parametrize("dtype", DTYPES)
def test_triangular_all_singleton_blocks(oracle, dtype):
    """Upper-triangular A: n singleton BTF blocks, only off-diagonal back-substitution."""
    r = np.random.default_rng(3)
    n, L = 12, 3
    iu, ju = np.triu_indices(n)
    keep = (iu == ju) | (r.random(iu.size) < 0.5)
    Ai, Aj = iu[keep].astype(np.int32), ju[keep].astype(np.int32)
    Ax = r.standard_normal((L, Ai.size)).astype(dtype)
    Ax[:, Ai == Aj] += 4.0
    b = r.standard_normal((L, n, 2)).astype(dtype)
    sym = oracle.analyze(Ai, Aj, n)
    info = oracle.symbolic_info(sym)
    assert info["nblocks"] == n and info["nzoff"] == Ai.size - n
    x = oracle.solve_with_symbol(Ai, Aj, Ax, b, sym)
    xt = oracle.solve_with_symbol(Ai, Aj, Ax, b, sym, transpose=True)
    oracle.free_symbolic(sym)
    for m in range(L):
        A = dense(Ai, Aj, Ax[m], n)
        assert relerr(x[m], np.linalg.solve(A, b[m])) < 1e-12
        assert relerr(xt[m], np.linalg.solve(A.T, b[m])) < 1e-12


def test_cfg1_cpu_latency_reference(oracle, capsys):
    """Config 1 on the CPU side (the figure DESIGN.md quotes next to tools/time_cfg1.py): the
    restated KLU's analyze + factor + solve of the n = 1000 system; checked against LAPACK."""
    import time
    from klujax_b200 import synth
    Ai, Aj, Ax = synth.random_dd()
    n = 1000
    b = np.random.default_rng(1).standard_normal((1, n, 1))
    t0 = time.perf_counter()
    x = oracle.solve(Ai, Aj, Ax, b)
    dt = time.perf_counter() - t0
    A = synth.dense_from_coo(Ai, Aj, Ax[0], n)
    assert np.linalg.norm(A @ x[0] - b[0]) / np.linalg.norm(b[0]) < 1e-12
    with capsys.disabled():
        print(f"\n[cfg1 CPU] analyze + factor + solve, one thread: {dt * 1e3:.1f} ms")
