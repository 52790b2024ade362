"""Parity tests proper: the CUDA path, called through the C-ABI by the API
mirror (klujax_b200.api), against the CPU oracle on the same seeded inputs.
Tolerances are BASELINE.json's: x within 1e-10 relative, residual
||Ax-b||/||b|| <= 1e-12 on well-conditioned systems (float64 and complex128).
The cases follow /root/reference/tests.py: shape matrix (:45-170), split API
(:174-294), refactor (:401-527), transpose solves (:555-625),
refactor_and_solve (:630-675)."""
import os

import numpy as np
import pytest

from klujax_b200 import synth
from util import dense, golden, relerr, residual

pytestmark = pytest.mark.gpu

X_TOL = 1e-10
RES_TOL = 1e-12
DTYPES = [np.float64, np.complex128]


@pytest.fixture(scope="module")
def K():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import klujax_b200
    from klujax_b200 import _cabi
    _cabi.lib()  # fail loudly if the CUDA library is missing
    return klujax_b200


def np_(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ known answers
def test_readme_known_answer(K):
    g = golden()["readme_5x5"]
    A = np.array(g["A_dense"], float)
    Ai, Aj = np.nonzero(A)
    x = np_(K.solve(Ai, Aj, A[Ai, Aj], np.array(g["b"], float)))
    assert x.shape == (5,)
    assert np.abs(x - np.array(g["x"])).max() < 1e-12


def test_docs_batched_known_answers(K):
    g = golden()
    k = g["batched_multi_rhs"]
    assert np.allclose(np_(K.solve(k["Ai"], k["Aj"], np.array(k["Ax"]), np.array(k["b"]))), k["x"], atol=1e-14)
    k = g["batched_multi_matrix"]
    assert np.allclose(np_(K.solve(k["Ai"], k["Aj"], np.array(k["Ax"]), np.array(k["b"]))), k["x"], atol=1e-14)
    k = g["batched_broadcast_b"]
    assert np.allclose(np_(K.solve(k["Ai"], k["Aj"], np.array(k["Ax"]), np.array(k["b"]))), k["x"], atol=1e-14)


def test_golden_oracle_vectors(K):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
    for nm in sorted({k.split("/")[0] for k in z.files}):
        Ai, Aj, Ax, b = (z[f"{nm}/{k}"] for k in ("Ai", "Aj", "Ax", "b"))
        with K.analyze(Ai, Aj, b.shape[1]) as sym:
            assert relerr(np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym)), z[f"{nm}/x"]) < X_TOL
            assert relerr(np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym)), z[f"{nm}/xt"]) < X_TOL


# ------------------------------------------------------------------ shape matrix
@pytest.mark.parametrize("dtype", DTYPES)
def test_shapes_1d_2d_3d(K, oracle, dtype):
    # (n_nz,) x (n_col,)
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 1, 1, dtype)
    x = np_(K.solve(Ai, Aj, Ax[0], b[0, :, 0]))
    assert x.shape == (5,)
    assert relerr(x, oracle.solve(Ai, Aj, Ax, b)[0, :, 0]) < X_TOL
    # (n_nz,) x (n_col, n_rhs)
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 1, 3, dtype)
    x = np_(K.solve(Ai, Aj, Ax[0], b[0]))
    assert x.shape == (5, 3) and relerr(x, oracle.solve(Ai, Aj, Ax, b)[0]) < X_TOL
    # (n_lhs, n_nz) x (n_lhs, n_col)
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 4, 1, dtype)
    x = np_(K.solve(Ai, Aj, Ax, b[:, :, 0]))
    assert x.shape == (4, 5) and relerr(x, oracle.solve(Ai, Aj, Ax, b)[:, :, 0]) < X_TOL
    # (n_lhs, n_nz) x (n_lhs, n_col, n_rhs)
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 3, 2, dtype)
    x = np_(K.solve(Ai, Aj, Ax, b))
    assert x.shape == (3, 5, 2) and relerr(x, oracle.solve(Ai, Aj, Ax, b)) < X_TOL
    # (n_nz,) x (n_lhs, n_col, n_rhs): one matrix broadcast over the batch
    x = np_(K.solve(Ai, Aj, Ax[0], b))
    ref = oracle.solve(Ai, Aj, np.broadcast_to(Ax[:1], Ax.shape), b)
    assert x.shape == (3, 5, 2) and relerr(x, ref) < X_TOL
    # dense check, like tests.py:49-51
    for m in range(3):
        assert relerr(ref[m], np.linalg.solve(dense(Ai, Aj, Ax[0], 5), b[m])) < 1e-12


def test_4d_raises_value_error(K):
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 2, 2, np.float64)
    with pytest.raises(ValueError):
        K.solve(Ai, Aj, Ax, b[None])
    with pytest.raises(ValueError):
        K.dot(Ai, Aj, Ax[None], b)


# ------------------------------------------------------------------ split API
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(3, 8, 1, 1), (5, 15, 3, 2), (5, 15, 4, 6), (40, 160, 5, 3)])
def test_every_entry_point_vs_oracle(K, oracle, dtype, shape):
    n, nz, L, r = shape
    Ai, Aj, Ax, b = synth.ref_style_case(n, nz, L, r, dtype)
    so = oracle.analyze(Ai, Aj, n)
    sym = K.analyze(Ai, Aj, n)
    # the numeric call may present the COO entries in another order (klujax.py:319-321)
    perm = np.random.default_rng(1).permutation(Ai.size)
    xo = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
    xto = oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)
    assert relerr(np_(K.solve_with_symbol(Ai[perm], Aj[perm], Ax[:, perm], b, sym)), xo) < X_TOL
    assert relerr(np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym)), xto) < X_TOL
    st, gr = K.last_status()
    assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
    num = K.factor(Ai, Aj, Ax, sym)
    assert num.handle.shape == (L,) and num.handle.dtype == np.uint64 and (num.handle != 0).all()
    assert relerr(np_(K.solve_with_numeric(num, b, sym)), xo) < X_TOL
    assert relerr(np_(K.tsolve_with_numeric(num, b, sym)), xto) < X_TOL
    # broadcast one factorization over the batch of b, and one b over the batch
    no = oracle.factor(Ai, Aj, Ax, so)
    one = K.KLUHandleManager(num.handle[:1], None, owner=False)
    assert relerr(np_(K.solve_with_numeric(one, b, sym)), oracle.solve_with_numeric(so, no[:1], b)) < X_TOL
    assert relerr(np_(K.solve_with_numeric(num, b[:1], sym)), oracle.solve_with_numeric(so, no, b[:1])) < X_TOL
    # a reversed subset of the handles
    if L > 1:
        rev = K.KLUHandleManager(num.handle[::-1].copy(), None, owner=False)
        assert relerr(np_(K.solve_with_numeric(rev, b, sym)), oracle.solve_with_numeric(so, no[::-1].copy(), b)) < X_TOL
    # refactor with new values, in place, same handles back
    Ax2 = Ax * 1.5 + 0.25 * np.roll(Ax, 1, axis=1)
    num2 = K.refactor(Ai, Aj, Ax2, num, sym)
    assert num2._owner is False and np.array_equal(num2.handle, num.handle)
    x2o = oracle.solve_with_symbol(Ai, Aj, Ax2, b, so)
    assert relerr(np_(K.solve_with_numeric(num2, b, sym)), x2o) < X_TOL
    # fused refactor + solve
    x3, num3 = K.refactor_and_solve(Ai, Aj, Ax, b, num, sym)
    assert num3._owner is False and np.array_equal(num3.handle, num.handle)
    assert relerr(np_(x3), xo) < X_TOL
    assert relerr(np_(K.tsolve_with_numeric(num3, b, sym)), xto) < X_TOL
    for m in range(L):
        assert residual(Ai, Aj, Ax[m], np_(x3)[m], b[m], n) < RES_TOL
    K.free_numeric(num)
    K.free_symbolic(sym)
    oracle.free_numeric(no)
    oracle.free_symbolic(so)


def test_solve_with_numeric_b_ranks(K, oracle):
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 3, 2, np.float64)
    so = oracle.analyze(Ai, Aj, 5)
    no = oracle.factor(Ai, Aj, Ax, so)
    with K.analyze(Ai, Aj, 5) as sym, K.factor(Ai, Aj, Ax, sym) as num:
        one = K.KLUHandleManager(num.handle[:1], None, owner=False)
        # 1-D b with one numeric -> (n_col,)
        x = np_(K.solve_with_numeric(one, b[0, :, 0], sym))
        assert x.shape == (5,) and relerr(x, oracle.solve_with_numeric(so, no[:1], b[:1, :, :1])[0, :, 0]) < X_TOL
        # 2-D b, first dim == number of numerics -> (n_lhs, n_col)   (klujax.cpp:1023)
        x = np_(K.solve_with_numeric(num, b[:, :, 0], sym))
        assert x.shape == (3, 5) and relerr(x, oracle.solve_with_numeric(so, no, b[:, :, :1])[:, :, 0]) < X_TOL
        # 2-D b with a single numeric -> (n_col, n_rhs)
        x = np_(K.solve_with_numeric(one, b[0], sym))
        assert x.shape == (5, 2) and relerr(x, oracle.solve_with_numeric(so, no[:1], b[:1])[0]) < X_TOL
    oracle.free_numeric(no)
    oracle.free_symbolic(so)


# ------------------------------------------------------------------ status path
@pytest.mark.parametrize("regime", ["0", "1"])
def test_singular_system_is_flagged_per_system(K, oracle, regime, monkeypatch):
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", regime)
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 4, 1, np.float64)
    Ax = Ax.copy()
    Ax[2] = 0.0  # system 2 is singular; the seed (system 0) is fine
    with K.analyze(Ai, Aj, 5) as sym:
        with pytest.raises(K.KlujaxCudaError):
            K.solve_with_symbol(Ai, Aj, Ax, b, sym)
        x = np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym, check=False))
        st, _ = K.last_status()
        assert np_(st).tolist() == [0, 0, 1, 0]
        so = oracle.analyze(Ai, Aj, 5)
        keep = [0, 1, 3]
        assert relerr(x[keep], oracle.solve_with_symbol(Ai, Aj, Ax[keep], b[keep], so)) < X_TOL
        oracle.free_symbolic(so)
        with pytest.raises(K.KlujaxCudaError):
            K.factor(Ai, Aj, Ax, sym)


def test_pattern_mismatch_is_invalid(K):
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 2, 1, np.float64)
    with K.analyze(Ai, Aj, 5) as sym:
        K.solve_with_symbol(Ai, Aj, Ax, b, sym)  # seeds
        Ai2 = Ai.copy()
        Ai2[0] = 7  # out of range
        with pytest.raises(K.KlujaxCudaError) as ei:
            K.solve_with_symbol(Ai2, Aj, Ax, b, sym)
        assert ei.value.code == -3


# ------------------------------------------------------------------ the named configs
@pytest.mark.parametrize("n,seed,r", [(64, 1, 1), (64, 1, 4), (128, 2, 1), (128, 2, 6)])
def test_sax_configs_vs_oracle(K, oracle, n, seed, r):
    """cfg2 / cfg5 shape (SAX photonic S-matrix systems, complex128)."""
    L = 96
    pat = synth.SaxPattern(n, seed=seed)
    Ax = pat.values(0, L)
    rng = np.random.default_rng(9)
    b = rng.standard_normal((L, n, r)) + 1j * rng.standard_normal((L, n, r))
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    with K.analyze(pat.Ai, pat.Aj, n) as sym:
        x = np_(K.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym))
        xt = np_(K.tsolve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym))
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
    assert relerr(x, oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)) < X_TOL
    assert relerr(xt, oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so, transpose=True)) < X_TOL
    for m in range(0, L, 7):
        assert residual(pat.Ai, pat.Aj, Ax[m], x[m], b[m], n) < RES_TOL
        assert residual(pat.Ai, pat.Aj, Ax[m], xt[m], b[m], n, transpose=True) < RES_TOL
    oracle.free_symbolic(so)


def test_cfg1_random_dd_solve(K, oracle):
    """cfg1: klujax.solve float64, n=1000, nnz~5000, n_lhs=1 (level-scheduled regime)."""
    Ai, Aj, Ax = synth.random_dd()
    b = np.random.default_rng(1).standard_normal((1000,))
    x = np_(K.solve(Ai, Aj, Ax[0], b))
    xo = oracle.solve(Ai, Aj, Ax, b[None, :, None])[0, :, 0]
    assert relerr(x, xo) < X_TOL
    assert residual(Ai, Aj, Ax[0], x[:, None], b[:, None], 1000) < RES_TOL


@pytest.mark.parametrize("dtype", DTYPES)
def test_level_regime_mna_btf(K, oracle, dtype):
    """cfg4 shape at reduced n: BTF blocks, refactor per step, solve + transpose solve."""
    pat = synth.MnaPattern(n=3000)
    n, L = pat.n, 5
    Ax0 = pat.values(0, L).astype(dtype)
    if dtype is np.complex128:
        Ax0 = Ax0 * (1 + 0.05j)
    rng = np.random.default_rng(4)
    b = rng.standard_normal((L, n, 2)).astype(dtype)
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    no = oracle.factor(pat.Ai, pat.Aj, Ax0, so)
    with K.analyze(pat.Ai, pat.Aj, n) as sym, K.factor(pat.Ai, pat.Aj, Ax0, sym) as num:
        for t in (1, 2):
            Ax = pat.values(t, L).astype(dtype)
            if dtype is np.complex128:
                Ax = Ax * (1 + 0.05j)
            num2 = K.refactor(pat.Ai, pat.Aj, Ax, num, sym)
            no = oracle.refactor(pat.Ai, pat.Aj, Ax, so, no)
            x = np_(K.solve_with_numeric(num2, b, sym))
            xt = np_(K.tsolve_with_numeric(num2, b, sym))
            assert relerr(x, oracle.solve_with_numeric(so, no, b)) < X_TOL
            assert relerr(xt, oracle.solve_with_numeric(so, no, b, transpose=True)) < X_TOL
            assert residual(pat.Ai, pat.Aj, Ax[1], x[1], b[1], n) < RES_TOL
        x3, _ = K.refactor_and_solve(pat.Ai, pat.Aj, Ax0, b, num, sym)
        assert relerr(np_(x3), oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax0, b, so)) < X_TOL
    oracle.free_numeric(no)
    oracle.free_symbolic(so)


def test_level_regime_poisson_multi_rhs(K, oracle):
    """cfg3 shape at reduced grid: factor once, solve_with_numeric with many RHS."""
    g, L, r = 48, 3, 40
    Ai, Aj, Ax = synth.poisson2d(g, L)
    n = g * g
    b = np.random.default_rng(3).standard_normal((L, n, r))
    so = oracle.analyze(Ai, Aj, n)
    no = oracle.factor(Ai, Aj, Ax, so)
    with K.analyze(Ai, Aj, n) as sym, K.factor(Ai, Aj, Ax, sym) as num:
        x = np_(K.solve_with_numeric(num, b, sym))
    assert relerr(x, oracle.solve_with_numeric(so, no, b)) < X_TOL
    assert residual(Ai, Aj, Ax[2], x[2], b[2], n) < RES_TOL
    oracle.free_numeric(no)
    oracle.free_symbolic(so)


@pytest.mark.parametrize("dtype", DTYPES)
def test_level_regime_on_small_cases(K, oracle, dtype, monkeypatch):
    """Same answers from both regimes on the reference's tiny test matrices."""
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", "1")
    Ai, Aj, Ax, b = synth.ref_style_case(20, 60, 37, 3, dtype)
    so = oracle.analyze(Ai, Aj, 20)
    with K.analyze(Ai, Aj, 20) as sym:
        assert relerr(np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym)), oracle.solve_with_symbol(Ai, Aj, Ax, b, so)) < X_TOL
        assert relerr(np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym)),
                      oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)) < X_TOL
    oracle.free_symbolic(so)


# ------------------------------------------------------------------ full-size properties
def test_cfg5_full_size_properties(K, oracle):
    """n_lhs = 65536, n_col = 128, complex128: too big for the oracle in seconds, so
    check size-independent properties: residual on every system, linearity, status."""
    import torch
    n, L = 128, 65536
    pat = synth.SaxPattern(n, seed=2)
    Ax = torch.from_numpy(pat.values_fast(L)).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    b = torch.randn(L, n, 1, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128)
    Ai, Aj = torch.from_numpy(pat.Ai).cuda(), torch.from_numpy(pat.Aj).cuda()
    with K.analyze(pat.Ai, pat.Aj, n) as sym:
        x = K.solve_with_symbol(Ai, Aj, Ax, b, sym)
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
        x2 = K.solve_with_symbol(Ai, Aj, Ax, 3.0 * b, sym)
    Ax_x = torch.zeros_like(b)
    Ax_x.index_add_(1, Ai.long(), Ax[:, :, None] * x[:, Aj.long(), :])
    res = (torch.linalg.vector_norm(Ax_x - b, dim=1) / torch.linalg.vector_norm(b, dim=1)).max()
    assert float(res) < RES_TOL
    assert float((x2 - 3.0 * x).abs().max() / x.abs().max()) < 1e-13
    # a sample of the full batch against the oracle (per-system klu_factor + klu_solve)
    idx = np.unique(np.linspace(0, L - 1, 384).astype(int))
    ti = torch.from_numpy(idx).cuda()
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, np_(Ax[ti]), np_(b[ti]), so)
    oracle.free_symbolic(so)
    assert relerr(np_(x[ti]), ref) < X_TOL


# ------------------------------------------------------------------ dot (next: f1)
@pytest.mark.parametrize("dtype", DTYPES)
def test_dot_vs_dense(K, dtype):
    Ai, Aj, Ax, x = synth.ref_style_case(5, 15, 3, 2, dtype)
    got = np_(K.dot(Ai, Aj, Ax, x))
    for m in range(3):
        assert relerr(got[m], dense(Ai, Aj, Ax[m], 5) @ x[m]) < 1e-13


@pytest.mark.parametrize("dtype", DTYPES)
def test_dot_is_deterministic_and_sums_duplicates_in_coo_order(K, dtype):
    """`dot` groups the entries by row with a stable sort and adds them in COO order (the order of the
    reference's serial loop, klujax.cpp:165-249): repeated calls agree bit for bit, duplicated and
    unsorted entries are summed, rows without entries come out as exact zeros."""
    rng = np.random.default_rng(17)
    n, nz, L, r = 300, 4000, 7, 3
    Ai = rng.integers(0, n - 5, nz).astype(np.int32)   # the last five rows have no entries
    Aj = rng.integers(0, n, nz).astype(np.int32)
    Ax = rng.standard_normal((L, nz)).astype(dtype)
    x = rng.standard_normal((L, n, r)).astype(dtype)
    if np.issubdtype(np.dtype(dtype), np.complexfloating):
        Ax = Ax + 1j * rng.standard_normal((L, nz))
        x = x + 1j * rng.standard_normal((L, n, r))
    a = np_(K.dot(Ai, Aj, Ax, x))
    b = np_(K.dot(Ai, Aj, Ax, x))
    assert np.array_equal(a, b)
    assert np.all(a[:, n - 5:, :] == 0)
    ref = np.zeros_like(x)
    for m in range(L):
        np.add.at(ref[m], Ai, Ax[m][:, None] * x[m][Aj])
    assert relerr(a, ref) < 1e-13


# ------------------------------------------------------------------ structural edge cases
def _upper_triangular(n, dtype, L, seed=3):
    r = np.random.default_rng(seed)
    iu, ju = np.triu_indices(n)
    keep = (iu == ju) | (r.random(iu.size) < 0.5)
    Ai, Aj = iu[keep].astype(np.int32), ju[keep].astype(np.int32)
    Ax = r.standard_normal((L, Ai.size)).astype(dtype)
    if np.issubdtype(np.dtype(dtype), np.complexfloating):
        Ax = Ax + 1j * r.standard_normal((L, Ai.size))
    Ax[:, Ai == Aj] += 4.0
    return Ai, Aj, Ax


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("regime", ["0", "1"])
def test_all_singleton_blocks_offdiagonal_backsubstitution(K, oracle, dtype, regime, monkeypatch):
    """Upper-triangular A: every BTF block is a singleton, all the work is the block
    back-substitution through the off-diagonal entries (klu_solve / klu_tsolve)."""
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", regime)
    n, L = 12, 5
    Ai, Aj, Ax = _upper_triangular(n, dtype, L)
    b = np.random.default_rng(0).standard_normal((L, n, 3)).astype(dtype)
    so = oracle.analyze(Ai, Aj, n)
    assert oracle.symbolic_info(so)["nblocks"] == n
    with K.analyze(Ai, Aj, n) as sym:
        assert relerr(np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym)), oracle.solve_with_symbol(Ai, Aj, Ax, b, so)) < X_TOL
        assert relerr(np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym)),
                      oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)) < X_TOL
    oracle.free_symbolic(so)


def test_one_by_one_and_tiny_blocks(K, oracle):
    # 1x1
    x = np_(K.solve(np.array([0]), np.array([0]), np.array([[2.0], [4.0], [-8.0]]), np.ones((3, 1, 2))))
    assert np.allclose(x[:, 0, 0], [0.5, 0.25, -0.125])
    # block diagonal of a 2x2 and a 3x3 block (natural order inside blocks of size <= 3)
    A = np.zeros((5, 5))
    A[:2, :2] = [[4, 1], [2, 5]]
    A[2:, 2:] = [[6, 1, 0], [1, 7, 2], [3, 0, 8]]
    Ai, Aj = np.nonzero(A)
    b = np.arange(10.0).reshape(5, 2)
    assert relerr(np_(K.solve(Ai, Aj, A[Ai, Aj], b)), np.linalg.solve(A, b)) < 1e-13


def test_many_right_hand_sides_in_the_shared_memory_regime(K, oracle):
    """n_rhs larger than the chunk the value array can hold: the kernel re-runs the
    solve-only program per chunk while the factors stay in shared memory."""
    Ai, Aj, Ax, b = synth.ref_style_case(20, 60, 3, 17, np.complex128)
    so = oracle.analyze(Ai, Aj, 20)
    with K.analyze(Ai, Aj, 20) as sym:
        assert relerr(np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym)), oracle.solve_with_symbol(Ai, Aj, Ax, b, so)) < X_TOL
        with K.factor(Ai, Aj, Ax, sym) as num:
            assert relerr(np_(K.tsolve_with_numeric(num, b, sym)),
                          oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)) < X_TOL
    oracle.free_symbolic(so)


def test_float64_sax_pattern(K, oracle):
    """The shared-memory regime in real arithmetic on the cfg2 pattern."""
    pat = synth.SaxPattern(64, seed=1)
    L = 40
    Ax = np.ascontiguousarray(pat.values(0, L).real) * 2.0
    Ax[:, pat.Ai == pat.Aj] = 1.0
    b = np.random.default_rng(2).standard_normal((L, 64, 2))
    so = oracle.analyze(pat.Ai, pat.Aj, 64)
    with K.analyze(pat.Ai, pat.Aj, 64) as sym:
        x = np_(K.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym))
    assert relerr(x, oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)) < X_TOL
    assert residual(pat.Ai, pat.Aj, Ax[3], x[3], b[3], 64) < RES_TOL
    oracle.free_symbolic(so)


def test_duplicate_entries_are_rejected(K):
    Ai, Aj = np.array([0, 0, 1], np.int32), np.array([0, 0, 1], np.int32)
    with K.analyze(Ai, Aj, 2) as sym:
        with pytest.raises(K.KlujaxCudaError) as ei:
            K.solve_with_symbol(Ai, Aj, np.ones((1, 3)), np.ones((1, 2, 1)), sym)
        assert ei.value.code == -3


def test_free_semantics(K):
    Ai, Aj, Ax, b = synth.ref_style_case(5, 15, 3, 1, np.float64)
    sym = K.analyze(Ai, Aj, 5)
    num = K.factor(Ai, Aj, Ax, sym)
    h = num.handle.copy()
    assert K.free_numeric(h[:1]) == 1           # element-wise free (klujax.cpp:1282-1288)
    with pytest.raises(K.KlujaxCudaError):      # a freed element is stale
        K.solve_with_numeric(h, b, sym)
    two = K.KLUHandleManager(h[1:], None, owner=False)
    assert np_(K.solve_with_numeric(two, b[1:], sym)).shape == (2, 5, 1)
    num.close()                                  # frees the rest; second close is a no-op
    num.close()
    assert K.free_symbolic(sym) == 0             # manager path returns 0 (klujax.py:251-253)
    assert K.free_symbolic(np.uint64(0)) == 0    # null handle: nothing to free


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_both_shared_memory_kernels(K, oracle, wide, dtype, monkeypatch):
    """One warp per system (kernels_small.cuh) and four warps per system (kernels_wide.cuh)
    give the oracle's answers on the same inputs, for every entry point."""
    monkeypatch.setenv("KLUJAX_CUDA_WIDE", wide)
    Ai, Aj, Ax, b = synth.ref_style_case(40, 160, 7, 5, dtype)
    so = oracle.analyze(Ai, Aj, 40)
    xo = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
    xto = oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)
    with K.analyze(Ai, Aj, 40) as sym:
        assert relerr(np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym)), xo) < X_TOL
        assert relerr(np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym)), xto) < X_TOL
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0 and 0.0 < float(gr.max()) <= 1000.0
        with K.factor(Ai, Aj, Ax, sym) as num:
            assert relerr(np_(K.solve_with_numeric(num, b, sym)), xo) < X_TOL
            assert relerr(np_(K.tsolve_with_numeric(num, b, sym)), xto) < X_TOL
            x3, _ = K.refactor_and_solve(Ai, Aj, Ax * 2.0, b, num, sym)
            assert relerr(np_(x3) * 2.0, xo) < X_TOL
    oracle.free_symbolic(so)
    if dtype is np.complex128:  # and on the cfg5 pattern
        pat = synth.SaxPattern(128, seed=2)
        Axs = pat.values(0, 12)
        bs = np.random.default_rng(1).standard_normal((12, 128, 2)) + 0j
        so = oracle.analyze(pat.Ai, pat.Aj, 128)
        with K.analyze(pat.Ai, pat.Aj, 128) as sym:
            assert relerr(np_(K.solve_with_symbol(pat.Ai, pat.Aj, Axs, bs, sym)),
                          oracle.solve_with_symbol(pat.Ai, pat.Aj, Axs, bs, so)) < X_TOL
        oracle.free_symbolic(so)


def test_host_batches_are_streamed_in_chunks(K):
    """Host-resident batches of >= 8192 systems go through the chunked two-stream path
    (copy of chunk k+1 overlaps the kernels of chunk k): same answers, same status."""
    import torch
    L = 8192 + 37  # ragged chunk sizes
    Ai, Aj, Ax1, b1 = synth.ref_style_case(12, 40, 1, 2, np.complex128)
    rng = np.random.default_rng(5)
    Ax = Ax1 * (1.0 + 0.1 * rng.standard_normal((L, 1)))
    b = b1 + rng.standard_normal((L, 12, 2))
    with K.analyze(Ai, Aj, 12) as sym:
        x_host = K.solve_with_symbol(Ai, Aj, Ax, b, sym)                      # numpy in: streamed
        st_host = K.last_status()[0].clone()
        x_dev = K.solve_with_symbol(Ai, Aj, torch.from_numpy(Ax).cuda(), torch.from_numpy(b).cuda(), sym)
        xt_host = K.tsolve_with_symbol(Ai, Aj, torch.from_numpy(Ax).pin_memory(), torch.from_numpy(b).pin_memory(), sym)
        xt_dev = K.tsolve_with_symbol(Ai, Aj, torch.from_numpy(Ax).cuda(), torch.from_numpy(b).cuda(), sym)
    assert torch.equal(x_host, x_dev) and torch.equal(xt_host, xt_dev)
    assert int(st_host.abs().max()) == 0
    m = 4321
    assert relerr(np_(x_host)[m], np.linalg.solve(dense(Ai, Aj, Ax[m], 12), b[m])) < 1e-12


@pytest.mark.gpu
def test_host_result_buffer_out(K):
    """`out=` (extension): a CPU result buffer filled by per-chunk device->host copies gives the
    same bytes as copying the device result, for the streamed and the unstreamed path."""
    import torch
    Ai, Aj, Ax1, b1 = synth.ref_style_case(12, 40, 1, 2, np.complex128)
    rng = np.random.default_rng(6)
    for L in (33, 8192 + 5):
        Ax = torch.from_numpy(Ax1 * (1.0 + 0.1 * rng.standard_normal((L, 1)))).pin_memory()
        b = torch.from_numpy(b1 + rng.standard_normal((L, 12, 2))).pin_memory()
        out = torch.empty((L, 12, 2), dtype=torch.complex128).pin_memory()
        with K.analyze(Ai, Aj, 12) as sym:
            ret = K.solve_with_symbol(Ai, Aj, Ax, b, sym, out=out)
            ref = K.solve_with_symbol(Ai, Aj, Ax.cuda(), b.cuda(), sym).cpu()
            assert ret.data_ptr() == out.data_ptr() and torch.equal(out, ref)
            K.tsolve_with_symbol(Ai, Aj, Ax, b, sym, out=out)
            assert torch.equal(out, K.tsolve_with_symbol(Ai, Aj, Ax.cuda(), b.cuda(), sym).cpu())
            with pytest.raises(ValueError):
                K.solve_with_symbol(Ai, Aj, Ax, b, sym, out=torch.empty(3, dtype=torch.complex128))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_fused_value_and_jvp_vjp(K, dtype):
    """One factorization serves the primal, the JVP and the transpose solve ("next" row f2):
    same numbers as the reference's composition of rules (klujax.py:1856-1897, 1948-1957,
    1712-1714), checked against dense linear algebra."""
    import torch
    L, n, r = 3, 12, 2
    Ai, Aj, Ax, b = synth.ref_style_case(n, 40, L, r, dtype)
    rng = np.random.default_rng(11)
    cplx = np.iscomplexobj(Ax)
    def rnd(shape):
        v = rng.standard_normal(shape)
        return v + 1j * rng.standard_normal(shape) if cplx else v
    t_Ax, t_b, ct = rnd(Ax.shape), rnd(b.shape), rnd(b.shape)
    with K.analyze(Ai, Aj, n) as sym:
        x, dx = K.solve_with_symbol_value_and_jvp(Ai, Aj, Ax, b, sym, t_Ax=t_Ax, t_b=t_b)
        x2, b_bar, Ax_bar = K.solve_with_symbol_value_and_vjp(Ai, Aj, Ax, b, sym, ct)
        # the reference's own composition: two more factorizations
        dx_ref = K.solve_with_symbol(Ai, Aj, Ax, torch.from_numpy(t_b).cuda() - K.dot(Ai, Aj, t_Ax, x), sym)
        b_bar_ref = K.tsolve_with_symbol(Ai, Aj, Ax, ct, sym)
    assert relerr(np_(dx), np_(dx_ref)) < 1e-12 and relerr(np_(b_bar), np_(b_bar_ref)) < 1e-12
    assert torch.equal(x, x2)
    for m in range(L):
        A = dense(Ai, Aj, Ax[m], n)
        tA = dense(Ai, Aj, t_Ax[m], n)
        xm = np.linalg.solve(A, b[m])
        assert relerr(np_(x)[m], xm) < 1e-11
        assert relerr(np_(dx)[m], np.linalg.solve(A, t_b[m] - tA @ xm)) < 1e-10
        bb = np.linalg.solve(A.T, ct[m])
        assert relerr(np_(b_bar)[m], bb) < 1e-10
        assert relerr(np_(Ax_bar)[m], -(bb[Ai] * xm[Aj]).sum(-1)) < 1e-10
    # the gather kernel alone, against numpy
    g = np_(K.dot_transpose_ax(Ai, Aj, ct, b))
    assert relerr(g, (ct[:, Ai] * b[:, Aj]).sum(-1)) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("regime", ["0", "1"])
def test_systems_that_need_other_pivots_are_reseeded(K, oracle, regime, monkeypatch):
    """The batch runs with the pivot sequence of its first system.  A system whose growth
    certificate fails (KLU's own rule would have exchanged rows: tiny or zero diagonal) is solved
    again with a pivot sequence of its own, so x matches the reference's per-system klu_factor."""
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", regime)
    n, L = 8, 6
    rng = np.random.default_rng(21)
    A = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.5)
    A += np.diag(4.0 + np.abs(A).sum(1))
    Ai, Aj = np.nonzero(A)
    Ai, Aj = Ai.astype(np.int32), Aj.astype(np.int32)
    Ax = np.tile(A[Ai, Aj], (L, 1)) * (1.0 + 0.05 * rng.standard_normal((L, Ai.size)))
    dpos = {int(i): k for k, (i, j) in enumerate(zip(Ai, Aj)) if i == j}
    Ax[2, dpos[3]] = 1e-9     # tiny diagonal: the static pivot would give |L| ~ 1e9
    Ax[4, dpos[1]] = 0.0      # zero diagonal: the static pivot is exactly singular
    Ax[4, dpos[5]] = 1e-12
    b = rng.standard_normal((L, n, 2))
    so = oracle.analyze(Ai, Aj, n)
    ref = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
    oracle.free_symbolic(so)
    with K.analyze(Ai, Aj, n) as sym:
        raw = np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym, check=False))
        st, gr = (np_(t).copy() for t in K.last_status())
        assert st.tolist() == [0, 0, 2, 0, 1, 0] and gr[2] > 1e6   # 2: pivot-parity certificate failed
        x = np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym))          # check=True: re-seeds 2 and 4
        st2, gr2 = (np_(t) for t in K.last_status())
        assert st2.tolist() == [0] * L and gr2.max() <= 1000.0 * (1 + 1e-9)
        assert relerr(x, ref) < X_TOL
        for m in (0, 1, 3, 5):
            assert np.array_equal(x[m], raw[m])                   # certified systems are untouched
        xt = np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym))
        for m in range(L):
            Am = dense(Ai, Aj, Ax[m], n)
            assert relerr(xt[m], np.linalg.solve(Am.T, b[m])) < 1e-9


@pytest.mark.gpu
def test_cfg4_full_size_properties(K, oracle):
    """BASELINE.json config 4 at full size (n = 20 000, 128 systems, planted BTF): refactor +
    solve + transpose solve, checked through size-independent properties (residuals of every
    system through the COO product, status, growth certificate, solve/tsolve adjointness)."""
    import torch
    pat = synth.MnaPattern(n=20000)
    n, L = pat.n, 128
    Ai, Aj = torch.from_numpy(pat.Ai).cuda(), torch.from_numpy(pat.Aj).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    b = torch.randn(L, n, 1, dtype=torch.float64, device="cuda", generator=g)
    c = torch.randn(L, n, 1, dtype=torch.float64, device="cuda", generator=g)
    with K.analyze(pat.Ai, pat.Aj, n) as sym, K.factor(Ai, Aj, torch.from_numpy(pat.values(0, L)).cuda(), sym) as num:
        Ax = torch.from_numpy(pat.values(3, L)).cuda()
        num2 = K.refactor(Ai, Aj, Ax, num, sym)
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
        x = K.solve_with_numeric(num2, b, sym)
        y = K.tsolve_with_numeric(num2, c, sym)
    bn = torch.linalg.vector_norm(b, dim=1)
    assert float((torch.linalg.vector_norm(K.dot(Ai, Aj, Ax, x) - b, dim=1) / bn).max()) < RES_TOL
    At_y = K.dot(Aj, Ai, Ax, y)  # transposed triplets
    assert float((torch.linalg.vector_norm(At_y - c, dim=1) / torch.linalg.vector_norm(c, dim=1)).max()) < RES_TOL
    # <A^-T c, b> == <c, A^-1 b>
    lhs, rhs = (y * b).sum((1, 2)), (c * x).sum((1, 2))
    assert float(((lhs - rhs).abs() / (lhs.abs() + rhs.abs() + 1e-300)).max()) < 1e-9
    # a sample of the batch against the oracle: klu_factor at t = 0, klu_refactor at t = 3, solves
    idx = [0, 41, 86, 127]
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    no = oracle.factor(pat.Ai, pat.Aj, pat.values(0, L)[idx], so)
    no = oracle.refactor(pat.Ai, pat.Aj, np_(Ax)[idx], so, no)
    assert relerr(np_(x)[idx], oracle.solve_with_numeric(so, no, np_(b)[idx])) < X_TOL
    assert relerr(np_(y)[idx], oracle.solve_with_numeric(so, no, np_(c)[idx], transpose=True)) < X_TOL
    oracle.free_numeric(no)
    oracle.free_symbolic(so)


@pytest.mark.gpu
def test_cfg3_full_grid_properties(K):
    """BASELINE.json config 3's matrix at full size (256 x 256 Poisson, n = 65 536), with a
    reduced batch (4 systems x 8 right-hand sides) so that the test stays short: factor once,
    solve_with_numeric, residual of every right-hand side and linearity."""
    import torch
    g, L, r = 256, 4, 8
    Ai_h, Aj_h, Ax_h = synth.poisson2d(g, L)
    n = g * g
    Ai, Aj, Ax = torch.from_numpy(Ai_h).cuda(), torch.from_numpy(Aj_h).cuda(), torch.from_numpy(Ax_h).cuda()
    gen = torch.Generator(device="cuda").manual_seed(2)
    b = torch.randn(L, n, r, dtype=torch.float64, device="cuda", generator=gen)
    with K.analyze(Ai_h, Aj_h, n) as sym, K.factor(Ai, Aj, Ax, sym) as num:
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
        x = K.solve_with_numeric(num, b, sym)
        x2 = K.solve_with_numeric(num, b[:, :, :1] - 2.0 * b[:, :, 1:2], sym)
    res = torch.linalg.vector_norm(K.dot(Ai, Aj, Ax, x) - b, dim=1) / torch.linalg.vector_norm(b, dim=1)
    assert float(res.max()) < RES_TOL
    assert float((x2 - (x[:, :, :1] - 2.0 * x[:, :, 1:2])).abs().max() / x.abs().max()) < 1e-12


@pytest.mark.gpu
def test_cfg3_full_rhs_width_vs_oracle(K, oracle):
    """Config 3 at its full matrix size and right-hand-side width (256 x 256 grid, n_rhs = 256),
    two systems: factor once, solve_with_numeric, x against the oracle at 1e-10."""
    import torch
    g, L, r = 256, 2, 256
    Ai_h, Aj_h, Ax_h = synth.poisson2d(g, L)
    n = g * g
    b_h = np.random.default_rng(3).standard_normal((L, n, r))
    so = oracle.analyze(Ai_h, Aj_h, n)
    no = oracle.factor(Ai_h, Aj_h, Ax_h, so)
    ref = oracle.solve_with_numeric(so, no, b_h)
    oracle.free_numeric(no)
    oracle.free_symbolic(so)
    with K.analyze(Ai_h, Aj_h, n) as sym, K.factor(Ai_h, Aj_h, Ax_h, sym) as num:
        x = np_(K.solve_with_numeric(num, torch.from_numpy(b_h).cuda(), sym))
    assert relerr(x, ref) < X_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("r", [1, 4])
def test_cfg2_full_batch_vs_oracle(K, oracle, r):
    """Config 2 at its full batch (n_col = 64, n_lhs = 4096), one and four right-hand sides:
    every system against the oracle."""
    n, L = 64, 4096
    pat = synth.SaxPattern(n, seed=1)
    Ax = pat.values_fast(L, seed=4)
    rng = np.random.default_rng(6)
    b = rng.standard_normal((L, n, r)) + 1j * rng.standard_normal((L, n, r))
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)
    reft = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so, transpose=True)
    oracle.free_symbolic(so)
    with K.analyze(pat.Ai, pat.Aj, n) as sym:
        assert relerr(np_(K.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym)), ref) < X_TOL
        assert relerr(np_(K.tsolve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym)), reft) < X_TOL


@pytest.mark.gpu
def test_rowreg_kernel_matches_the_shared_memory_kernels(K, oracle, monkeypatch):
    """The register-resident kernel (opt-in, generated PTX + driver JIT) on the cfg 2 shape."""
    n, L = 64, 512
    pat = synth.SaxPattern(n, seed=1)
    Ax = pat.values_fast(L, seed=8)
    rng = np.random.default_rng(2)
    b = rng.standard_normal((L, n, 2)) + 1j * rng.standard_normal((L, n, 2))
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)
    oracle.free_symbolic(so)
    monkeypatch.setenv("KLUJAX_CUDA_ROWREG", "1")
    monkeypatch.setenv("KLUJAX_CUDA_ROWREG_REQUIRE", "1")
    with K.analyze(pat.Ai, pat.Aj, n) as sym:
        x = np_(K.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym))
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
    assert relerr(x, ref) < X_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_register_tail_stage_matches_the_interpreter(K, oracle, dtype, monkeypatch):
    """The dense-tail register stage of the multi-warp kernel (opt-in, profiles/tail_stage_r02.md)
    on the cfg 5 shape: same solution as the oracle and as the default interpreted schedule."""
    n, L = 128, 300
    pat = synth.SaxPattern(n, seed=2)
    Ax = pat.values_fast(L, seed=7)
    rng = np.random.default_rng(3)
    b = rng.standard_normal((L, n, 1)) + 1j * rng.standard_normal((L, n, 1))
    if dtype == np.float64:
        Ax, b = np.ascontiguousarray(Ax.real + Ax.imag), np.ascontiguousarray(b.real)
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)
    oracle.free_symbolic(so)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KLUJAX_CUDA_TAIL", mode)
        with K.analyze(pat.Ai, pat.Aj, n) as sym:
            out[mode] = np_(K.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym))
            st, gr = K.last_status()
            assert int(st.abs().max()) == 0 and float(gr.max()) <= 1000.0
        assert relerr(out[mode], ref) < X_TOL
    assert relerr(out["1"], out["0"]) < X_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("n", [256, 320])
def test_large_shared_memory_systems(K, oracle, n):
    """Systems that fill most of an SM's shared memory (one or two resident systems): n = 256
    still packs for the multi-warp kernel, n = 320 exceeds its control-word table and falls back
    to the one-warp kernel inside the same regime."""
    pat = synth.SaxPattern(n, seed=4)
    L = 5
    Ax = pat.values(0, L)
    b = np.random.default_rng(8).standard_normal((L, n, 1)) + 0j
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)
    oracle.free_symbolic(so)
    with K.analyze(pat.Ai, pat.Aj, n) as sym:
        x = np_(K.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, sym))
        st, gr = K.last_status()
        assert int(st.abs().max()) == 0
    assert relerr(x, ref) < X_TOL


@pytest.mark.gpu
def test_fuzz_structures_all_kernels(K, oracle, monkeypatch):
    """Seeded fuzz over structures the schedule passes (step balancing, bank-aware lane order,
    slot compaction) have to survive: random sparse patterns with and without a dominant
    diagonal, arrow-heads, block-triangular chains, dense small matrices; both dtypes, one and
    several right-hand sides, both shared-memory kernels and the level-scheduled regime; x against
    the oracle and dense LAPACK."""
    rng = np.random.default_rng(1234)

    def pattern(kind, n):
        A = np.zeros((n, n))
        if kind == "random":
            A = rng.standard_normal((n, n)) * (rng.random((n, n)) < min(0.5, 4.0 / n))
        elif kind == "arrow":
            A[0, :] = rng.standard_normal(n); A[:, 0] = rng.standard_normal(n)
            A[-1, :] = rng.standard_normal(n)
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
        p = rng.permutation(n)
        return A[p]  # rows permuted: the diagonal is hidden, the transversal has to find it

    cases = [("random", 37), ("random", 90), ("arrow", 48), ("btf", 64), ("dense", 24), ("band", 120),
             ("random", 150), ("btf", 130)]
    for ci, (kind, n) in enumerate(cases):
        A = pattern(kind, n)
        Ai, Aj = (v.astype(np.int32) for v in np.nonzero(A))
        for dtype in (np.float64, np.complex128):
            L, r = (6, 1) if ci % 2 == 0 else (3, 5)
            Ax = np.tile(A[Ai, Aj], (L, 1)) * (1.0 + 0.05 * rng.standard_normal((L, Ai.size)))
            b = rng.standard_normal((L, n, r))
            if dtype is np.complex128:
                Ax = Ax * (1.0 + 0.1j)
                b = b + 1j * rng.standard_normal((L, n, r))
            so = oracle.analyze(Ai, Aj, n)
            ref = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
            reft = oracle.solve_with_symbol(Ai, Aj, Ax, b, so, transpose=True)
            oracle.free_symbolic(so)
            # The transversal need not recover the dominant diagonal of a row-permuted matrix,
            # so KLU's tol = 1e-3 pivoting may run on small pivots: the bar is then the oracle's
            # own distance from the dense LAPACK solution, not 1e-10.
            true = np.stack([np.linalg.solve(dense(Ai, Aj, Ax[m], n), b[m]) for m in range(L)])
            truet = np.stack([np.linalg.solve(dense(Ai, Aj, Ax[m], n).T, b[m]) for m in range(L)])
            tol = max(X_TOL, 10.0 * relerr(ref, true))
            tolt = max(X_TOL, 10.0 * relerr(reft, truet))
            for wide in ("0", "1", "level"):
                if wide == "level":  # the level-scheduled regime on the same systems
                    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", "1")
                else:
                    monkeypatch.delenv("KLUJAX_CUDA_FORCE_REGIME", raising=False)
                    monkeypatch.setenv("KLUJAX_CUDA_WIDE", wide)
                with K.analyze(Ai, Aj, n) as sym:
                    x = np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym))
                    xt = np_(K.tsolve_with_symbol(Ai, Aj, Ax, b, sym))
                    with K.factor(Ai, Aj, Ax, sym) as num:
                        xn = np_(K.solve_with_numeric(num, b, sym))
                assert relerr(x, ref) < tol and relerr(x, true) < tol, (kind, n, dtype, wide)
                assert relerr(xt, reft) < tolt and relerr(xt, truet) < tolt, (kind, n, dtype, wide)
                assert relerr(xn, ref) < tol, (kind, n, dtype, wide)


@pytest.mark.gpu
def test_one_process_sharded_over_the_visible_gpus(K, oracle):
    """solve_with_symbol_sharded: the n_lhs batch cut into contiguous blocks over every visible
    GPU, one process, replicated symbolic handles, no data-path collective (the reference's only
    multi-device test is /root/reference/tests.py:472-492).  With one GPU it degenerates to one
    shard; with several the blocks run concurrently."""
    import torch
    n, L = 64, 1031  # ragged over any device count
    pat = synth.SaxPattern(n, seed=1)
    Ax = pat.values_fast(L, seed=3)
    rng = np.random.default_rng(4)
    b = rng.standard_normal((L, n, 2)) + 1j * rng.standard_normal((L, n, 2))
    so = oracle.analyze(pat.Ai, pat.Aj, n)
    ref = oracle.solve_with_symbol(pat.Ai, pat.Aj, Ax, b, so)
    oracle.free_symbolic(so)
    devs = list(range(torch.cuda.device_count()))
    with K.analyze_sharded(pat.Ai, pat.Aj, n, devs) as ssym:
        x = K.solve_with_symbol_sharded(pat.Ai, pat.Aj, torch.from_numpy(Ax).pin_memory(),
                                        torch.from_numpy(b).pin_memory(), ssym)
        assert relerr(x.numpy(), ref) < X_TOL
        parts = K.solve_with_symbol_sharded(pat.Ai, pat.Aj, Ax, b, ssym, gather=False)
        assert len(parts) == len(devs) and sum(p.shape[0] for p in parts if p is not None) == L
        if len(devs) > 1:
            assert {p.device.index for p in parts if p is not None} == set(devs)
            # a handle is bound to its device
            with torch.cuda.device(devs[1]):
                with pytest.raises(K.KlujaxCudaError):
                    K.solve_with_symbol(pat.Ai, pat.Aj, Ax[:4], b[:4], ssym.handles[0])
