"""Pivot-parity certificate and re-seeding (ADVICE
high): the static pivot sequence is certified per system, a failed certificate is a status code
(2 = KLUJAX_CUDA_PIVOT_DRIFT), the fused solves re-seed below the C-ABI, factor / refactor fail
loudly.  The reference pivots every system afresh (klu_factor at klujax.cpp:310/441/679): the
oracle does the same, so x must match it for every system."""
import numpy as np
import pytest

from util import dense, relerr

pytestmark = pytest.mark.gpu
X_TOL = 1e-10


@pytest.fixture(scope="module")
def K():
    import torch
    assert torch.cuda.is_available()
    import klujax_b200
    from klujax_b200 import _cabi
    _cabi.lib()
    return klujax_b200


def np_(t):
    return t.detach().cpu().numpy()


def _offdiag_case():
    """5x5 with a dense 3x3 diagonal block {2,3,4} (natural order inside a block of size <= 3,
    klu_analyze) whose first column has a tiny diagonal: KLU's threshold test fails for it and
    the column maximum, row 4, becomes the pivot (an off-diagonal static pivot)."""
    n = 5
    A = np.zeros((n, n))
    A[0, 0] = A[1, 1] = 3.0
    A[0, 1], A[0, 3], A[1, 4] = 0.4, 0.7, -0.6          # coupling to later blocks only
    A[2:, 2:] = [[1e-7, 0.3, 0.2],
                 [0.5, 3.0, 0.4],
                 [2.0, 0.3, 3.0]]
    Ai, Aj = np.nonzero(A)
    return n, Ai.astype(np.int32), Aj.astype(np.int32), A


@pytest.mark.parametrize("regime", ["0", "1"])
def test_offdiagonal_static_pivot_certificate(K, oracle, regime, monkeypatch):
    from klujax_b200 import _cabi
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", regime)
    n, Ai, Aj, A = _offdiag_case()
    pos = {(int(i), int(j)): k for k, (i, j) in enumerate(zip(Ai, Aj))}
    base = A[Ai, Aj]
    rng = np.random.default_rng(8)
    L = 5
    Ax = np.tile(base, (L, 1))
    Ax[1] *= 1.0 + 1e-3 * rng.standard_normal(Ax.shape[1])   # same pivots: certified
    Ax[2, pos[(3, 2)]] = 5.0     # a larger candidate than the static pivot: KLU would take row 3
    Ax[3, pos[(2, 2)]] = 1.0     # the diagonal candidate passes the threshold test again
    Ax[4] *= 1.0 + 1e-3 * rng.standard_normal(Ax.shape[1])
    b = rng.standard_normal((L, n, 2))
    so = oracle.analyze(Ai, Aj, n)
    ref = oracle.solve_with_symbol(Ai, Aj, Ax, b, so)
    oracle.free_symbolic(so)
    with K.analyze(Ai, Aj, n) as sym:
        raw = np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym, check=False))
        st, gr = (np_(t).copy() for t in K.last_status())
        info = _cabi.symbolic_info(int(np.asarray(sym.handle).reshape(-1)[0]), False)
        assert info["noffdiag"] >= 1, "the seed system must need an off-diagonal pivot for this test"
        assert st.tolist() == [0, 0, 2, 2, 0], (st, gr)
        assert gr[0] <= 1000.0 and gr[1] <= 1000.0 and gr[2] > 1000.0 and gr[3] > 1000.0
        x = np_(K.solve_with_symbol(Ai, Aj, Ax, b, sym))       # check=True: re-seeds 2 and 3
        st2, _ = (np_(t) for t in K.last_status())
        assert st2.tolist() == [0] * L
        assert relerr(x, ref) < X_TOL
        for m in (0, 1, 4):
            assert np.array_equal(x[m], raw[m])
        # factor(check=True): klu_factor pivots every system (klujax.cpp:679) - systems 2 and 3 get
        # pivot sequences of their own in child batches; every entry point takes the mixed handles
        so = oracle.analyze(Ai, Aj, n)
        no = oracle.factor(Ai, Aj, Ax, so)
        with K.factor(Ai, Aj, Ax, sym) as num:
            hs = np.asarray(num.handle).reshape(-1)
            assert len({int(h) >> 32 for h in hs}) >= 2
            assert relerr(np_(K.solve_with_numeric(num, b, sym)), ref) < X_TOL
            reft = oracle.solve_with_numeric(so, no, b, transpose=True)
            assert relerr(np_(K.tsolve_with_numeric(num, b, sym)), reft) < X_TOL
            # klu_refactor keeps the pivot order of each handle: same in the oracle
            Ax2 = Ax * (1.0 + 0.01 * rng.standard_normal(Ax.shape))
            no = oracle.refactor(Ai, Aj, Ax2, so, no)
            ref2 = oracle.solve_with_numeric(so, no, b)
            num2 = K.refactor(Ai, Aj, Ax2, num, sym)
            assert relerr(np_(K.solve_with_numeric(num2, b, sym)), ref2) < X_TOL
            x3, _ = K.refactor_and_solve(Ai, Aj, Ax2, b, num, sym)
            assert relerr(np_(x3), ref2) < X_TOL
            # a permuted subset of the handles, one right-hand side broadcast
            sel = [3, 0, 2]
            xs = np_(K.solve_with_numeric(hs[sel], b[:1], sym))
            assert relerr(xs, oracle.solve_with_numeric(so, no, np.repeat(b[:1], 5, 0))[sel]) < X_TOL
        oracle.free_numeric(no)
        oracle.free_symbolic(so)
        # check=False keeps ONE pivot sequence and only reports the certificate
        num = K.factor(Ai, Aj, Ax, sym, check=False)
        st3, _ = (np_(t).copy() for t in K.last_status())
        assert st3.tolist() == [0, 0, 2, 2, 0]
        K.free_numeric(num)


def test_device_resident_handle_arrays(K, oracle):
    """The _dev entry points: numeric handles as u64 device buffers (what XLA delivers on CUDA)."""
    import ctypes as C
    import torch
    from klujax_b200 import _cabi, synth
    lib = _cabi.lib()
    n, L, r = 20, 9, 2
    Ai, Aj, Ax, b = synth.ref_style_case(n, 60, L, r, np.float64)
    so = oracle.analyze(Ai, Aj, n)
    no = oracle.factor(Ai, Aj, Ax, so)
    ref = oracle.solve_with_numeric(so, no, b)
    dev = torch.device("cuda")
    Ai_d, Aj_d = torch.tensor(Ai, device=dev), torch.tensor(Aj, device=dev)
    Ax_d, b_d = torch.tensor(Ax, device=dev), torch.tensor(b, device=dev)
    x_d = torch.empty_like(b_d)
    num_d = torch.zeros(L, dtype=torch.int64, device=dev)      # u64 handles
    st = torch.empty(L, dtype=torch.int32, device=dev)
    gr = torch.empty(L, dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = _cabi.analyze(Ai, Aj, n)
    fb = C.c_int32(-1)
    _cabi.check(lib.klujax_cuda_factor_dev_f64(C.c_uint64(h), L, Ai.size, p(Ai_d), p(Aj_d), p(Ax_d), p(num_d), p(st),
                                               p(gr), 1, C.byref(fb), stream))
    _cabi.check(lib.klujax_cuda_solve_with_numeric_dev_f64(C.c_uint64(h), L, p(num_d), L, n, r, p(b_d), p(x_d), stream))
    torch.cuda.synchronize()
    assert (num_d != 0).all() and int(st.abs().max()) == 0
    assert relerr(x_d.cpu().numpy(), ref) < X_TOL
    Ax2 = Ax * 1.01
    no = oracle.refactor(Ai, Aj, Ax2, so, no)
    ref2 = oracle.solve_with_numeric(so, no, b)
    out_d = torch.zeros_like(num_d)
    _cabi.check(lib.klujax_cuda_refactor_and_solve_dev_f64(C.c_uint64(h), L, n, r, Ai.size, p(Ai_d), p(Aj_d),
                                                           p(torch.tensor(Ax2, device=dev)), p(b_d), p(num_d), p(x_d),
                                                           p(out_d), p(st), p(gr), stream))
    torch.cuda.synchronize()
    assert torch.equal(out_d, num_d)
    assert relerr(x_d.cpu().numpy(), ref2) < X_TOL
    freed = C.c_int32(0)
    _cabi.check(lib.klujax_cuda_free_numeric_dev(L, p(num_d), C.byref(freed), stream))
    assert freed.value == 1
    rc = lib.klujax_cuda_solve_with_numeric_dev_f64(C.c_uint64(h), L, p(num_d), L, n, r, p(b_d), p(x_d), stream)
    assert rc == -3            # stale handles
    _cabi.free_symbolic(h)
    oracle.free_numeric(no)
    oracle.free_symbolic(so)


@pytest.mark.parametrize("regime", ["0", "1"])
def test_single_system_calls_with_changing_pivots(K, oracle, regime, monkeypatch):
    """A Newton / transient loop: analyze once, then one system per call.  The pivot sequence
    was fixed by the first call; a later system that needs a row exchange must still get the
    reference's answer."""
    monkeypatch.setenv("KLUJAX_CUDA_FORCE_REGIME", regime)
    n = 8
    rng = np.random.default_rng(21)
    A = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.5)
    A += np.diag(4.0 + np.abs(A).sum(1))
    Ai, Aj = np.nonzero(A)
    Ai, Aj = Ai.astype(np.int32), Aj.astype(np.int32)
    dpos = {int(i): k for k, (i, j) in enumerate(zip(Ai, Aj)) if i == j}
    A0 = A[Ai, Aj]
    A1 = A0.copy(); A1[dpos[3]] = 1e-9         # tiny diagonal
    A2 = A0.copy(); A2[dpos[1]] = 0.0          # zero diagonal: the static pivot is singular
    b = rng.standard_normal((n,))
    with K.analyze(Ai, Aj, n) as sym:
        for Axm in (A0, A1, A2, A0):
            x = np_(K.solve_with_symbol(Ai, Aj, Axm, b, sym))
            assert relerr(x, np.linalg.solve(dense(Ai, Aj, Axm, n), b)) < 1e-9
            xt = np_(K.tsolve_with_symbol(Ai, Aj, Axm, b, sym))
            assert relerr(xt, np.linalg.solve(dense(Ai, Aj, Axm, n).T, b)) < 1e-9
    for Axm in (A0, A1, A2):                     # the one-shot solve keeps a pattern cache
        x = np_(K.solve(Ai, Aj, Axm, b))
        assert relerr(x, np.linalg.solve(dense(Ai, Aj, Axm, n), b)) < 1e-9


def test_broadcast_matrix_is_not_materialised_and_reseeds(K):
    """One matrix, many right-hand sides (klujax.py:1844-1848): stride 0 on the device; a failed
    certificate of the one matrix redoes the whole call."""
    n = 8
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.6)
    A += np.diag(4.0 + np.abs(A).sum(1))
    Ai, Aj = np.nonzero(A)
    Ai, Aj = Ai.astype(np.int32), Aj.astype(np.int32)
    dpos = {int(i): k for k, (i, j) in enumerate(zip(Ai, Aj)) if i == j}
    A0 = A[Ai, Aj]
    A1 = A0.copy(); A1[dpos[2]] = 1e-10
    b = rng.standard_normal((7, n, 3))
    with K.analyze(Ai, Aj, n) as sym:
        for Axm in (A0, A1):
            x = np_(K.solve_with_symbol(Ai, Aj, Axm, b, sym))
            assert x.shape == (7, n, 3)
            D = dense(Ai, Aj, Axm, n)
            for m in range(7):
                assert relerr(x[m], np.linalg.solve(D, b[m])) < 1e-9
        # many matrices, one right-hand side
        Ax = np.stack([A0 * (1 + 0.01 * k) for k in range(4)])
        x = np_(K.solve_with_symbol(Ai, Aj, Ax, b[:1], sym))
        for m in range(4):
            assert relerr(x[m], np.linalg.solve(dense(Ai, Aj, Ax[m], n), b[0])) < 1e-9
