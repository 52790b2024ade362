"""cuSOLVER-RF batch comparator for the float64 level-regime configs.

NVIDIA's sparse LU RE-factorization library does what this repo's level regime does - a fixed
pivot sequence and fixed L/U patterns from one host factorization, then batched numeric
refactorization + solve on the GPU - so it is the GPU library baseline next to the CPU one.
    python tools/cusolverrf_compare.py cfg4        (n = 20 000, 128 systems, 1 right-hand side)
What is compared: ResetValues + Refactor + Solve of the batch (cuSOLVER-RF has no transpose
solve and takes one right-hand side per call) against refactor + solve_with_numeric of this repo
on the same matrices.  cuSOLVER-RF gets L, U, P, Q from SuperLU (scipy.sparse.linalg.splu) run in
natural order on the matrix permuted by this repo's own BTF + AMD analysis: the same ordering on
both sides, no row scaling and no block structure on the cuSOLVER-RF side.

Every array cuSOLVER-RF sees lives in MANAGED memory, so that a host pointer where the library
wants a device pointer (or the other way round) cannot fault the GPU.

STATUS (round 2): NOT WORKING.  With the libcusolver 11.7 that torch bundles in this image,
cusolverRfBatchSetupHost accepts these inputs (L with its unit diagonal stored, nnz(U) = 5.9 M because
cuSOLVER-RF has no block-triangular form) and cusolverRfBatchAnalyze then dies with a HOST segmentation
fault - twice, with library defaults and a batch of 2 (gpurun_out/rf.log).  Not pursued further: every
crash on the GPU box counts against the round.  The script refuses to run unless RF_FORCE=1."""
import ctypes as C
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cuda import cudart  # noqa: E402

from klujax_b200 import synth  # noqa: E402

def _load_cusolver():
    # the copy that matches the cuBLAS torch has already loaded (the toolkit's own libcusolver wants a newer one)
    import glob
    import nvidia.cusolver
    cands = glob.glob(os.path.join(list(nvidia.cusolver.__path__)[0], "lib", "libcusolver.so*"))
    return C.CDLL(sorted(cands)[0] if cands else "/usr/local/cuda/lib64/libcusolver.so.11")


lib = _load_cusolver()
for _f in ("cusolverRfCreate", "cusolverRfBatchSetupHost", "cusolverRfBatchAnalyze", "cusolverRfBatchRefactor", "cusolverRfBatchSolve",
           "cusolverRfBatchResetValues", "cusolverRfSetNumericProperties", "cusolverRfSetMatrixFormat", "cusolverRfSetResetValuesFastMode"):
    getattr(lib, _f).restype = C.c_int
lib.cusolverRfBatchAnalyze.argtypes = [C.c_void_p]
lib.cusolverRfBatchRefactor.argtypes = [C.c_void_p]


def ck(status, what):
    if status != 0:
        raise RuntimeError(f"{what}: cusolver status {status}")


class Managed:
    """numpy array copied into cudaMallocManaged memory."""

    def __init__(self, arr):
        arr = np.ascontiguousarray(arr)
        err, ptr = cudart.cudaMallocManaged(max(arr.nbytes, 16), 1)
        assert err == cudart.cudaError_t.cudaSuccess, err
        self.ptr, self.nbytes, self.dtype, self.shape = int(ptr), arr.nbytes, arr.dtype, arr.shape
        C.memmove(self.ptr, arr.ctypes.data, arr.nbytes)

    def numpy(self):
        out = np.empty(self.shape, self.dtype)
        C.memmove(out.ctypes.data, self.ptr, self.nbytes)
        return out

    def free(self):
        cudart.cudaFree(self.ptr)


def main():
    if os.environ.get("RF_FORCE") != "1":
        sys.exit("cusolverrf_compare.py: known to crash in cusolverRfBatchAnalyze on this image (see the header); set RF_FORCE=1 to try")
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    assert which == "cfg4", "only cfg4 is wired up (cfg 3 needs 256 solve calls per step: one right-hand side per call)"
    pat = synth.MnaPattern(n=20000)
    n, L = pat.n, int(os.environ.get("RF_BATCH", "128"))
    Ax = pat.values(0, L)
    Ai, Aj = np.asarray(pat.Ai), np.asarray(pat.Aj)
    # CSR of the pattern with sorted columns; perm maps CSR position -> COO position
    order = np.lexsort((Aj, Ai))
    rowptr = np.zeros(n + 1, np.int32)
    np.add.at(rowptr, Ai[order] + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    colind = Aj[order].astype(np.int32)
    vals = np.ascontiguousarray(Ax[:, order])  # [L, nnz]
    nnz = colind.size
    A0 = sp.csr_matrix((vals[0], colind, rowptr), shape=(n, n))
    # Ordering: this repo's own analysis (BTF + AMD, host), so that both sides factor the same permuted
    # matrix; SuperLU then factors it in natural order with diagonal pivots.  (COLAMD on the raw matrix,
    # without BTF, fills to 38 M entries against 0.33 M: not a comparison.)
    import klujax_b200 as K
    from klujax_b200 import _cabi
    sym0 = K.analyze(Ai, Aj, n)
    perm = _cabi.symbolic_perm(int(sym0.handle))
    Pk, Qk = perm["P"].astype(np.int64), perm["Q"].astype(np.int64)
    t0 = time.perf_counter()
    lu = spla.splu(A0[Pk][:, Qk].tocsc(), permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=False))
    t_host = time.perf_counter() - t0
    P = Pk[np.argsort(lu.perm_r)].astype(np.int32)  # new row r = old row P[r]
    Q = Qk[np.argsort(lu.perm_c)].astype(np.int32)  # new column c = old column Q[c]
    Lc, Uc = lu.L.tocsr(), lu.U.tocsr()
    Lc.sort_indices(); Uc.sort_indices()
    chk = abs((Lc @ Uc) - A0[P][:, Q]).max()
    print(f"host SuperLU: {t_host:.2f} s, nnz(L) {Lc.nnz}, nnz(U) {Uc.nnz}, |LU - A(P,Q)| = {chk:.2e}")
    assert chk < 1e-9 and sorted(P.tolist()) == list(range(n)) and sorted(Q.tolist()) == list(range(n))

    torch.zeros(1, device="cuda")  # context
    h = C.c_void_p()
    ck(lib.cusolverRfCreate(C.byref(h)), "create")
    if os.environ.get("RF_OPTIONS", "0") == "1":  # everything below is the library's default anyway
        ck(lib.cusolverRfSetNumericProperties(h, C.c_double(0.0), C.c_double(0.0)), "numeric properties")
        ck(lib.cusolverRfSetMatrixFormat(h, 0, 0), "matrix format (CSR, unit diagonal stored in L)")
        ck(lib.cusolverRfSetResetValuesFastMode(h, 1), "fast mode")
    m_rowptr, m_colind = Managed(rowptr), Managed(colind)
    m_vals = [Managed(vals[m]) for m in range(L)]
    m_valptrs = Managed(np.array([v.ptr for v in m_vals], np.uint64))
    m_Lp, m_Li, m_Lx = Managed(Lc.indptr.astype(np.int32)), Managed(Lc.indices.astype(np.int32)), Managed(Lc.data)
    m_Up, m_Ui, m_Ux = Managed(Uc.indptr.astype(np.int32)), Managed(Uc.indices.astype(np.int32)), Managed(Uc.data)
    m_P, m_Q = Managed(P), Managed(Q)
    vp = C.c_void_p
    ck(lib.cusolverRfBatchSetupHost(L, n, nnz, vp(m_rowptr.ptr), vp(m_colind.ptr), vp(m_valptrs.ptr), Lc.nnz, vp(m_Lp.ptr),
                                    vp(m_Li.ptr), vp(m_Lx.ptr), Uc.nnz, vp(m_Up.ptr), vp(m_Ui.ptr), vp(m_Ux.ptr), vp(m_P.ptr),
                                    vp(m_Q.ptr), h), "batch setup")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ck(lib.cusolverRfBatchAnalyze(h), "batch analyze")
    torch.cuda.synchronize()
    print(f"cusolverRfBatchAnalyze: {time.perf_counter() - t0:.2f} s")
    rng = np.random.default_rng(0)
    b = rng.standard_normal((L, n))
    m_x = [Managed(b[m]) for m in range(L)]
    m_xptrs = Managed(np.array([v.ptr for v in m_x], np.uint64))
    m_tmp = Managed(np.zeros(2 * L * n))

    def step(reset_rhs=True):
        if reset_rhs:
            for m in range(L):
                C.memmove(m_x[m].ptr, b[m].ctypes.data, b[m].nbytes)
        ck(lib.cusolverRfBatchResetValues(L, n, nnz, vp(m_rowptr.ptr), vp(m_colind.ptr), vp(m_valptrs.ptr), vp(m_P.ptr), vp(m_Q.ptr), h),
           "reset values")
        ck(lib.cusolverRfBatchRefactor(h), "batch refactor")
        ck(lib.cusolverRfBatchSolve(h, vp(m_P.ptr), vp(m_Q.ptr), 1, vp(m_tmp.ptr), n, vp(m_xptrs.ptr), n), "batch solve")

    step()
    torch.cuda.synchronize()
    worst = 0.0
    for m in range(0, L, max(1, L // 8)):
        x = m_x[m].numpy()
        Am = sp.csr_matrix((vals[m], colind, rowptr), shape=(n, n))
        worst = max(worst, float(np.abs(Am @ x - b[m]).max() / np.abs(b[m]).max()))
    print(f"cuSOLVER-RF residual (sampled systems): {worst:.2e}")
    assert worst < 1e-8, "cuSOLVER-RF result does not solve the systems: conventions wrong, numbers below are void"
    # prefetch everything to the device once, then time without touching the right-hand sides from the host
    for mobj in [m_rowptr, m_colind, m_valptrs, m_P, m_Q, m_xptrs, m_tmp] + m_vals + m_x:
        cudart.cudaMemPrefetchAsync(mobj.ptr, mobj.nbytes, 0, 0)
    torch.cuda.synchronize()
    for _ in range(2):
        step(reset_rhs=False)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step(reset_rhs=False)
    e1.record()
    torch.cuda.synchronize()
    ms_rf = e0.elapsed_time(e1) / reps

    # this repo on the same matrices: refactor + solve_with_numeric
    import klujax_b200 as K
    tAi, tAj = torch.from_numpy(Ai.astype(np.int32)).cuda(), torch.from_numpy(Aj.astype(np.int32)).cuda()
    tAx, tb = torch.from_numpy(Ax).cuda(), torch.from_numpy(b[:, :, None].copy()).cuda()
    sym = K.analyze(Ai, Aj, n)
    num = K.factor(tAi, tAj, tAx, sym)
    for _ in range(2):
        K.refactor(tAi, tAj, tAx, num, sym); K.solve_with_numeric(num, tb, sym)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        K.refactor(tAi, tAj, tAx, num, sym)
        x = K.solve_with_numeric(num, tb, sym)
    e1.record()
    torch.cuda.synchronize()
    ms_us = e0.elapsed_time(e1) / reps
    print(f"{which}: cuSOLVER-RF ResetValues + BatchRefactor + BatchSolve: {ms_rf:.2f} ms per step of {L} systems")
    print(f"{which}: klujax_b200 refactor + solve_with_numeric:            {ms_us:.2f} ms per step of {L} systems")


if __name__ == "__main__":
    main()
