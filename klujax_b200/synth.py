"""Seeded synthetic workload generators for the BASELINE.json configs.

cfg1: random diagonally dominant f64, n=1000, nnz~5000, L=1          (klujax.solve)
cfg2: SAX-style photonic circuit, c128, n=64,  L=4096               (solve_with_symbol)
cfg3: 2-D 5-point Poisson 256^2, f64, L=64, r=256                   (factor + solve_with_numeric)
cfg4: MNA-like with planted BTF structure, f64, n=20000, L=128      (refactor per step + tsolve)
cfg5: cfg2's generator at n=128, L=65536                            (solve_with_symbol, sharded)

All COO outputs are coalesced and lexsorted by (row, col) the way
klujax.coalesce leaves them (/root/reference/klujax.py:153-168).
numpy only; no reference code is imported.
"""
from __future__ import annotations

import numpy as np


def _coalesce_sorted(Ai, Aj):
    order = np.lexsort((Aj, Ai))
    return Ai[order].astype(np.int32), Aj[order].astype(np.int32), order


# --------------------------------------------------------------------------- cfg1
def random_dd(n=1000, n_off=4000, seed=0, dtype=np.float64, n_lhs=1):
    """Random sparse, strictly row-diagonally-dominant matrix (cfg1)."""
    r = np.random.default_rng(seed)
    i = r.integers(0, n, n_off)
    j = r.integers(0, n, n_off)
    keep = i != j
    i, j = i[keep], j[keep]
    key = np.unique(i.astype(np.int64) * n + j)
    i, j = (key // n).astype(np.int32), (key % n).astype(np.int32)
    nz_off = i.size
    Ai = np.concatenate([i, np.arange(n, dtype=np.int32)])
    Aj = np.concatenate([j, np.arange(n, dtype=np.int32)])
    Ai, Aj, order = _coalesce_sorted(Ai, Aj)
    cplx = np.issubdtype(dtype, np.complexfloating)
    Ax = np.empty((n_lhs, Ai.size), dtype)
    for m in range(n_lhs):
        v = r.standard_normal(nz_off)
        if cplx:
            v = v + 1j * r.standard_normal(nz_off)
        rowsum = np.zeros(n)
        np.add.at(rowsum, i, np.abs(v))
        full = np.concatenate([v, rowsum + 1.0]).astype(dtype)
        Ax[m] = full[order]
    return Ai, Aj, Ax


# ----------------------------------------------------------------------- cfg2 / cfg5
class SaxPattern:
    """A = I - C.S for S block-diagonal (ports x ports dense blocks) and C a
    perfect matching on all but n_ext ports (generator G2)."""

    def __init__(self, n=64, ports=4, n_ext=4, seed=1):
        assert n % ports == 0 and (n - n_ext) % 2 == 0
        self.n, self.ports = n, ports
        r = np.random.default_rng(seed)
        perm = r.permutation(n - n_ext)
        pairs = perm.reshape(-1, 2)
        partner = -np.ones(n, np.int64)
        partner[pairs[:, 0]] = pairs[:, 1]
        partner[pairs[:, 1]] = pairs[:, 0]
        self.partner = partner
        # (C.S)[i, j] = S[partner(i), j]  -> nonzero iff j in block(partner(i))
        rows, cols, sr, sc = [], [], [], []
        for i in range(n):
            p = partner[i]
            if p < 0:
                continue
            blk = p // ports
            for q in range(ports):
                j = blk * ports + q
                rows.append(i); cols.append(j); sr.append(p); sc.append(j)
        rows = np.array(rows); cols = np.array(cols)
        sr = np.array(sr); sc = np.array(sc)
        # add the identity; merge where (i, i) already present
        key_off = rows.astype(np.int64) * n + cols
        diag_key = np.arange(n, dtype=np.int64) * (n + 1)
        allkey = np.unique(np.concatenate([key_off, diag_key]))
        self.Ai = (allkey // n).astype(np.int32)
        self.Aj = (allkey % n).astype(np.int32)
        self._pos_off = np.searchsorted(allkey, key_off)
        self._pos_diag = np.searchsorted(allkey, diag_key)
        self._s_idx = (sr // ports, sr % ports, sc % ports)
        self.nnz = allkey.size

    def values(self, m0, count, scale=0.1):
        """Ax[count, nnz] complex128; system m uses rng(100 + m)."""
        nb = self.n // self.ports
        Ax = np.zeros((count, self.nnz), np.complex128)
        for t in range(count):
            r = np.random.default_rng(100 + m0 + t)
            S = scale * (r.standard_normal((nb, self.ports, self.ports))
                         + 1j * r.standard_normal((nb, self.ports, self.ports)))
            Ax[t, self._pos_diag] = 1.0
            np.add.at(Ax[t], self._pos_off, -S[self._s_idx])
        return Ax

    def values_fast(self, count, seed=12345, scale=0.1):
        """Same distribution, one vectorised draw (for the big bench batches)."""
        nb = self.n // self.ports
        r = np.random.default_rng(seed)
        S = scale * (r.standard_normal((count, nb, self.ports, self.ports))
                     + 1j * r.standard_normal((count, nb, self.ports, self.ports)))
        Ax = np.zeros((count, self.nnz), np.complex128)
        Ax[:, self._pos_diag] = 1.0
        vals = -S[:, self._s_idx[0], self._s_idx[1], self._s_idx[2]]
        # _pos_off may repeat the diagonal position at most once per entry: use add.at per column
        np.add.at(Ax, (slice(None), self._pos_off), vals)
        return Ax


# --------------------------------------------------------------------------- cfg3
def poisson2d(g=256, n_lhs=64, dtype=np.float64):
    """5-point Laplacian on a g x g grid plus a per-system diagonal shift."""
    n = g * g
    idx = np.arange(n).reshape(g, g)
    rows = [idx.ravel()]; cols = [idx.ravel()]; vals = [np.full(n, 4.0)]
    for a, b in ((idx[:, :-1], idx[:, 1:]), (idx[:-1, :], idx[1:, :])):
        rows += [a.ravel(), b.ravel()]
        cols += [b.ravel(), a.ravel()]
        vals += [np.full(a.size, -1.0), np.full(a.size, -1.0)]
    Ai = np.concatenate(rows); Aj = np.concatenate(cols); v = np.concatenate(vals)
    Ai, Aj, order = _coalesce_sorted(Ai, Aj)
    v = v[order]
    diag = Ai == Aj
    Ax = np.tile(v.astype(dtype), (n_lhs, 1))
    Ax[:, diag] += (1e-3 * np.arange(n_lhs))[:, None]
    return Ai, Aj, Ax


# --------------------------------------------------------------------------- cfg4
class MnaPattern:
    """MNA-like circuit matrix with planted block-triangular structure hidden by
    random row/column permutations (generator G4)."""

    def __init__(self, n=20000, seed=3, big_frac=0.6, symmetric_hide=True):
        # symmetric_hide=True (benchmark default): rows and columns get the SAME
        # random permutation, so the dominant diagonal stays a zero-free diagonal
        # and KLU's maxtrans cheap assignment keeps it (diagonal pivots, no
        # growth).  With independent permutations the matching is a different
        # transversal and KLU's tol=1e-3 threshold pivoting accepts small pivots
        # (oracle residual ~5e-7): kept only as a structural BTF stress case.
        r = np.random.default_rng(seed)
        sizes = [int(big_frac * n)]
        tot = sizes[0]
        while tot < n:
            s = 1 if r.random() < 0.7 else int(r.integers(2, 40))
            s = min(s, n - tot)
            sizes.append(s); tot += s
        offs = np.concatenate([[0], np.cumsum(sizes)])
        rows, cols, vals = [], [], []
        nb = len(sizes)
        for bi, s in enumerate(sizes):
            o = offs[bi]
            if s > 1:
                k = np.arange(s)
                rows.append(o + k); cols.append(o + (k + 1) % s); vals.append(-r.random(s))
                a = r.integers(0, s, 2 * s)
                c = np.clip(a + r.integers(-30, 31, 2 * s), 0, s - 1)
                g = -r.random(2 * s)
                rows += [o + a, o + c]; cols += [o + c, o + a]; vals += [g, g]
            if bi < nb - 1:
                ncpl = max(1, s // 2)
                rr = o + r.integers(0, s, ncpl)
                cc = r.integers(offs[bi + 1], n, ncpl)
                rows.append(rr); cols.append(cc); vals.append(-r.random(ncpl))
        i = np.concatenate(rows); j = np.concatenate(cols); v = np.concatenate(vals)
        keep = i != j
        i, j, v = i[keep], j[keep], v[keep]
        key = i.astype(np.int64) * n + j
        ukey, inv = np.unique(key, return_inverse=True)
        uv = np.zeros(ukey.size); np.add.at(uv, inv, v)
        i = (ukey // n); j = (ukey % n)
        p = r.permutation(n); q = r.permutation(n)
        if symmetric_hide:
            q = p
        # A <- A[p][:, q]  : new row a = position of old row in p
        pinv = np.empty(n, np.int64); pinv[p] = np.arange(n)
        qinv = np.empty(n, np.int64); qinv[q] = np.arange(n)
        self.n = n
        self.sizes = sizes
        self._off_row_old = i
        self._off_val = uv
        d = np.arange(n)
        Ai = np.concatenate([pinv[i], pinv[d]])
        Aj = np.concatenate([qinv[j], qinv[d]])
        self.Ai, self.Aj, self._order = _coalesce_sorted(Ai, Aj)
        self.nnz = self.Ai.size

    def values(self, t, n_lhs, dtype=np.float64):
        """Ax[n_lhs, nnz] at transient step t: off-diagonals scaled by
        1 + 0.1 sin(t + m); diagonal = sum |row| + 1e-3."""
        n = self.n
        Ax = np.empty((n_lhs, self.nnz), dtype)
        rowabs = np.zeros(n)
        np.add.at(rowabs, self._off_row_old, np.abs(self._off_val))
        for m in range(n_lhs):
            s = 1.0 + 0.1 * np.sin(t + m)
            full = np.concatenate([self._off_val * s, rowabs * s + 1e-3])
            Ax[m] = full[self._order]
        return Ax


# ------------------------------------------------------- reference test-suite style
def ref_style_case(n_col, n_nz, n_lhs, n_rhs, dtype, seed=33):
    """Small random COO + 10.0 diagonal, coalesced — the shape of the fixtures in
    /root/reference/tests.py:299-358 (numpy RNG instead of jax PRNGKey)."""
    r = np.random.default_rng(seed)
    cplx = np.issubdtype(np.dtype(dtype), np.complexfloating)
    Ai = np.concatenate([r.integers(0, n_col, n_nz), np.arange(n_col)])
    Aj = np.concatenate([r.integers(0, n_col, n_nz), np.arange(n_col)])
    key = Ai.astype(np.int64) * n_col + Aj
    ukey, inv = np.unique(key, return_inverse=True)
    Ax = np.zeros((n_lhs, ukey.size), dtype)
    for m in range(n_lhs):
        v = r.standard_normal(n_nz)
        if cplx:
            v = v + 1j * r.standard_normal(n_nz)
        full = np.concatenate([v, np.full(n_col, 10.0)])
        np.add.at(Ax[m], inv, full.astype(dtype))
    b = r.standard_normal((n_lhs, n_col, n_rhs))
    if cplx:
        b = b + 1j * r.standard_normal((n_lhs, n_col, n_rhs))
    return ((ukey // n_col).astype(np.int32), (ukey % n_col).astype(np.int32),
            Ax, b.astype(dtype))


def dense_from_coo(Ai, Aj, Ax_m, n):
    A = np.zeros((n, n), Ax_m.dtype)
    np.add.at(A, (Ai, Aj), Ax_m)
    return A
