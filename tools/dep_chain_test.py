"""Chain-only systems (bidiagonal / banded) through the level regime: time per dependency link of the
dependency-driven kernels against the level kernels (KLUJAX_CUDA_DEP=0/1).  Usage: python tools/dep_chain_test.py [band]"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("KLUJAX_CUDA_FORCE_REGIME", "1")
import klujax_b200 as K
band = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, L = 20000, int(os.environ.get("CHAIN_L", "1"))
rows, cols = [np.arange(n)], [np.arange(n)]
for d in range(1, band + 1):
    rows += [np.arange(d, n), np.arange(0, n - d)]; cols += [np.arange(0, n - d), np.arange(d, n)]
Ai = np.concatenate(rows).astype(np.int32); Aj = np.concatenate(cols).astype(np.int32)
rng = np.random.default_rng(0)
Ax = rng.uniform(-1, 1, (L, Ai.size)); Ax[:, :n] = 2.0 * band + 2.0
b = rng.standard_normal((L, n, 1))
sym = K.analyze(Ai, Aj, n)
tAi, tAj, tAx, tb = (torch.from_numpy(a).cuda() for a in (Ai, Aj, Ax, b))
num = K.factor(tAi, tAj, tAx, sym)
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_ref = timed(lambda: K.refactor(tAi, tAj, tAx, num, sym))
t_sol = timed(lambda: K.solve_with_numeric(num, tb, sym))
x = K.solve_with_numeric(num, tb, sym).cpu().numpy()
import scipy.sparse as sp
A0 = sp.csr_matrix((Ax[0], (Ai, Aj)), shape=(n, n))
res = np.abs(A0 @ x[0, :, 0] - b[0, :, 0]).max() / np.abs(b[0]).max()
from klujax_b200 import _cabi
info = _cabi.symbolic_info(int(sym.handle), False)
print(f"band {band} L {L} DEP={os.environ.get('KLUJAX_CUDA_DEP','1')}: refactor {t_ref:.3f} ms, solve {t_sol:.3f} ms, residual {res:.1e}, levels {info['levels']} nblocks {info['nblocks']}")
