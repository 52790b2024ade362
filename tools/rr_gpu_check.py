"""Quick GPU check of the register-resident kernel against the shared-memory kernels and a dense
solve: cfg 2 / cfg 5 shapes, timing of both paths.  (Builder tool; the parity tests proper live in
tests/test_gpu_parity.py.)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klujax_b200 as K
from klujax_b200 import synth

def run(n, L, r, seed, reps=5):
    pat = synth.SaxPattern(n=n, seed=seed)
    Ax = pat.values_fast(L, seed=7)
    rng = np.random.default_rng(3)
    b = rng.standard_normal((L, n, r)) + 1j * rng.standard_normal((L, n, r))
    Ai = torch.tensor(pat.Ai, device="cuda"); Aj = torch.tensor(pat.Aj, device="cuda")
    Axd = torch.tensor(Ax, device="cuda"); bd = torch.tensor(b, device="cuda")
    res = {}
    for mode in ("1", "0"):
        os.environ["KLUJAX_CUDA_ROWREG"] = mode
        sym = K.analyze(pat.Ai, pat.Aj, n)
        t0 = time.time()
        x = K.solve_with_symbol(Ai, Aj, Axd, bd, sym)
        torch.cuda.synchronize()
        t_first = time.time() - t0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            e0.record(); x = K.solve_with_symbol(Ai, Aj, Axd, bd, sym, check=False); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        st, gr = K.last_status()
        res[mode] = (x.cpu().numpy(), min(ts), t_first, int(st.abs().max()), float(gr.max()))
    x1, x0 = res["1"][0], res["0"][0]
    m = min(L, 64)
    err = 0.0
    for q in range(m):
        A = synth.dense_from_coo(pat.Ai, pat.Aj, Ax[q], n)
        xr = np.linalg.solve(A, b[q])
        err = max(err, np.abs(x1[q] - xr).max() / np.abs(xr).max())
    print(f"n={n} L={L} r={r}: rowreg {res['1'][1]:.3f} ms (first call {res['1'][2]:.1f} s, status {res['1'][3]}, growth {res['1'][4]:.3f})"
          f" | shared-memory kernels {res['0'][1]:.3f} ms (growth {res['0'][4]:.3f}) | max rel err vs dense {err:.2e}"
          f" | rowreg vs smem {np.abs(x1 - x0).max():.2e}", flush=True)

if __name__ == "__main__":
    import klujax_b200._cabi as cabi
    os.environ.setdefault("KLUJAX_CUDA_RR_VERBOSE", "1")
    if len(sys.argv) > 1:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 2 if int(sys.argv[1]) == 128 else 1)
    else:
        run(64, 4096, 1, 1)
        run(64, 4096, 4, 1)
        run(128, 8192, 1, 2)
        run(128, 65536, 1, 2)
