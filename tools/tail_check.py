"""Builder tool: the register tail stage of the wide kernel on / off (cfg 5 shape): parity vs dense, timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klujax_b200 as K
from klujax_b200 import synth
n, L = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pat = synth.SaxPattern(n=n, seed=2)
Ax = pat.values_fast(L, seed=7)
rng = np.random.default_rng(3)
b = rng.standard_normal((L, n, 1)) + 1j * rng.standard_normal((L, n, 1))
Ai = torch.tensor(pat.Ai, device="cuda"); Aj = torch.tensor(pat.Aj, device="cuda")
Axd = torch.tensor(Ax, device="cuda"); bd = torch.tensor(b, device="cuda")
res = {}
for mode in ("1", "0"):
    os.environ["KLUJAX_CUDA_TAIL"] = mode
    sym = K.analyze(pat.Ai, pat.Aj, n)
    x = K.solve_with_symbol(Ai, Aj, Axd, bd, sym); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); x = K.solve_with_symbol(Ai, Aj, Axd, bd, sym, check=False); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    st, gr = K.last_status()
    res[mode] = (x.cpu().numpy(), min(ts), int(st.abs().max()), float(gr.max()))
err = 0.0
for q in range(min(L, 48)):
    A = synth.dense_from_coo(pat.Ai, pat.Aj, Ax[q], n)
    err = max(err, np.abs(res["1"][0][q] - np.linalg.solve(A, b[q])).max() / np.abs(np.linalg.solve(A, b[q])).max())
print(f"tail stage ON {res['1'][1]:.3f} ms (status {res['1'][2]}, growth {res['1'][3]:.3f}) | OFF {res['0'][1]:.3f} ms (growth {res['0'][3]:.3f})"
      f" | ON vs dense {err:.2e} | ON vs OFF {np.abs(res['1'][0]-res['0'][0]).max():.2e}")
