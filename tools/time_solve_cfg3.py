"""Times solve_with_numeric of config 3 (256^2 Poisson, 64 systems x 256 right-hand sides).
Arguments: NAME=VALUE environment settings to compare, e.g. KLUJAX_CUDA_LEVEL_C=1 KLUJAX_CUDA_LEVEL_C=2."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klujax_b200 as K
from klujax_b200 import synth

g, L, r = 256, 64, 256
Ai_h, Aj_h, Ax_h = synth.poisson2d(g, L)
n = g * g
Ai, Aj, Ax = torch.from_numpy(Ai_h).cuda(), torch.from_numpy(Aj_h).cuda(), torch.from_numpy(Ax_h).cuda()
b = torch.randn(L, n, r, dtype=torch.float64, device="cuda")
sym = K.analyze(Ai_h, Aj_h, n)
num = K.factor(Ai, Aj, Ax, sym)
ref = None
for setting in (sys.argv[1:] or ["X=0"]):
    name, val = setting.split("=")
    os.environ[name] = val
    x = K.solve_with_numeric(num, b, sym); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2): x = K.solve_with_numeric(num, b, sym)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = x.clone()
    # residual of system 0, first 4 columns
    import numpy as np, scipy.sparse as sp
    A0 = sp.coo_matrix((Ax_h[0], (Ai_h, Aj_h)), shape=(n, n)).tocsr()
    xs = x[0, :, :4].cpu().numpy(); bs = b[0, :, :4].cpu().numpy()
    res = np.abs(A0 @ xs - bs).max() / np.abs(bs).max()
    print("%s: solve %.1f ms | residual %.1e | vs first setting %.1e" % (setting, e0.elapsed_time(e1) / 2, res, float((x - ref).abs().max())), flush=True)
    del os.environ[name]
