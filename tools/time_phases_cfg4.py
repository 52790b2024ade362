"""Times refactor / solve_with_numeric / tsolve_with_numeric of config 4 separately (CUDA events)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klujax_b200 as K
from klujax_b200 import synth, _cabi

pat = synth.MnaPattern(n=20000)
n, L = pat.n, 128
Ai, Aj = torch.from_numpy(pat.Ai).cuda(), torch.from_numpy(pat.Aj).cuda()
Ax = torch.from_numpy(pat.values(0, L)).cuda()
b = torch.randn(L, n, 1, dtype=torch.float64, device="cuda")
sym = K.analyze(pat.Ai, pat.Aj, n)
num = K.factor(Ai, Aj, Ax, sym)
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("refactor %.2f ms" % timed(lambda: K.refactor(Ai, Aj, Ax, num, sym, check=False) if "check" in K.refactor.__code__.co_varnames else K.refactor(Ai, Aj, Ax, num, sym)))
print("solve    %.2f ms" % timed(lambda: K.solve_with_numeric(num, b, sym)))
print("tsolve   %.2f ms" % timed(lambda: K.tsolve_with_numeric(num, b, sym)))
print(_cabi.symbolic_info(int(sym.handle), False))
