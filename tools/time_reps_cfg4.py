import sys, os, torch
sys.path.insert(0, "/root/repo")
import klujax_b200 as K
from klujax_b200 import synth
pat = synth.MnaPattern(n=20000)
n, L = pat.n, 128
Ai, Aj = torch.from_numpy(pat.Ai).cuda(), torch.from_numpy(pat.Aj).cuda()
Ax = torch.from_numpy(pat.values(0, L)).cuda()
b = torch.randn(L, n, 1, dtype=torch.float64, device="cuda")
sym = K.analyze(pat.Ai, pat.Aj, n)
num = K.factor(Ai, Aj, Ax, sym)
def each(f, reps=8):
    f(); torch.cuda.synchronize(); out=[]
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); out.append(round(e0.elapsed_time(e1),2))
    return out
print("solve ", each(lambda: K.solve_with_numeric(num, b, sym)))
print("tsolve", each(lambda: K.tsolve_with_numeric(num, b, sym)))
print("solve ", each(lambda: K.solve_with_numeric(num, b, sym)))
print("tsolve", each(lambda: K.tsolve_with_numeric(num, b, sym)))
print("refac ", each(lambda: K.refactor(Ai, Aj, Ax, num, sym)))
def step():
    K.refactor(Ai, Aj, Ax, num, sym); K.solve_with_numeric(num, b, sym); K.tsolve_with_numeric(num, b, sym)
print("step  ", each(step))
