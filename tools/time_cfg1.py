"""Config 1 (the reference's own CPU-runnable case: n = 1000, one system, `solve`): per-call latency."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import klujax_b200 as K
from klujax_b200 import synth

Ai, Aj, Ax = synth.random_dd()
n = 1000
b = np.random.default_rng(1).standard_normal((1, n, 1))
Ax_d, b_d = torch.from_numpy(Ax).cuda(), torch.from_numpy(b).cuda()
def t(f, reps):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("solve (analyze + seed + schedule + kernels, every call): %.1f ms" % t(lambda: K.solve(Ai, Aj, Ax_d, b_d), 3))
sym = K.analyze(Ai, Aj, n)
print("solve_with_symbol (handle reused):                      %.2f ms" % t(lambda: K.solve_with_symbol(Ai, Aj, Ax_d, b_d, sym), 10))
num = K.factor(Ai, Aj, Ax_d, sym)
print("solve_with_numeric (factors reused):                    %.2f ms" % t(lambda: K.solve_with_numeric(num, b_d, sym), 10))
# (The CPU figures quoted next to these in DESIGN.md come from the oracle, which only tests/,
# smoke() and bench.py's CPU baseline may execute: `python bench.py --impl reference` style runs.)
