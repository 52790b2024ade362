"""Times solve_with_numeric of config 3 (256^2 Poisson, 64 systems x 256 right-hand sides)."""
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
for threads in sys.argv[1:]:
    os.environ["KLUJAX_CUDA_LEVEL_THREADS"] = threads
    K.solve_with_numeric(num, b, sym); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2): K.solve_with_numeric(num, b, sym)
    e1.record(); torch.cuda.synchronize()
    print("threads %s: solve %.1f ms" % (threads, e0.elapsed_time(e1) / 2), flush=True)
