"""Writes the SAX-style test pattern of size n (klujax_b200.synth.SaxPattern, seed 2) as text for the
C++ planner tools: first line `n nz`, then `i j re im` per entry.  Usage: python tools/dump_sax_pattern.py 128"""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.makedirs('/tmp/pl', exist_ok=True)
from klujax_b200 import synth
n=int(sys.argv[1])
p=synth.SaxPattern(n, seed=2)
Ax=p.values(0,1)[0]
with open(f'/tmp/pl/sax{n}.txt','w') as f:
    f.write(f"{n} {len(p.Ai)}\n")
    for i,j,v in zip(p.Ai,p.Aj,Ax): f.write(f"{i} {j} {float(v.real)!r} {float(v.imag)!r}\n")
