"""Regenerates tests/golden/known_answers.json and tests/golden/oracle_vectors.npz.

known_answers.json holds the reference's own published known-answer systems
(README.md:62-86, docs/examples/batched.md) — transcribed, because the reference
cannot be imported in this container (no jax / SuiteSparse).
oracle_vectors.npz pins the ORACLE itself: seeded inputs of the reference's test
shapes (/root/reference/tests.py:299-358 style) with the oracle's outputs, so a
later change to oracle/ that alters results is caught on the GPU box too, where
neither scipy cross-checks nor /root/reference are assumed.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from klujax_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    out = {}
    cases = [("f64_n5_L3_r2", 5, 15, 3, 2, np.float64), ("c128_n5_L3_r2", 5, 15, 3, 2, np.complex128),
             ("f64_n3_L1_r1", 3, 8, 1, 1, np.float64), ("c128_n20_L2_r3", 20, 60, 2, 3, np.complex128)]
    for name, n, nz, L, r, dt in cases:
        Ai, Aj, Ax, b = synth.ref_style_case(n, nz, L, r, dt)
        sym = O.analyze(Ai, Aj, n)
        x = O.solve_with_symbol(Ai, Aj, Ax, b, sym)
        xt = O.solve_with_symbol(Ai, Aj, Ax, b, sym, transpose=True)
        O.free_symbolic(sym)
        for k, v in dict(Ai=Ai, Aj=Aj, Ax=Ax, b=b, x=x, xt=xt).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
