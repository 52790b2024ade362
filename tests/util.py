import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with open(os.path.join(HERE, "golden", "known_answers.json")) as f:
        return json.load(f)


def relerr(x, y):
    x, y = np.asarray(x), np.asarray(y)
    d = np.abs(y).max()
    return float(np.abs(x - y).max() / (d if d > 0 else 1.0))


def dense(Ai, Aj, Ax_m, n):
    A = np.zeros((n, n), Ax_m.dtype)
    np.add.at(A, (Ai, Aj), Ax_m)
    return A


def residual(Ai, Aj, Ax_m, x_m, b_m, n, transpose=False):
    """max over right-hand sides of ||A x - b|| / ||b|| using the COO entries."""
    r = np.zeros_like(b_m)
    if transpose:
        np.add.at(r, Aj, Ax_m[:, None] * x_m[Ai])
    else:
        np.add.at(r, Ai, Ax_m[:, None] * x_m[Aj])
    num = np.linalg.norm(r - b_m, axis=0)
    den = np.linalg.norm(b_m, axis=0)
    return float((num / np.where(den > 0, den, 1.0)).max())
