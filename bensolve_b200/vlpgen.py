"""Synthetic vector linear programs in bensolve's .vlp format (BASELINE configs 3-4).

    min  P x   s.t.  B x >= -1,  -box <= x <= box          (ordering cone R^q_+, c = (1,...,1))

B: m x n, rows uniform on the sphere (so 0 is interior); P: the first q rows of a random
orthogonal n x n matrix.  Format facts used (bslv_vlp.c:275-588): 1-based indices; an unspecified
row is free, an unspecified column is FIXED AT 0, so every variable gets a `j` line.
"""
from __future__ import annotations

import numpy as np


def random_vlp(q: int, m: int, n: int, seed: int = 1, box: float = 10.0):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((m, n))
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    P = Q[:q]
    return B, P, box


def write_vlp(path: str, B: np.ndarray, P: np.ndarray, box: float) -> None:
    m, n = B.shape
    q = P.shape[0]
    with open(path, "w") as f:
        f.write(f"p vlp min {m} {n} {m * n} {q} {q * n}\n")
        for i in range(m):
            f.write("".join(f"a {i + 1} {j + 1} {B[i, j]:.17g}\n" for j in range(n)))
        for i in range(q):
            f.write("".join(f"o {i + 1} {j + 1} {P[i, j]:.17g}\n" for j in range(n)))
        for i in range(m):
            f.write(f"i {i + 1} l -1\n")
        for j in range(n):
            f.write(f"j {j + 1} d {-box:.17g} {box:.17g}\n")
        f.write("e\n")
