"""Compare two sets of bensolve result files (<base>_{img,adj,inc}_{p,d}.sol) up to the things the
reference itself leaves order-dependent: row order (= slot numbering) and the order of
entries within a row.  Coordinates must agree to `tol` relative.  Exit code 0 = same solution."""
from __future__ import annotations

import sys

import numpy as np


def load_img(path):
    rows = [l.split() for l in open(path) if l.strip()]
    return [(int(r[0]), tuple(float(x) for x in r[1:])) for r in rows]


def load_idx(path):
    return [tuple(int(x) for x in l.split()) for l in open(path)]


def canon_order(img, digits=9):
    key = lambda i: (img[i][0], tuple(round(x, digits) + 0.0 for x in img[i][1]))
    order = sorted(range(len(img)), key=key)
    inv = {old: new for new, old in enumerate(order)}
    return order, inv


def compare(base_a, base_b, tol=1e-9):
    out = []
    maps = {}
    for side in ("p", "d"):
        ia, ib = load_img(f"{base_a}_img_{side}.sol"), load_img(f"{base_b}_img_{side}.sol")
        if len(ia) != len(ib):
            return [f"img_{side}: {len(ia)} vs {len(ib)} rows"]
        oa, inva = canon_order(ia)
        ob, invb = canon_order(ib)
        for ra, rb in zip(oa, ob):
            if ia[ra][0] != ib[rb][0]:
                return [f"img_{side}: point/direction flags differ"]
            x, y = np.array(ia[ra][1]), np.array(ib[rb][1])
            err = np.abs(x - y) / np.maximum(1.0, np.maximum(np.abs(x), np.abs(y)))
            if err.size and err.max() > tol:
                return [f"img_{side}: coordinates differ by {err.max():.3e}"]
        maps[side] = (inva, invb, oa, ob)
    for side, other in (("p", "d"), ("d", "p")):
        inva, invb, oa, ob = maps[side]
        aa, ab = load_idx(f"{base_a}_adj_{side}.sol"), load_idx(f"{base_b}_adj_{side}.sol")
        ca = {inva[i]: frozenset(inva[x] for x in row) for i, row in enumerate(aa)}
        cb = {invb[i]: frozenset(invb[x] for x in row) for i, row in enumerate(ab)}
        if ca != cb:
            out.append(f"adj_{side}: adjacency differs")
        # inc_<side>: one row per facet (= row of the OTHER side's image), listing vertices of <side>
        fa_inv, fb_inv = maps[other][0], maps[other][1]
        na, nb = load_idx(f"{base_a}_inc_{side}.sol"), load_idx(f"{base_b}_inc_{side}.sol")
        da = {fa_inv[f]: frozenset(inva[x] for x in row) for f, row in enumerate(na)}
        db = {fb_inv[f]: frozenset(invb[x] for x in row) for f, row in enumerate(nb)}
        if da != db:
            out.append(f"inc_{side}: incidence differs")
    return out


if __name__ == "__main__":
    diffs = compare(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1e-9)
    print("SAME" if not diffs else "DIFFERENT: " + "; ".join(diffs))
    sys.exit(1 if diffs else 0)
