"""GPU probe of the wave path: one device-resident batch (b200_poly_add_batch) over a synthetic tangent polytope."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bensolve_b200 import capi, polytopes as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--n", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--reserve", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    lib = capi.load_product()
    tr = P.tangent_polytope(a.dim, a.n, a.seed)
    for rep in range(a.repeat):
        e = capi.PolyEngine(lib, a.dim)
        if a.reserve:
            e.reserve(a.reserve, a.reserve * (a.dim + 2), a.reserve * (a.dim + 2))
        for i in range(a.dim):
            e.add(tr.vals[i], 0)
        assert e.init_approx() == 0
        t0 = time.perf_counter()
        rcs = e.add_batch(tr.vals[a.dim:])
        dt = time.perf_counter() - t0
        st = e.stats()
        st.update(wall_s=dt, cuts_per_s=st["cuts"] / dt, dim=a.dim, n=a.n)
        print(json.dumps(st), flush=True)
        e.kill()


if __name__ == "__main__":
    main()
