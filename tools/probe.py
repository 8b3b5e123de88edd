"""GPU probe: replay one synthetic trace through the C ABI and print timing + engine statistics."""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bensolve_b200 import capi, polytopes as P  # noqa: E402


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("cuts", "redundant", "vertex_evals", "rows_scanned", "minus", "zero", "zpp", "edge_vertices",
                                          "copies", "pair_tests", "new_adjacent_pairs", "algorithmic_bytes", "kernel_launches",
                                          "compactions", "live_vertices", "slots", "facets")] + [("classify_ms", C.c_double), ("cut_ms", C.c_double)]


def get_stats(lib, eng):
    st = Stats()
    lib.b200_poly_get_stats.argtypes = [C.POINTER(capi.PolyArgs), C.POINTER(Stats)]
    lib.b200_poly_get_stats(C.byref(eng.args), C.byref(st))
    return {n: getattr(st, n) for n, _ in Stats._fields_}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--timing", type=int, default=1)
    ap.add_argument("--every", type=int, default=0)
    ap.add_argument("--reserve", type=int, default=0, help="rows to pre-size (0 = grow on demand)")
    a = ap.parse_args()
    lib = capi.load_product()
    tr = P.tangent_polytope(a.dim, a.n, a.seed)
    e = capi.PolyEngine(lib, a.dim)
    lib.b200_poly_set_flags.argtypes = [C.POINTER(capi.PolyArgs), C.c_uint]
    lib.b200_poly_set_flags(C.byref(e.args), a.timing)
    if a.reserve:
        e.reserve(a.reserve, a.reserve * (a.dim + 2), a.reserve * (a.dim + 2))
    t0 = time.perf_counter()
    last = [t0, 0.0, 0.0]

    def on_cut(i, rc):
        if a.every and (i + 1) % a.every == 0:
            st = get_stats(lib, e)
            now = time.perf_counter()
            print(json.dumps(dict(i=i + 1, live=st["live_vertices"], slots=st["slots"], wall_ms_per_cut=1e3 * (now - last[0]) / a.every,
                                  gpu_cut_ms=(st["cut_ms"] - last[1]) / a.every, classify_ms=(st["classify_ms"] - last[2]) / a.every)), flush=True)
            last[0], last[1], last[2] = now, st["cut_ms"], st["classify_ms"]
    P.replay(e, tr, on_cut=on_cut)
    dt = time.perf_counter() - t0
    st = get_stats(lib, e)
    st.update(wall_s=dt, cuts_per_s=(st["cuts"]) / dt, dim=a.dim, n=a.n)
    print(json.dumps(st))
    e.kill()


if __name__ == "__main__":
    main()
