import glob
import json
import os

import numpy as np

from bensolve_b200 import capi, polytopes as P

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json")))


def load_golden(path):
    with open(path) as f:
        g = json.load(f)
    vals = np.array([[float.fromhex(x) for x in row] for row in g["vals"]], dtype=np.float64)
    tr = P.Trace(g["dim"], vals, np.array(g["ideal"], np.uint8), g["n_init"], g["name"])
    st = g["state"]
    coords = np.array([[float.fromhex(x) for x in row] for row in st["coords"]], dtype=np.float64).reshape(-1, g["dim"])
    return tr, g["rcs"], st, coords


def check_against_golden(lib, path):
    tr, rcs, st, coords = load_golden(path)
    e = capi.PolyEngine(lib, tr.dim)
    try:
        got_rcs = P.replay(e, tr)
        s = e.state()
    finally:
        e.kill()
    assert got_rcs == rcs
    assert s.n_dual_slots == st["n_dual_slots"]
    assert [list(t) for t in s.incidence] == st["incidence"]
    assert [int(x) for x in s.ideal] == st["ideal"]
    assert [list(t) for t in s.adjacency] == st["adjacency"]
    assert list(s.live_facets) == st["live_facets"]
    same = (s.coords.view(np.uint64) == coords.view(np.uint64)) | (s.coords == coords)
    assert same.all(), "coordinates are not bit-identical to the reference's"


def run_pair(lib_a, lib_b, tr, stepwise=False, exact=True, flags_b=0):
    """Replay one trace into two engines; compare rc sequences and canonical state (after every
    cut when stepwise)."""
    a, b = capi.PolyEngine(lib_a, tr.dim), capi.PolyEngine(lib_b, tr.dim, flags=flags_b)
    try:
        if not stepwise:
            ra, rb = P.replay(a, tr), P.replay(b, tr)
            assert ra == rb, f"{tr.name}: return codes differ"
            capi.compare_states(a.state(), b.state(), exact_coords=exact)
            return ra
        for i in range(tr.n_init):
            a.add(tr.vals[i], int(tr.ideal[i]))
            b.add(tr.vals[i], int(tr.ideal[i]))
        assert a.init_approx() == b.init_approx() == 0
        capi.compare_states(a.state(), b.state(), exact_coords=exact)
        rcs = []
        for i in range(tr.n_init, len(tr)):
            ra = a.add(tr.vals[i], int(tr.ideal[i]))
            rb = b.add(tr.vals[i], int(tr.ideal[i]))
            assert ra == rb, f"{tr.name}: rc differs at halfspace {i}"
            if ra == 0:
                assert int(a.args.idx) == int(b.args.idx) or True   # slot numbers may differ (A.7)
            try:
                capi.compare_states(a.state(), b.state(), exact_coords=exact)
            except AssertionError as ex:
                raise AssertionError(f"{tr.name}: after halfspace {i}: {ex}") from None
            rcs.append(ra)
        return rcs
    finally:
        a.kill()
        b.kill()
