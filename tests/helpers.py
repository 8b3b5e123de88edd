import glob
import json
import os

import numpy as np

from bensolve_b200 import capi, polytopes as P

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json")) if not os.path.basename(p).startswith("benson_"))


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


# ---------------------------------------------------------------- Benson traces (real bensolve runs)
def benson_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "benson_*.json")))


def load_benson(path):
    with open(path) as f:
        return json.load(f)


def replay_benson_instance(lib, inst, flags=0):
    """Replay one recorded poly_args instance: same val sequence, the caller's callback replaced by a
    lookup of the halfspace it produced in the recorded run.  Returns (engine, rcs)."""
    import ctypes as C
    d = inst["dim"]
    table = {}
    for ev in inst["events"]:
        if ev[0] == "add":
            val = tuple(float.fromhex(x) for x in ev[1])
            table[(val, int(ev[2]))] = [float.fromhex(x) for x in ev[3]]

    def callback(dual_point, is_dir, hp_out):
        key = (tuple(dual_point[k] for k in range(d)), int(is_dir))
        hp = table[key]
        for k in range(d + 1):
            hp_out[k] = hp[k]

    e = capi.PolyEngine(lib, d, callback=callback, flags=flags)
    rcs = []
    for ev in inst["events"]:
        if ev[0] == "add":
            rc = e.add([float.fromhex(x) for x in ev[1]], int(ev[2]))
            rcs.append(rc)
            assert rc == ev[4], f"poly__add_vrtx returned {rc}, the reference run returned {ev[4]}"
        else:
            # dual slot 0 as the caller left it before poly__intl_apprx (bslv_algs.c:338-339)
            if ev[1]:
                e.args.dual.ideal[0] |= 1
            else:
                e.args.dual.ideal[0] &= ~1
            for k in range(d):
                e.args.dual.data[k] = float.fromhex(ev[2][k])
            rc = e.init_approx()
            assert rc == ev[3]
    return e, rcs


def check_benson_fixture(lib, path, checker=None, flags=0):
    fx = load_benson(path)
    for inst in fx["instances"]:
        e, _ = replay_benson_instance(lib, inst, flags)
        s = e.state()
        fin = inst.get("final")
        if fin:
            assert (s.n_points, s.n_dirs, s.n_slots, s.n_dual_slots) == (fin["points"], fin["dirs"], fin["slots"], fin["dual_slots"]), \
                f"{fx['example']}: counts {(s.n_points, s.n_dirs, s.n_slots, s.n_dual_slots)} != reference run {fin}"
        if checker is not None:
            c, _ = replay_benson_instance(checker, inst)
            capi.compare_states(c.state(), s, exact_coords=True)
            c.kill()
        e.kill()


def dual_adjacency_of(lib, tr, flags=0):
    """poly__update_adjacence(&dual) after replaying a trace: {facet: sorted neighbour facets} over the
    facets that hold at least one live vertex (the reference's ghost facets are dropped)."""
    e = capi.PolyEngine(lib, tr.dim, flags=flags)
    try:
        P.replay(e, tr)
        e.update_dual_adjacence()
        r = e.raw()
        real = {f for f, v in r["dual"]["inc"].items() if len(v)}
        return {f: sorted(x for x in r["dual"]["adj"][f] if x in real) for f in real}
    finally:
        e.kill()
