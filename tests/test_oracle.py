"""CPU tests of the oracle: the restatement (oracle/poly_oracle.c) against the unmodified reference
object (oracle/_ref) and against the committed golden fixtures generated from that object."""
import ctypes as C

import numpy as np
import pytest

from bensolve_b200 import capi, polytopes as P
from helpers import check_against_golden, golden_files, run_pair
from traces import medium_traces, small_traces, stepwise_traces


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-5])
def test_oracle_reproduces_golden(oracle_lib, path):
    check_against_golden(oracle_lib, path)


@pytest.mark.parametrize("path", golden_files()[:4], ids=lambda p: p.split("/")[-1][:-5])
def test_reference_reproduces_golden(ref_lib, path):
    check_against_golden(ref_lib, path)


@pytest.mark.parametrize("tr", small_traces(), ids=lambda t: t.name)
def test_oracle_matches_reference(ref_lib, oracle_lib, tr):
    run_pair(ref_lib, oracle_lib, tr, exact=True)


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_oracle_matches_reference_after_every_cut(ref_lib, oracle_lib, tr):
    run_pair(ref_lib, oracle_lib, tr, stepwise=True, exact=True)


@pytest.mark.parametrize("tr", medium_traces()[:5], ids=lambda t: t.name)
def test_oracle_matches_reference_medium(ref_lib, oracle_lib, tr):
    run_pair(ref_lib, oracle_lib, tr, exact=True)


def test_cube_known_answers(oracle_lib):
    """SURVEY App. C cube probes: the cut sum(y) <= d-2 through d on-plane vertices yields d copies
    with d+1 incidences each, all mutually adjacent; the supporting halfspace sum(y) <= d is redundant."""
    for d in (3, 4, 5):
        tr = P.cube_with_cuts(d)
        e = capi.PolyEngine(oracle_lib, d)
        rcs = P.replay(e, tr)
        s = e.state()
        e.kill()
        assert rcs[-1] == 1 and rcs[-2] == 0
        f = 2 * d + 1                      # dual slot of the cut sum(y) <= d-2 (slot 0 = facet at infinity)
        on = [i for i, inc in enumerate(s.incidence) if f in inc]
        assert len(on) == d
        for i in on:
            assert len(s.incidence[i]) == d + 1
            assert set(on) - {i} <= set(s.adjacency[i])
        assert s.n_points == 2 ** d - 1 and s.n_dirs == 0


def test_start_simplex_layout(oracle_lib):
    """after poly__intl_apprx slot 0 is the point, slots 1..d directions, d+1 facets."""
    tr = P.tangent_polytope(4, 4, 3)
    e = capi.PolyEngine(oracle_lib, 4)
    P.replay(e, tr)
    r = e.raw()
    e.kill()
    assert r["primal"]["cnt"] == 5 and list(r["primal"]["ideal"]) == [0, 1, 1, 1, 1]
    assert all(len(r["primal"]["inc"][s]) == 4 and len(r["primal"]["adj"][s]) == 4 for s in range(5))


def test_reachability_equals_flat_classification(oracle_lib):
    """The CUDA engine classifies every vertex instead of walking the graph from the first violated
    one.  The oracle walks; this checks on every cut of the suite that the walk
    reaches every non-PLUS vertex, i.e. that both traversals visit the same set."""
    class Stats(C.Structure):
        _fields_ = [(n, C.c_size_t) for n in ("n_minus", "n_zero", "n_zp", "n_edge", "n_copies", "n_unreached", "n_pairs", "n_adj")]
    oracle_lib.oracle_last_cut_stats.restype = C.POINTER(Stats)
    for tr in small_traces():
        e = capi.PolyEngine(oracle_lib, tr.dim)
        unreached = []
        P.replay(e, tr, on_cut=lambda i, rc: unreached.append(oracle_lib.oracle_last_cut_stats().contents.n_unreached))
        e.kill()
        assert sum(unreached) == 0, tr.name


def test_get_vrtx_order(ref_lib, oracle_lib):
    tr = P.tangent_polytope(3, 12, 2)
    for lib in (ref_lib, oracle_lib):
        e = capi.PolyEngine(lib, 3)
        P.replay(e, tr)
        seen = []
        while True:
            rc, idx, ideal, val = e.get_vrtx()
            if rc:
                break
            seen.append(idx)
            e.mark_solution(idx)
        live = np.nonzero(e.raw()["primal"]["used"])[0].tolist()
        e.kill()
        assert seen == live
