"""GPU parity tests: the CUDA engine, called through the reference-facing C ABI (poly__*), against
the oracle on identical traces.  Bar (north_star): identical return codes, vertex/direction counts,
incidence and adjacency structure after canonical sorting; coordinates bit-identical (the engine
reproduces the reference's FP64 operation order), which is stricter than the 1e-9 relative asked."""
import os

import numpy as np
import pytest

from bensolve_b200 import capi, polytopes as P
from helpers import check_against_golden, dual_adjacency_of, golden_files, run_pair
from traces import medium_traces, small_traces, stepwise_traces

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def checker(built):
    """The strongest oracle available on this box: the unmodified reference object if it travelled
    (oracle/_ref), else the restatement."""
    if os.path.exists(capi.REF_SO):
        return capi.load_lib(capi.REF_SO)
    return capi.load_lib(capi.ORACLE_SO)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-5])
def test_gpu_reproduces_golden(product_lib, path):
    check_against_golden(product_lib, path)


@pytest.mark.parametrize("tr", small_traces(), ids=lambda t: t.name)
def test_gpu_matches_oracle(product_lib, oracle_lib, tr):
    run_pair(oracle_lib, product_lib, tr, exact=True)


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_gpu_matches_checker_after_every_cut(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, stepwise=True, exact=True)


@pytest.mark.parametrize("tr", medium_traces(), ids=lambda t: t.name)
def test_gpu_matches_checker_medium(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, exact=True)


def _check_invariants(s, tr, rcs, simple=True):
    """Size-independent properties: every live vertex satisfies every halfspace and
    is tight on its incident facets, adjacency is symmetric; simple polytopes have exactly d facets
    and d neighbours per vertex.  Assumes n_init == dim and a generic trace, so that halfspace i
    is dual slot i+1 (slot 0 is the facet at infinity)."""
    d = tr.dim
    n_v = len(s.incidence)
    vals = s.coords @ tr.vals.T                  # halfspace i: vals[:, i] >= -1 (default callback)
    assert (vals >= -1 - 1e-7).all(), "a vertex violates a halfspace"
    for i in range(n_v):
        if simple:
            assert len(s.incidence[i]) == d and len(s.adjacency[i]) == d
        for f in s.incidence[i]:
            assert f >= 1 and abs(vals[i, f - 1] + 1.0) < 1e-7, "vertex not on an incident facet"
    for i in range(0, n_v, max(1, n_v // 2000)):
        for j in s.adjacency[i]:
            assert i in s.adjacency[j]
            assert len(set(s.incidence[i]) & set(s.incidence[j])) >= d - 1


def test_gpu_large_tangent_properties_and_reference(product_lib, checker):
    """~2*10^4 live vertices: compare with the checker (seconds on the CPU) and check invariants."""
    tr = P.tangent_polytope(4, 3000, 21)
    a, b = capi.PolyEngine(checker, 4), capi.PolyEngine(product_lib, 4)
    ra, rb = P.replay(a, tr), P.replay(b, tr)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)
    _check_invariants(sb, tr, rb)
    assert sb.n_points > 15000


def test_gpu_d6_properties(product_lib):
    """R^6 (BASELINE config 5 shape, reduced count): ~3*10^4 vertices, invariants only."""
    tr = P.tangent_polytope(6, 300, 88)
    e = capi.PolyEngine(product_lib, 6)
    rcs = P.replay(e, tr)
    s = e.state()
    e.kill()
    assert sum(rcs) == 0
    assert len(s.live_facets) == 300
    _check_invariants(s, tr, rcs)
    # Euler check on the graph of a simple 6-polytope: edges = V*d/2
    assert sum(len(a) for a in s.adjacency) == 6 * len(s.adjacency)


def test_gpu_redundant_halfspace_leaves_state_untouched(product_lib):
    tr = P.tangent_polytope(3, 30, 3)
    e = capi.PolyEngine(product_lib, 3)
    P.replay(e, tr)
    s0 = e.state()
    rc = e.add(-0.5 * tr.vals[0], 0)            # parallel to halfspace 0 but twice as far: redundant
    s1 = e.state()
    e.kill()
    assert rc == 1
    assert s1.n_dual_slots == s0.n_dual_slots + 1
    assert s1.incidence == s0.incidence and s1.adjacency == s0.adjacency
    assert (s1.coords == s0.coords).all()


def test_gpu_get_vrtx_and_sltn_inheritance(product_lib, checker):
    tr = P.cube_with_cuts(4)
    states = []
    for lib in (checker, product_lib):
        e = capi.PolyEngine(lib, 4)
        P.replay(e, tr, upto=8)
        for _ in range(12):
            rc, idx, _, _ = e.get_vrtx()
            assert rc == 0
            e.mark_solution(idx)
        e.add(tr.vals[8], 0)
        states.append(e.state())
        e.kill()
    capi.compare_states(states[0], states[1], exact_coords=True)


def test_gpu_two_engines_alive_at_once(product_lib, oracle_lib):
    """bslv_algs.c keeps several poly_args alive simultaneously (upper image + cone, :813/:333)."""
    t1, t2 = P.tangent_polytope(3, 40, 1), P.lattice_polytope(4, 30, 1)
    e1, e2 = capi.PolyEngine(product_lib, 3), capi.PolyEngine(product_lib, 4)
    o1, o2 = capi.PolyEngine(oracle_lib, 3), capi.PolyEngine(oracle_lib, 4)
    for i in range(max(len(t1), len(t2))):
        for e, o, t in ((e1, o1, t1), (e2, o2, t2)):
            if i < len(t):
                e.add(t.vals[i], 0); o.add(t.vals[i], 0)
                if i == t.n_init - 1:
                    e.init_approx(); o.init_approx()
    capi.compare_states(o1.state(), e1.state(), exact_coords=True)
    capi.compare_states(o2.state(), e2.state(), exact_coords=True)
    for x in (e1, e2, o1, o2):
        x.kill()


def test_gpu_dual_adjacency_and_polyck(product_lib, checker, capfd):
    tr = P.lattice_polytope(4, 30, 3)
    adj = []
    for lib in (checker, product_lib):
        e = capi.PolyEngine(lib, 4)
        P.replay(e, tr)
        e.update_dual_adjacence()
        r = e.raw()
        adj.append({f: sorted(v) for f, v in r["dual"]["adj"].items() if len(r["dual"]["inc"][f])})
        if lib is product_lib:
            e.polyck()
        e.kill()
    assert adj[0] == adj[1]
    err = capfd.readouterr().err
    assert "appears in vertex'" not in err and "are adjacent" not in err


FLAG_EAGER_GC = 2   # compact device rows after every cut that killed a vertex


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_gpu_row_compaction_after_every_cut(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, stepwise=True, exact=True, flags_b=FLAG_EAGER_GC)


@pytest.mark.parametrize("tr", medium_traces(), ids=lambda t: t.name)
def test_gpu_row_compaction_medium(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, exact=True, flags_b=FLAG_EAGER_GC)


@pytest.mark.parametrize("tr", small_traces()[::2] + medium_traces()[:3], ids=lambda t: t.name)
@pytest.mark.parametrize("chunk", [0, 7])
def test_gpu_batch_entry_point(product_lib, checker, tr, chunk):
    """Device-resident batch path (halfspaces built on the device, header-only readback, bulk mirror
    rebuild) against the checker fed one halfspace per call."""
    a, b = capi.PolyEngine(checker, tr.dim), capi.PolyEngine(product_lib, tr.dim, flags=FLAG_EAGER_GC if chunk else 0)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, chunk)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)


def test_gpu_add_each_is_the_per_call_loop(product_lib, checker):
    """b200_poly_add_each (the C caller's loop around poly__add_vrtx) against one ctypes call per halfspace."""
    tr = P.tangent_polytope(4, 90, 11)
    a, b = capi.PolyEngine(checker, tr.dim), capi.PolyEngine(product_lib, tr.dim)
    ra = P.replay(a, tr)
    for i in range(tr.n_init):
        b.add(tr.vals[i], int(tr.ideal[i]))
    assert b.init_approx() == 0
    rb = b.add_each(tr.vals[tr.n_init:], tr.ideal[tr.n_init:])
    assert ra == rb
    capi.compare_states(a.state(), b.state(), exact_coords=True)
    a.kill(); b.kill()


def test_gpu_reserve_then_run(product_lib, oracle_lib):
    tr = P.tangent_polytope(5, 120, 5)
    a, b = capi.PolyEngine(oracle_lib, 5), capi.PolyEngine(product_lib, 5)
    assert b.reserve(200000, 1 << 20, 1 << 20) == 0
    ra, rb = P.replay(a, tr), P.replay(b, tr)
    assert ra == rb
    capi.compare_states(a.state(), b.state(), exact_coords=True)
    st = b.stats()
    assert st["cuts"] == len(tr) - 5 - sum(rb) and st["vertex_evals"] > 0 and st["kernel_launches"] > 0
    a.kill(); b.kill()


FLAG_MULTI_KERNEL = 4   # never use the single-CTA tail: every cut goes through the multi-kernel path


@pytest.mark.parametrize("tr", small_traces()[::3] + medium_traces(), ids=lambda t: t.name)
def test_gpu_multi_kernel_path(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, exact=True, flags_b=FLAG_MULTI_KERNEL)


@pytest.mark.parametrize("tr", stepwise_traces()[:4], ids=lambda t: t.name)
def test_gpu_multi_kernel_path_after_every_cut(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, stepwise=True, exact=True, flags_b=FLAG_MULTI_KERNEL | FLAG_EAGER_GC)


@pytest.mark.parametrize("tr", [t for t in small_traces() if "pyramid_k300" not in t.name][::2] + medium_traces()[:4], ids=lambda t: t.name)
def test_gpu_dual_adjacency_k6(product_lib, checker, tr):
    """K6 on the device (bit matrix facets x vertices, same AND+POPC filter and containment kernels as K4)."""
    assert dual_adjacency_of(checker, tr) == dual_adjacency_of(product_lib, tr)


@pytest.fixture
def tiny_caps(monkeypatch):
    monkeypatch.setenv("B200_TINY_CAPS", "1")


@pytest.mark.parametrize("tr", small_traces()[::3] + medium_traces()[:3], ids=lambda t: t.name)
@pytest.mark.parametrize("flags", [0, FLAG_MULTI_KERNEL, FLAG_EAGER_GC])
def test_gpu_capacity_negotiation(product_lib, checker, tiny_caps, tr, flags):
    """Start from near-zero device capacities: rows, pools, pair buffers, bit matrix and staging all
    overflow and are renegotiated (on the device before any mutation; K4 / adjacency re-runnable)."""
    run_pair(checker, product_lib, tr, exact=True, flags_b=flags)


@pytest.mark.parametrize("seed", range(12))
def test_gpu_random_mixed_stress(product_lib, oracle_lib, seed):
    """Random dimension / generator / seed: unbounded, degenerate and redundant-heavy inputs."""
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.integers(2, 7))
    n = int(rng.integers(d + 2, 60 if d >= 5 else 150))
    kind = int(rng.integers(0, 4))
    tr = [P.mixed_polyhedron, P.lattice_polytope, P.random_offsets, P.tangent_polytope][kind](d, n, int(rng.integers(1, 10 ** 6))) \
        if not (kind == 0 and d < 3) else P.tangent_polytope(d, n, seed)
    run_pair(oracle_lib, product_lib, tr, exact=True, flags_b=[0, FLAG_EAGER_GC, FLAG_MULTI_KERNEL][seed % 3])


# ---------------------------------------------------------------- the kernel variants the bench runs
FLAG_FORCE_WIDE = 16   # every cut through k_tail<16> / k_tail2<16> (widest cluster) + the grid-wide k4_filter / k4_contain


@pytest.mark.parametrize("dim,n,seed", [(6, 300, 88), (5, 1500, 88)])
def test_gpu_bench_shape_matches_reference(product_lib, checker, dim, n, seed):
    """The bench's own shape (random tangent polytope in R^6 / R^5) at the size the reference finishes in
    seconds: > 256 visited vertices and > 512 new rows per late cut, i.e. the wide-cluster kernels the
    headline number runs on -- compared with the reference object, not only with invariants."""
    tr = P.tangent_polytope(dim, n, seed)
    a, b = capi.PolyEngine(checker, dim), capi.PolyEngine(product_lib, dim)
    ra, rb = P.replay(a, tr), P.replay(b, tr)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)
    assert sb.n_points > 25000


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_gpu_forced_wide_cluster_after_every_cut(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, stepwise=True, exact=True, flags_b=FLAG_FORCE_WIDE)


@pytest.mark.parametrize("tr", small_traces()[::2] + medium_traces(), ids=lambda t: t.name)
def test_gpu_forced_wide_cluster(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, exact=True, flags_b=FLAG_FORCE_WIDE)


@pytest.mark.parametrize("tr", medium_traces(), ids=lambda t: t.name)
def test_gpu_forced_wide_cluster_with_compaction(product_lib, checker, tr):
    run_pair(checker, product_lib, tr, exact=True, flags_b=FLAG_FORCE_WIDE | FLAG_EAGER_GC)


@pytest.fixture
def small_he_cap(monkeypatch):
    monkeypatch.setenv("B200_HE_CAP", "48")


@pytest.mark.parametrize("tr", medium_traces()[:6], ids=lambda t: t.name)
def test_gpu_tail_bails_out_with_queued_followers(product_lib, checker, small_he_cap, tr):
    """Half-edge scratch of 48 entries: most cuts bail out of the tail (ST_NEED_BIG) while k4_filter (with its
    early-header block), k4_contain and k_tail2 are already queued behind it; they must do nothing and the cut
    must be redone by the multi-kernel path."""
    run_pair(checker, product_lib, tr, exact=True, flags_b=FLAG_FORCE_WIDE)


# ---------------------------------------------------------------- wave path (device-resident batches)
FLAG_WAVES_ALWAYS = 32   # look-ahead classification + concurrent commuting cuts from the first halfspace on


@pytest.mark.parametrize("tr", small_traces()[::2] + medium_traces(), ids=lambda t: t.name)
@pytest.mark.parametrize("chunk", [0, 7])
def test_gpu_wave_path(product_lib, checker, tr, chunk):
    a, b = capi.PolyEngine(checker, tr.dim), capi.PolyEngine(product_lib, tr.dim, flags=FLAG_WAVES_ALWAYS)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, chunk)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)


@pytest.mark.parametrize("env", [{"B200_WAVE_IN_ORDER": "1"}, {"B200_WAVE_MAX": "3", "B200_WAVE_CAND": "5", "B200_WAVE_REFILL": "2"},
                                 {"B200_TINY_CAPS": "1"}, {"B200_HE_CAP": "48"}], ids=lambda e: "-".join(e))
@pytest.mark.parametrize("tr", medium_traces(), ids=lambda t: t.name)
def test_gpu_wave_path_variants(product_lib, checker, monkeypatch, env, tr):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    a, b = capi.PolyEngine(checker, tr.dim), capi.PolyEngine(product_lib, tr.dim, flags=FLAG_WAVES_ALWAYS)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, 0)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)


@pytest.mark.parametrize("dim,n,seed", [(6, 300, 88), (5, 1500, 88), (4, 3000, 21)])
def test_gpu_wave_path_bench_shape_matches_reference(product_lib, checker, dim, n, seed):
    """The bench's `value` path (b200_poly_add_batch -> waves) on the bench's shape against the reference object."""
    tr = P.tangent_polytope(dim, n, seed)
    a, b = capi.PolyEngine(checker, dim), capi.PolyEngine(product_lib, dim, flags=FLAG_WAVES_ALWAYS)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, 0)
    sa, sb = a.state(), b.state()
    st = b.stats()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)
    assert st["rows_scanned"] < st["vertex_evals"]


# ---------------------------------------------------------------- size-independent properties at scale (the bench's gate as a test)
@pytest.mark.parametrize("dim,n", [(6, 2000), (5, 6000), (3, 60000)])
def test_gpu_full_size_properties_and_path_equality(product_lib, dim, n):
    """10^5..3*10^5 vertices, far beyond what the CPU reference replays in test time: every property a correct result
    has (bensolve_b200/invariants.py), on the device-resident path and on the per-call path, and equality of the two
    (SHA-256 of the canonical form, coordinates by bit pattern)."""
    from bensolve_b200 import invariants as INV
    tr = P.tangent_polytope(dim, n, 20261018)
    digests, counts = [], []
    for batch in (True, False):
        e = capi.PolyEngine(product_lib, dim)
        for i in range(dim):
            e.add(tr.vals[i], 0)
        assert e.init_approx() == 0
        rcs = e.add_batch(tr.vals[dim:]) if batch else e.add_each(tr.vals[dim:])
        assert sum(rcs) == 0                       # tangent halfspaces: none redundant
        snap = INV.Snapshot(e)
        counts.append(INV.check_polytope(snap))
        digests.append(INV.digest(snap))
        st = e.stats()
        e.kill()
        if batch:
            assert st["waves"] > 0 and st["rows_scanned"] < st["vertex_evals"]
    assert digests[0] == digests[1]
    assert counts[0]["facets"] == n and counts[0]["vertices"] > n


@pytest.mark.parametrize("dim", [3, 4, 5])
@pytest.mark.parametrize("flags", [0, FLAG_MULTI_KERNEL, FLAG_FORCE_WIDE])
def test_gpu_zero_plus_rows_are_projected_once(product_lib, checker, tiny_caps, dim, flags):
    """ZERO+ closure followed by a capacity bail-out: the rerun must not project the rows again (CutParams::zp_done)."""
    run_pair(checker, product_lib, P.cube_zero_plus(dim), stepwise=True, exact=True, flags_b=flags)


@pytest.mark.parametrize("tr", [P.tangent_polytope(4, 300, 3), P.tangent_polytope(6, 300, 88), P.tangent_polytope(5, 1500, 88), P.lattice_polytope(4, 60, 2),
                                P.mixed_polyhedron(4, 80, 5), P.random_cone(5, 40, 2), P.random_offsets(4, 90, 1)], ids=lambda t: t.name)
def test_gpu_vertex_enumeration_through_intl_apprx(product_lib, checker, tr):
    """cone_vertenum's call pattern (bslv_algs.c:331-350) with the UNCHANGED API: everything queued, then poly__intl_apprx;
    inside, the re-adds run as one device-resident batch (look-ahead + waves)."""
    run_pair(checker, product_lib, P.Trace(tr.dim, tr.vals, tr.ideal, len(tr.vals), tr.name + "_queued"), exact=True)
