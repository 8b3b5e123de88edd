"""Size-independent properties of an engine's state, vectorised (numpy), for polytopes too large to replay
through the CPU reference: the parity gate `bench.py` applies to the full cut sequence and the `-m gpu`
tests apply at 10^5..10^7 vertices.

Nothing here knows how the state was produced: the arrays are read through the reference's own struct layout
(`polytope.used/ideal/data/incidence/adjacence`, bslv_poly.h:55-69) after `b200_poly_materialise`, i.e. exactly
what `poly__polyck` (bslv_poly.c:940-990) walks.  The properties are SURVEY App. B's plus those of a bounded
polytope cut from random tangent halfspaces: every vertex lies on >= d facets and has >= d
neighbours, with equality except for the handful of vertices that came within the reference's 1e-9 of a later
hyperplane and were copied onto it (bslv_poly.c:573-588, :666-674) -- those are counted and reported.
"""
from __future__ import annotations

import ctypes as C
import hashlib

import numpy as np

SENTINEL = np.int64(1) << np.int64(62)


def _bits(words, cnt):
    nw = (cnt + 63) // 64
    if nw == 0:
        return np.zeros(0, bool)
    w = np.ctypeslib.as_array(words, shape=(nw,)).astype(np.uint64)
    return np.unpackbits(w.view(np.uint8), bitorder="little")[:cnt].astype(bool)


def _csr(lst_ptr, cnt):
    """poly_list[cnt] -> (lengths[cnt], flat uint64 array, offsets[cnt]) without a Python loop.  The product points
    every list into one slab (b200_poly_materialise), which is what makes the flat view possible."""
    if cnt == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.uint64), np.zeros(0, np.int64)
    raw = np.ctypeslib.as_array(C.cast(lst_ptr, C.POINTER(C.c_uint64)), shape=(cnt, 3))
    lens = raw[:, 0].astype(np.int64)
    ptrs = raw[:, 2].astype(np.uint64)
    nz = lens > 0
    if not nz.any():
        return lens, np.zeros(0, np.uint64), np.zeros(cnt, np.int64)
    base = int(ptrs[nz].min())
    off = np.zeros(cnt, np.int64)
    off[nz] = ((ptrs[nz] - np.uint64(base)) // np.uint64(8)).astype(np.int64)
    total = int((off[nz] + lens[nz]).max())
    if total != int(lens.sum()):
        raise ValueError("lists are not laid out in one slab (not the product's materialised layout)")
    flat = np.ctypeslib.as_array((C.c_uint64 * total).from_address(base)).copy()
    return lens, flat, off


def _flatten(flat, off, lens, idx):
    """(owner index into idx, position within the list, value) of every entry of the lists of `idx`."""
    ln = lens[idx]
    owner = np.repeat(np.arange(len(idx)), ln)
    within = np.arange(int(ln.sum())) - np.repeat(np.cumsum(ln) - ln, ln)
    vals = flat[np.repeat(off[idx], ln) + within].astype(np.int64)
    return owner, within, vals


class Snapshot:
    """Slot-indexed arrays of one engine (primal side + the facets' vertex lists)."""

    def __init__(self, engine):
        engine.materialise()
        a, d = engine.args, engine.dim
        P, D = a.primal, a.dual
        self.d = d
        self.S, self.F = int(P.cnt), int(D.cnt)
        self.used = _bits(P.used, self.S)
        self.ideal = _bits(P.ideal, self.S)
        self.data = np.ctypeslib.as_array(P.data, shape=(self.S * d,)).reshape(self.S, d).copy()
        self.inc_len, self.inc, self.inc_off = _csr(P.incidence, self.S)
        self.adj_len, self.adj, self.adj_off = _csr(P.adjacence, self.S)
        self.fused = _bits(D.used, self.F)
        self.fideal = _bits(D.ideal, self.F)
        self.fdata = np.ctypeslib.as_array(D.data, shape=(self.F * d,)).reshape(self.F, d).copy()
        self.flen, self.fvert, self.foff = _csr(D.incidence, self.F)
        self.live = np.nonzero(self.used)[0]


def _padded(owner, within, vals, n, width):
    m = np.full((n, width), SENTINEL, np.int64)
    m[owner, within] = vals
    m.sort(axis=1)
    return m


def check_polytope(s: Snapshot, sample: int = 20000, tol: float = 1e-7, seed: int = 0, max_edges: int = 12_000_000) -> dict:
    """Raises AssertionError on the first violated property; returns the counts it verified.

    Default-callback semantics (cone_polar, bslv_poly.c:30-39): facet f is the halfspace fdata[f].y >= -1."""
    d, live = s.d, s.live
    V = len(live)
    assert V > 0
    assert not s.ideal[live].any(), "a bounded polytope has no ideal vertices"
    il, al = s.inc_len[live], s.adj_len[live]
    assert (il >= d).all(), "a live vertex lies on fewer than d facets"
    assert (al >= d).all(), "a live vertex has fewer than d neighbours"
    simple = (il == d) & (al == d)
    n_deg = int((~simple).sum())
    assert n_deg <= max(8, V // 10000), f"{n_deg} of {V} vertices are not simple: more than on-plane coincidences explain"
    io, iw, iv = _flatten(s.inc, s.inc_off, s.inc_len, live)          # vertex row, position, facet
    ao, aw, av = _flatten(s.adj, s.adj_off, s.adj_len, live)          # vertex row, position, neighbour slot
    assert s.used[av].all(), "adjacency points at a dead slot"
    assert s.fused[iv].all(), "a live vertex lies on a dead facet"
    wi = int(il.max())
    inc_m = _padded(io, iw, iv, V, wi)                                # [V, wi] sorted facet ids, SENTINEL-padded
    real = inc_m[:, 1:] != SENTINEL
    assert ((inc_m[:, 1:] > inc_m[:, :-1]) | ~real).all(), "duplicate facet in an incidence list"
    # ---- adjacency is symmetric and loop-free: the set of (u,v) equals the set of (v,u)
    u = live[ao]
    v = av
    assert (u != v).all(), "self loop"
    key_uv = np.sort(u.astype(np.uint64) << np.uint64(32) | v.astype(np.uint64))
    key_vu = np.sort(v.astype(np.uint64) << np.uint64(32) | u.astype(np.uint64))
    assert (key_uv[1:] != key_uv[:-1]).all(), "duplicate neighbour"
    assert (key_uv == key_vu).all(), "adjacency is not symmetric"
    E = len(u) // 2
    # ---- neighbours share >= d-1 facets (edge_test's necessary condition, bslv_poly.c:482-485); exactly d-1 between
    # simple vertices
    slot2row = np.full(s.S, -1, np.int64)
    slot2row[live] = np.arange(V)
    ru, rv = ao, slot2row[v]
    if len(ru) > max_edges:               # beyond ~10^7 directed edges: a uniform sample of them
        pick_e = np.random.default_rng(seed + 1).choice(len(ru), size=max_edges, replace=False)
        ru, rv = ru[pick_e], rv[pick_e]
    bad = 0
    CH = 1 << 18
    for b in range(0, len(ru), CH):
        a_inc, b_inc = inc_m[ru[b:b + CH]], inc_m[rv[b:b + CH]]
        shared = ((a_inc[:, :, None] == b_inc[:, None, :]) & (a_inc[:, :, None] != SENTINEL)).sum(axis=(1, 2))
        both_simple = simple[ru[b:b + CH]] & simple[rv[b:b + CH]]
        bad += int(((shared != d - 1) & both_simple).sum()) + int((shared < d - 1).sum())
    assert bad == 0, f"{bad} adjacent pairs do not share d-1 facets"
    # ---- no two live vertices carry the same facet set
    order = np.lexsort(inc_m.T[::-1])
    srt = inc_m[order]
    assert ((srt[1:] != srt[:-1]).any(axis=1)).all(), "two live vertices with the same incidence set"
    # ---- facet -> vertex lists are the transpose of the incidence lists (App. B)
    assert int(s.flen[s.fused].sum()) == len(iv), "facet lists and incidence lists disagree in size"
    fo, _, fv = _flatten(s.fvert, s.foff, s.flen, np.arange(s.F))
    k1 = np.sort(fo.astype(np.uint64) << np.uint64(32) | fv.astype(np.uint64))
    k2 = np.sort(iv.astype(np.uint64) << np.uint64(32) | live[io].astype(np.uint64))
    assert len(k1) == len(k2) and (k1 == k2).all(), "facet lists are not the transpose of the incidence lists"
    # ---- geometry: every vertex is tight on its own facets, and (sampled) feasible for all halfspaces
    x = s.data[live]
    worst = 0.0
    for b in range(0, len(iv), 1 << 22):
        t = np.einsum("ij,ij->i", s.fdata[iv[b:b + (1 << 22)]], x[io[b:b + (1 << 22)]])
        worst = max(worst, float(np.abs(t + 1.0).max()))
    assert worst <= tol, f"a vertex is off one of its facets by {worst:.3e}"
    rng = np.random.default_rng(seed)
    pick = rng.choice(V, size=min(sample, V), replace=False)
    fl = np.nonzero(s.fused)[0]
    lo_cnt = np.zeros(len(pick), np.int64)
    hi_cnt = np.zeros(len(pick), np.int64)
    min_slack = np.inf
    for b in range(0, len(fl), 8192):        # (unused dual slots: redundant halfspaces, and the d queued start halfspaces
        slack = x[pick] @ s.fdata[fl[b:b + 8192]].T + 1.0          # poly__intl_apprx retires and re-adds, bslv_poly.c:190-197)
        min_slack = min(min_slack, float(slack.min()))
        lo_cnt += (np.abs(slack) <= 0.5e-9).sum(axis=1)
        hi_cnt += (np.abs(slack) <= 2e-9).sum(axis=1)
    assert min_slack >= -tol, f"a vertex violates a halfspace by {min_slack:.3e}"
    # a vertex lies on no other facet: the halfspaces tight within the reference's 1e-9 are its incidence list
    assert ((lo_cnt <= il[pick]) & (il[pick] <= hi_cnt)).all(), "a sampled vertex is tight on a halfspace missing from its incidence list"
    return {"vertices": int(V), "edges": int(E), "facets": int(s.fused.sum()), "non_simple_vertices": n_deg,
            "max_facet_residual": worst, "feasibility_sample": int(len(pick)), "shared_facet_check_edges": int(len(ru))}


check_simple_polytope = check_polytope


def canonical(s: Snapshot):
    """Slot-number independent form: live vertices ordered by their sorted facet tuple; adjacency in that numbering."""
    live = s.live
    V = len(live)
    io, iw, iv = _flatten(s.inc, s.inc_off, s.inc_len, live)
    inc_m = _padded(io, iw, iv, V, int(s.inc_len[live].max()))
    order = np.lexsort(inc_m.T[::-1])
    rank = np.empty(V, np.int64)
    rank[order] = np.arange(V)
    slot2canon = np.full(s.S, -1, np.int64)
    slot2canon[live] = rank
    ao, aw, av = _flatten(s.adj, s.adj_off, s.adj_len, live)
    adj_m = _padded(ao, aw, slot2canon[av], V, int(s.adj_len[live].max()))
    return inc_m[order], s.data[live][order], adj_m[order]


def digest(s: Snapshot) -> str:
    """SHA-256 over the canonical incidence, coordinates (bit patterns) and adjacency: equal digests = same polytope."""
    inc, x, adj = canonical(s)
    h = hashlib.sha256()
    for arr in (inc, x.view(np.uint64), adj):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()
