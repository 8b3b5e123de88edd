"""Size-independent properties of an engine's state, vectorised (numpy), for polytopes too large to replay
through the CPU reference: the parity gate `bench.py` applies to the full cut sequence and the `-m gpu`
tests apply at 10^5..10^6 vertices.

Nothing here knows how the state was produced: the arrays are read through the reference's own struct layout
(`polytope.used/ideal/data/incidence/adjacence`, bslv_poly.h:55-69) after `b200_poly_materialise`, i.e. exactly
what `poly__polyck` (bslv_poly.c:940-990) walks.  The properties are SURVEY App. B's plus, for a simple polytope
(random tangent halfspaces), |inc(v)| = |adj(v)| = d and E = V*d/2.
"""
from __future__ import annotations

import ctypes as C
import hashlib

import numpy as np


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


def check_simple_polytope(s: Snapshot, sample: int = 20000, tol: float = 1e-7, seed: int = 0) -> dict:
    """Raises AssertionError on the first violated property; returns the counts it verified.

    Default-callback semantics (cone_polar, bslv_poly.c:30-39): facet f is the halfspace fdata[f].y >= -1."""
    d, live = s.d, s.live
    V = len(live)
    assert V > 0
    assert not s.ideal[live].any(), "a bounded polytope has no ideal vertices"
    # ---- every live vertex is simple: d facets, d neighbours
    assert (s.inc_len[live] == d).all(), "a live vertex does not lie on exactly d facets"
    assert (s.adj_len[live] == d).all(), "a live vertex does not have exactly d neighbours"
    inc = s.inc[(s.inc_off[live][:, None] + np.arange(d)[None, :])].astype(np.int64)      # [V, d] facet ids
    adj = s.adj[(s.adj_off[live][:, None] + np.arange(d)[None, :])].astype(np.int64)      # [V, d] slots
    assert s.used[adj].all(), "adjacency points at a dead slot"
    assert s.fused[inc].all(), "a live vertex lies on a dead facet"
    assert (np.diff(np.sort(inc, axis=1), axis=1) > 0).all(), "duplicate facet in an incidence list"
    # ---- adjacency is symmetric and loop-free: the multiset of (u,v) equals the multiset of (v,u)
    u = np.repeat(live, d)
    v = adj.reshape(-1)
    assert (u != v).all(), "self loop"
    key_uv = np.sort(u.astype(np.uint64) << np.uint64(32) | v.astype(np.uint64))
    key_vu = np.sort(v.astype(np.uint64) << np.uint64(32) | u.astype(np.uint64))
    assert (np.diff(key_uv) > 0).all(), "duplicate neighbour"
    assert (key_uv == key_vu).all(), "adjacency is not symmetric"
    E = len(u) // 2
    assert 2 * E == V * d
    # ---- neighbours share exactly d-1 facets (edge_test's necessary condition, bslv_poly.c:482-485; simple: exactly)
    slot2row = np.full(s.S, -1, np.int64)
    slot2row[live] = np.arange(V)
    inc_s = np.sort(inc, axis=1)
    bad = 0
    CH = 1 << 18
    for b in range(0, len(u), CH):
        a_inc = inc_s[slot2row[u[b:b + CH]]]
        b_inc = inc_s[slot2row[v[b:b + CH]]]
        shared = (a_inc[:, :, None] == b_inc[:, None, :]).sum(axis=(1, 2))
        bad += int((shared != d - 1).sum())
    assert bad == 0, f"{bad} adjacent pairs do not share exactly d-1 facets"
    # ---- no two live vertices carry the same facet set (a simple polytope's vertex is its facet set)
    order = np.lexsort(inc_s.T[::-1])
    srt = inc_s[order]
    assert (np.abs(np.diff(srt, axis=0)).sum(axis=1) > 0).all(), "two live vertices with the same incidence set"
    # ---- facet -> vertex lists are the transpose of the incidence lists (App. B)
    assert int(s.flen[s.fused].sum()) == V * d, "facet lists and incidence lists disagree in size"
    ff = np.repeat(np.arange(s.F), s.flen)
    within = np.arange(len(ff)) - np.repeat(np.cumsum(s.flen) - s.flen, s.flen)
    fv = s.fvert[np.repeat(s.foff, s.flen) + within]
    k1 = np.sort(ff.astype(np.uint64) << np.uint64(32) | fv.astype(np.uint64))
    k2 = np.sort(inc.reshape(-1).astype(np.uint64) << np.uint64(32) | np.repeat(live, d).astype(np.uint64))
    assert len(k1) == len(k2) and (k1 == k2).all(), "facet lists are not the transpose of the incidence lists"
    # ---- geometry: every vertex is tight on its own facets, and (sampled) feasible for all halfspaces
    x = s.data[live]
    worst = 0.0
    for j in range(d):
        t = np.einsum("ij,ij->i", s.fdata[inc[:, j]], x)
        worst = max(worst, float(np.abs(t + 1.0).max()))
    assert worst <= tol, f"a vertex is off one of its facets by {worst:.3e}"
    rng = np.random.default_rng(seed)
    pick = rng.choice(V, size=min(sample, V), replace=False)
    slack = x[pick] @ s.fdata[s.fused].T + 1.0               # (unused dual slots: redundant halfspaces, and the d queued
    #                                                           start halfspaces poly__intl_apprx retires and re-adds, bslv_poly.c:190-197)
    assert float(slack.min()) >= -tol, f"a vertex violates a halfspace by {float(slack.min()):.3e}"
    # a vertex lies on no other facet: exactly d of the live facets are tight
    tight = (np.abs(slack) <= 1e-9).sum(axis=1)
    assert (tight == d).all(), "a sampled vertex is tight on a halfspace missing from its incidence list"
    return {"vertices": int(V), "edges": int(E), "facets": int(s.fused.sum()), "max_facet_residual": worst,
            "feasibility_sample": int(len(pick))}


def canonical(s: Snapshot):
    """Slot-number independent form: live vertices ordered by their sorted facet tuple; adjacency in that numbering."""
    d, live = s.d, s.live
    V = len(live)
    inc = np.sort(s.inc[(s.inc_off[live][:, None] + np.arange(d)[None, :])].astype(np.int64), axis=1)
    order = np.lexsort(inc.T[::-1])
    rank = np.empty(V, np.int64)
    rank[order] = np.arange(V)
    slot2canon = np.full(s.S, -1, np.int64)
    slot2canon[live] = rank
    adj = slot2canon[s.adj[(s.adj_off[live][:, None] + np.arange(d)[None, :])].astype(np.int64)]
    adj = np.sort(adj, axis=1)
    return inc[order], s.data[live][order], adj[order]


def digest(s: Snapshot) -> str:
    """SHA-256 over the canonical incidence, coordinates (bit patterns) and adjacency: equal digests = same polytope."""
    inc, x, adj = canonical(s)
    h = hashlib.sha256()
    for arr in (inc, x.view(np.uint64), adj):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()
