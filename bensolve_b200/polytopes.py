"""Synthetic halfspace traces for the cut path (BASELINE.json configs 3-5).

A *trace* is the sequence of calls entering ``poly__add_vrtx`` recorded at the ``val`` level
(SURVEY section 4): ``dim``, then per call ``val[dim]`` and the ``ideal`` flag, plus the position of
the ``poly__intl_apprx`` call.  With the default callback (cone_polar, bslv_poly.c:30-39) a dual
point ``d`` means the halfspace ``d.y >= -1`` (``ideal=0``) or ``d.y >= 0`` (``ideal=1``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Trace:
    dim: int
    vals: np.ndarray      # [n, dim] float64 dual points
    ideal: np.ndarray     # [n] uint8
    n_init: int           # number of calls issued before poly__intl_apprx (>= dim)
    name: str = ""

    def __len__(self):
        return len(self.vals)


def tangent_polytope(dim: int, n: int, seed: int = 1) -> Trace:
    """Config 5: n halfspaces a.y <= 1 tangent to the unit ball, normals uniform on S^{dim-1}.

    Every halfspace is irredundant and the polytope is simple with probability 1.  The first
    ``dim+1`` normals are a regular-ish simplex so that the start polyhedron becomes bounded early."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, dim))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    return Trace(dim, -a, np.zeros(n, np.uint8), dim, f"tangent_d{dim}_n{n}_s{seed}")


def random_offsets(dim: int, n: int, seed: int = 1, spread: float = 0.5) -> Trace:
    """Random normals with offsets in [1, 1+spread]: many halfspaces are redundant (return code 1)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, dim))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = 1.0 + spread * rng.random(n)
    # a.y <= b  <=>  (-a/b).y >= -1
    return Trace(dim, -a / b[:, None], np.zeros(n, np.uint8), dim, f"offsets_d{dim}_n{n}_s{seed}")


def cube(dim: int) -> Trace:
    """+-e_i . y <= 1 (SURVEY App. C cube probes)."""
    vals = np.concatenate([-np.eye(dim), np.eye(dim)])
    return Trace(dim, vals, np.zeros(2 * dim, np.uint8), dim, f"cube_d{dim}")


def cube_with_cuts(dim: int) -> Trace:
    """Cube, then sum(y) <= dim-2 (passes through dim on-plane vertices: ZERO copies) and the
    supporting halfspace sum(y) <= dim (redundant, rc 1)."""
    t = cube(dim)
    a1 = np.ones(dim) / (dim - 2.0) if dim > 2 else np.ones(dim) / 0.5
    a2 = np.ones(dim) / float(dim)
    vals = np.concatenate([t.vals, -a1[None], -a2[None]])
    return Trace(dim, vals, np.zeros(len(vals), np.uint8), dim, f"cubecuts_d{dim}")


def cube_zero_plus(dim: int, gap: float = 5e-10) -> Trace:
    """Cube, then a halfspace that cuts off the vertex (1,...,1) and passes `gap` ABOVE its dim
    neighbours: they fall in the reference's projection band (1e-11 < s <= 1e-9, bslv_poly.c:666-674),
    are moved onto the hyperplane in place and then copied like on-plane vertices."""
    t = cube(dim)
    d = -(1.0 - gap) / (dim - 2.0) * np.ones(dim)      # neighbours of (1,..,1) have coordinate sum dim-2
    vals = np.concatenate([t.vals, d[None]])
    return Trace(dim, vals, np.zeros(len(vals), np.uint8), dim, f"cubezp_d{dim}")


def pyramid(k: int, tilt: float = 0.5) -> Trace:
    """R^3 pyramid over a regular k-gon: its apex lies on k facets (highly degenerate vertex).  The last
    halfspace passes exactly through the apex and cuts the base, so the apex is an on-plane vertex whose
    copy inherits a subset of k incidences (bslv_poly.c:634-652)."""
    th = 2 * np.pi * (np.arange(k) + 0.5) / k
    sides = np.stack([np.cos(th), np.sin(th), np.ones(k)], axis=1)     # n.y <= 1, tight at the apex (0,0,1)
    base = np.array([[0.0, 0.0, -1.0]])                                 # z >= -1
    cut = np.array([[tilt, 0.0, 1.0]])                                  # through the apex, cuts the base
    vals = np.concatenate([-sides[:3], -base, -sides[3:], -cut])
    return Trace(3, vals, np.zeros(len(vals), np.uint8), 3, f"pyramid_k{k}")


def lattice_polytope(dim: int, n: int, seed: int = 1, kmax: int = 2, bmax: int = 3) -> Trace:
    """Degenerate companion: small-integer normals and offsets, so many
    vertices lie exactly on later hyperplanes (ZERO copies, reduced incidence, ghost facets)."""
    rng = np.random.default_rng(seed)
    rows = [(-np.eye(dim)[i]) for i in range(dim)] + [np.eye(dim)[i] for i in range(dim)]
    rows = [r / float(bmax) for r in rows]          # box |y_i| <= bmax
    while len(rows) < n:
        a = rng.integers(-kmax, kmax + 1, size=dim).astype(np.float64)
        if not a.any():
            continue
        b = float(rng.integers(1, bmax * kmax + 1))
        rows.append(-a / b)                          # a.y <= b
    vals = np.asarray(rows[:n])
    return Trace(dim, vals, np.zeros(len(vals), np.uint8), dim, f"lattice_d{dim}_n{n}_s{seed}")


def random_cone(dim: int, n: int, seed: int = 1) -> Trace:
    """cone_vertenum template (bslv_algs.c:331-350): every generator is an ideal dual point, i.e. a
    homogeneous halfspace g.y >= 0; generators are drawn around e_dim so the cone is pointed."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, dim)) * 0.6
    g[:, -1] = np.abs(g[:, -1]) + 1.0
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    return Trace(dim, g, np.ones(n, np.uint8), n, f"cone_d{dim}_n{n}_s{seed}")


def mixed_polyhedron(dim: int, n: int, seed: int = 1) -> Trace:
    """Unbounded polyhedron in the style of an upper image: a few ideal generators (recession cone
    constraints) plus tangent halfspaces, interleaved."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, dim))
    a[:, -1] = np.abs(a[:, -1]) + 0.2            # every normal has a positive last component
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    ideal = (rng.random(n) < 0.25).astype(np.uint8)
    ideal[:dim] = 0
    vals = a.copy()                                # a.y >= -1  (or >= 0 when ideal)
    return Trace(dim, vals, ideal, dim, f"mixed_d{dim}_n{n}_s{seed}")


def replay(engine, trace: Trace, upto: int | None = None, on_cut=None):
    """Feed a trace into a PolyEngine the way bslv_algs.c does: queue n_init points,
    poly__intl_apprx, then one poly__add_vrtx per remaining point.  Returns the list of return
    codes (0 = cut, 1 = redundant) of the post-init calls."""
    n = len(trace) if upto is None else min(upto, len(trace))
    rcs = []
    for i in range(min(trace.n_init, n)):
        engine.add(trace.vals[i], int(trace.ideal[i]))
    if n < trace.n_init:
        return rcs
    rc = engine.init_approx()
    if rc:
        raise RuntimeError(f"poly__intl_apprx failed on trace {trace.name}")
    if on_cut is not None:
        on_cut(trace.n_init - 1, 0)
    for i in range(trace.n_init, n):
        rc = engine.add(trace.vals[i], int(trace.ideal[i]))
        rcs.append(rc)
        if on_cut is not None:
            on_cut(i, rc)
    return rcs


def replay_batched(engine, trace: Trace, chunk: int = 0):
    """Same trace through the batch entry point b200_poly_add_batch (device-resident runs of
    `chunk` halfspaces; 0 = everything after initialisation in one call)."""
    for i in range(trace.n_init):
        engine.add(trace.vals[i], int(trace.ideal[i]))
    if engine.init_approx():
        raise RuntimeError(f"poly__intl_apprx failed on trace {trace.name}")
    rest = np.arange(trace.n_init, len(trace))
    rcs = []
    step = chunk if chunk > 0 else max(1, len(rest))
    for s in range(0, len(rest), step):
        idx = rest[s:s + step]
        rcs += engine.add_batch(trace.vals[idx], trace.ideal[idx])
    return rcs
