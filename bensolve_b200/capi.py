"""ctypes mirror of the reference's polyhedron-engine ABI (bslv_poly.h:49-118).

The same struct layouts and entry points are exported by three shared objects:

* ``bensolve_b200/libbslv_poly_b200.so`` -- the product: CUDA cut engine behind ``poly__*``
* ``oracle/_ref/libref_poly.so``         -- the unmodified reference ``bslv_poly.c`` (test oracle)
* ``oracle/libpoly_oracle.so``           -- our CPU restatement of the cut path (test oracle)

so one ``PolyEngine`` wrapper drives all of them.  Only tests, ``smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may load anything under ``oracle/``;
the product loader (:func:`load_product`) never does and raises if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
PRODUCT_SO = os.path.join(HERE, "libbslv_poly_b200.so")
REF_SO = os.path.join(REPO, "oracle", "_ref", "libref_poly.so")
ORACLE_SO = os.path.join(REPO, "oracle", "libpoly_oracle.so")

BTCNT = 64  # bslv_poly.h:42  (CHAR_BIT*sizeof(size_t) on LP64)


class PolyList(C.Structure):  # bslv_poly.h:49-53, 24 bytes
    _fields_ = [("cnt", C.c_size_t), ("blcks", C.c_size_t), ("data", C.POINTER(C.c_size_t))]


class Polytope(C.Structure):  # bslv_poly.h:55-69, 112 bytes
    pass


V2H = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double))

Polytope._fields_ = [
    ("dim", C.c_size_t),
    ("dim_primg", C.c_size_t),
    ("cnt", C.c_size_t),
    ("blcks", C.c_size_t),
    ("ip", C.POINTER(C.c_double)),
    ("data", C.POINTER(C.c_double)),
    ("data_primg", C.POINTER(C.c_double)),
    ("adjacence", C.POINTER(PolyList)),
    ("incidence", C.POINTER(PolyList)),
    ("ideal", C.POINTER(C.c_size_t)),
    ("used", C.POINTER(C.c_size_t)),
    ("sltn", C.POINTER(C.c_size_t)),
    ("dual", C.POINTER(Polytope)),
    ("v2h", C.c_void_p),
]


class InitData(C.Structure):  # bslv_poly.h:81
    _fields_ = [
        ("H", C.POINTER(C.c_double)),
        ("R", C.POINTER(C.c_double)),
        ("alph", C.POINTER(C.c_double)),
        ("queue", PolyList),
        ("gnrtrs", PolyList),
        ("intlsd", C.c_uint, 1),
    ]


class PolyArgs(C.Structure):  # bslv_poly.h:71-82, 392 bytes
    _fields_ = [
        ("dim", C.c_size_t),
        ("dim_primg_prml", C.c_size_t),
        ("dim_primg_dl", C.c_size_t),
        ("ideal", C.c_uint, 1),
        ("idx", C.c_size_t),
        ("val", C.POINTER(C.c_double)),
        ("val_primg_prml", C.POINTER(C.c_double)),
        ("val_primg_dl", C.POINTER(C.c_double)),
        ("eps", C.c_double),
        ("primal", Polytope),
        ("dual", Polytope),
        ("primalV2dualH", C.c_void_p),
        ("dualV2primalH", C.c_void_p),
        ("init_data", InitData),
    ]


class Permutation(C.Structure):  # bslv_poly.h:84-88
    _fields_ = [("cnt", C.c_size_t), ("data", C.POINTER(C.c_size_t)), ("inv", C.POINTER(C.c_size_t))]


class Stats(C.Structure):  # b200_stats, include/bensolve_b200.h
    _fields_ = [(n, C.c_uint64) for n in (
        "cuts", "redundant", "vertex_evals", "rows_scanned", "minus", "zero", "zero_plus_projected", "edge_vertices",
        "copies", "pair_tests", "new_adjacent_pairs", "algorithmic_bytes", "kernel_launches", "compactions",
        "live_vertices", "slots", "facets")] + [("classify_ms", C.c_double), ("cut_ms", C.c_double)] + [(n, C.c_uint64) for n in (
        "waves", "wave_cuts", "lookahead_passes", "sharded_passes", "sharded_cuts", "sharded_pair_tests")]


assert C.sizeof(PolyList) == 24 and C.sizeof(Polytope) == 112 and C.sizeof(PolyArgs) == 392

# the 16 entry points bslv_algs.o imports + the exported helpers
BOUNDARY_SYMBOLS = [
    "poly__set_default_args", "poly__initialise", "poly__add_vrtx", "poly__intl_apprx",
    "poly__get_vrtx", "poly__update_adjacence", "poly__swap", "poly__plot",
    "poly__initialise_permutation", "poly__kill_permutation", "poly__vrtx2file",
    "poly__primg2file", "poly__adj2file", "poly__inc2file", "poly__kill", "poly__polyck",
]


def _bind(lib):
    P = C.POINTER
    lib.poly__set_default_args.argtypes = [P(PolyArgs), C.c_size_t]
    lib.poly__set_default_args.restype = None
    for name in ("poly__initialise", "poly__kill"):
        getattr(lib, name).argtypes = [P(PolyArgs)]
        getattr(lib, name).restype = None
    for name in ("poly__add_vrtx", "poly__intl_apprx", "poly__get_vrtx"):
        getattr(lib, name).argtypes = [P(PolyArgs)]
        getattr(lib, name).restype = C.c_int
    lib.poly__update_adjacence.argtypes = [P(Polytope)]
    lib.poly__update_adjacence.restype = None
    if hasattr(lib, "poly__polyck"):
        lib.poly__polyck.argtypes = [P(PolyArgs)]
        lib.poly__polyck.restype = None
    if hasattr(lib, "poly__swap"):
        lib.poly__swap.argtypes = [P(PolyArgs), P(PolyArgs)]
        lib.poly__swap.restype = None
    if hasattr(lib, "poly__initialise_permutation"):
        lib.poly__initialise_permutation.argtypes = [P(Polytope), P(Permutation)]
        lib.poly__initialise_permutation.restype = None
        lib.poly__kill_permutation.argtypes = [P(Permutation)]
        lib.poly__kill_permutation.restype = None
        for name in ("poly__vrtx2file", "poly__primg2file", "poly__adj2file"):
            getattr(lib, name).argtypes = [P(Polytope), P(Permutation), C.c_char_p, C.c_char_p]
            getattr(lib, name).restype = None
        lib.poly__inc2file.argtypes = [P(Polytope), P(Permutation), P(Permutation), C.c_char_p, C.c_char_p]
        lib.poly__inc2file.restype = None
    if hasattr(lib, "poly__plot"):
        lib.poly__plot.argtypes = [P(Polytope), C.c_char_p]
        lib.poly__plot.restype = None
    # product-only extension entry points (include/bensolve_b200.h)
    if hasattr(lib, "b200_poly_materialise"):
        lib.b200_poly_materialise.argtypes = [P(PolyArgs)]
        lib.b200_poly_materialise.restype = C.c_int
        lib.b200_poly_set_flags.argtypes = [P(PolyArgs), C.c_uint]
        lib.b200_poly_set_flags.restype = C.c_int
        lib.b200_poly_add_batch.argtypes = [P(PolyArgs), P(C.c_double), P(C.c_ubyte), C.c_size_t, P(C.c_int)]
        lib.b200_poly_add_batch.restype = C.c_long
        lib.b200_poly_add_batch_device.argtypes = [P(PolyArgs), C.c_void_p, C.c_void_p, C.c_size_t, P(C.c_int)]
        lib.b200_poly_add_batch_device.restype = C.c_long
        lib.b200_poly_reserve.argtypes = [P(PolyArgs), C.c_size_t, C.c_size_t, C.c_size_t]
        lib.b200_poly_reserve.restype = C.c_int
        lib.b200_poly_classify_bench.argtypes = [P(PolyArgs), P(C.c_double), C.c_int, C.c_int]
        lib.b200_poly_classify_bench.restype = C.c_double
        lib.b200_poly_get_stats.argtypes = [P(PolyArgs), P(Stats)]
        lib.b200_poly_get_stats.restype = C.c_int
        lib.b200_last_error.restype = C.c_char_p
    return lib


def load_lib(path: str):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return _bind(C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW))


def load_product():
    """Load the CUDA engine.  No fallback: a missing extension is an error."""
    if not os.path.exists(PRODUCT_SO):
        raise RuntimeError(
            f"{PRODUCT_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " -- there is no CPU fallback for the cut step")
    return load_lib(PRODUCT_SO)


def is_elem(words, idx: int) -> int:  # IS_ELEM, bslv_poly.h:45
    return (words[idx // BTCNT] >> (idx % BTCNT)) & 1


@dataclass
class PolyState:
    """Canonical (slot-number independent) view of one engine's primal/dual state."""
    dim: int
    coords: np.ndarray            # [n_live, dim] float64, canonical order
    ideal: np.ndarray             # [n_live] uint8
    sltn: np.ndarray              # [n_live] uint8
    incidence: list               # per live vertex: sorted tuple of dual slot ids
    adjacency: list               # per live vertex: sorted tuple of canonical vertex ids
    live_facets: tuple            # dual slots that are used AND hold >=1 live vertex (ghosts dropped)
    facet_vertices: dict          # dual slot -> sorted tuple of canonical vertex ids
    slots: np.ndarray             # [n_live] original slot numbers (diagnostics only)
    n_slots: int = 0              # primal.cnt
    n_dual_slots: int = 0         # dual.cnt
    dual_used: tuple = field(default_factory=tuple)   # raw used bits of dual slots (incl. ghosts)

    @property
    def n_points(self):
        return int((self.ideal == 0).sum())

    @property
    def n_dirs(self):
        return int((self.ideal == 1).sum())


class PolyEngine:
    """Drives one ``poly_args`` through the reference API (same calls as bslv_algs.c:331-350)."""

    def __init__(self, lib, dim: int, callback=None, dim_primg_prml: int = 0, dim_primg_dl: int = 0, flags: int = 0):
        self.lib = lib
        self.dim = dim
        self.args = PolyArgs()
        self._args_ref = C.byref(self.args)
        lib.poly__set_default_args(C.byref(self.args), dim)
        self._cb = None
        if callback is not None:
            self._cb = V2H(callback)
            self.args.dualV2primalH = C.cast(self._cb, C.c_void_p)
        self.args.dim_primg_prml = dim_primg_prml
        self.args.dim_primg_dl = dim_primg_dl
        lib.poly__initialise(C.byref(self.args))
        if flags:
            lib.b200_poly_set_flags(C.byref(self.args), flags)   # product / emulation double only
        self.alive = True

    # -- the calls bslv_algs.c makes -------------------------------------------------------
    def add(self, val, ideal: int = 0) -> int:
        """poly__add_vrtx: append a dual point (= halfspace via the callback) and cut."""
        if isinstance(val, np.ndarray) and val.dtype == np.float64 and val.flags.c_contiguous:
            C.memmove(self.args.val, val.ctypes.data, 8 * self.dim)     # what a C caller's memcpy into args->val does
        else:
            for k in range(self.dim):
                self.args.val[k] = float(val[k])
        self.args.ideal = int(ideal)
        return self.lib.poly__add_vrtx(self._args_ref)

    def add_each(self, vals, ideal=None):
        """b200_poly_add_each: the C caller's loop, one poly__add_vrtx per row of `vals` (host memory)."""
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        n = len(vals)
        rcs = (C.c_int * n)()
        idl = None if ideal is None else np.ascontiguousarray(ideal, dtype=np.uint8)
        self.lib.b200_poly_add_each.restype = C.c_long
        self.lib.b200_poly_add_each.argtypes = [C.POINTER(PolyArgs), C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        self.lib.b200_poly_add_each(self._args_ref, vals.ctypes.data, None if idl is None else idl.ctypes.data, n, rcs)
        return list(rcs)

    def add_batch(self, vals, ideal=None):
        """b200_poly_add_batch: host arrays in, one device-resident pass, mirror coherent at return."""
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        n = len(vals)
        rcs = (C.c_int * n)()
        idl = None
        if ideal is not None:
            ideal = np.ascontiguousarray(ideal, dtype=np.uint8)
            idl = ideal.ctypes.data_as(C.POINTER(C.c_ubyte))
        got = self.lib.b200_poly_add_batch(C.byref(self.args), vals.ctypes.data_as(C.POINTER(C.c_double)), idl, n, rcs)
        if got < 0:
            raise RuntimeError(self.lib.b200_last_error().decode())
        return list(rcs)

    def add_batch_device(self, d_vals_ptr: int, d_ideal_ptr: int, n: int):
        rcs = (C.c_int * n)()
        got = self.lib.b200_poly_add_batch_device(C.byref(self.args), d_vals_ptr, d_ideal_ptr or None, n, rcs)
        if got < 0:
            raise RuntimeError(self.lib.b200_last_error().decode())
        return list(rcs)

    def reserve(self, vertices: int, inc_entries: int, adj_entries: int):
        return self.lib.b200_poly_reserve(C.byref(self.args), vertices, inc_entries, adj_entries)

    def stats(self) -> dict:
        st = Stats()
        self.lib.b200_poly_get_stats(C.byref(self.args), C.byref(st))
        return {n: getattr(st, n) for n, _ in Stats._fields_}

    def classify_bench(self, hp, iters: int, flush_l2: bool) -> float:
        arr = (C.c_double * (self.dim + 1))(*[float(x) for x in hp])
        return self.lib.b200_poly_classify_bench(C.byref(self.args), arr, iters, 1 if flush_l2 else 0)

    def init_approx(self) -> int:
        return self.lib.poly__intl_apprx(C.byref(self.args))

    def get_vrtx(self):
        rc = self.lib.poly__get_vrtx(C.byref(self.args))
        if rc:
            return rc, None, None, None
        return rc, int(self.args.idx), int(self.args.ideal), [self.args.val[k] for k in range(self.dim)]

    def mark_solution(self, idx: int):  # ST_BT(primal.sltn, idx), bslv_algs.c:1076
        self.args.primal.sltn[idx // BTCNT] |= 1 << (idx % BTCNT)

    def update_dual_adjacence(self):
        self.lib.poly__update_adjacence(C.byref(self.args.dual))

    def polyck(self):
        self.lib.poly__polyck(C.byref(self.args))

    def kill(self):
        if self.alive:
            self.lib.poly__kill(C.byref(self.args))
            self.alive = False

    def __del__(self):
        try:
            self.kill()
        except Exception:
            pass

    # -- state extraction --------------------------------------------------------------------
    def materialise(self):
        """Make host incidence/adjacence lists current (no-op for the CPU engines)."""
        if hasattr(self.lib, "b200_poly_materialise"):
            rc = self.lib.b200_poly_materialise(C.byref(self.args))
            if rc:
                raise RuntimeError(f"b200_poly_materialise failed rc={rc}")

    @staticmethod
    def _bits(words, cnt):
        nw = (cnt + BTCNT - 1) // BTCNT
        if nw == 0:
            return np.zeros(0, np.uint8)
        w = np.ctypeslib.as_array(words, shape=(nw,)).astype(np.uint64)
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")
        return bits[:cnt].astype(np.uint8)

    @staticmethod
    def _lists(lst_ptr, idxs):
        out = {}
        for i in idxs:
            l = lst_ptr[int(i)]
            n = int(l.cnt)
            out[int(i)] = [int(l.data[j]) for j in range(n)]
        return out

    def raw(self):
        """Raw slot-indexed view: dict with used/ideal/sltn bits, data and lists of live slots."""
        self.materialise()
        a = self.args
        d = self.dim
        out = {}
        for name, poly in (("primal", a.primal), ("dual", a.dual)):
            cnt = int(poly.cnt)
            used = self._bits(poly.used, cnt)
            ideal = self._bits(poly.ideal, cnt)
            sltn = self._bits(poly.sltn, cnt)
            data = (np.ctypeslib.as_array(poly.data, shape=(cnt * d,)).reshape(cnt, d).copy()
                    if cnt else np.zeros((0, d)))
            live = np.nonzero(used)[0]
            out[name] = dict(cnt=cnt, used=used, ideal=ideal, sltn=sltn, data=data,
                             inc=self._lists(poly.incidence, live), adj=self._lists(poly.adjacence, live))
        return out

    def state(self) -> PolyState:
        r = self.raw()
        p, du = r["primal"], r["dual"]
        live = np.nonzero(p["used"])[0]
        inc = {int(s): tuple(sorted(p["inc"][int(s)])) for s in live}
        # canonical order: incidence tuple, ideal flag, then coordinates (ties only on degenerate copies)
        key = lambda s: (inc[int(s)], int(p["ideal"][s]), tuple(np.round(p["data"][s], 9)))
        order = sorted(live, key=key)
        canon = {int(s): i for i, s in enumerate(order)}
        adjacency = []
        for s in order:
            nb = p["adj"][int(s)]
            for x in nb:
                if x not in canon:
                    raise AssertionError(f"adjacency of live slot {s} points at dead slot {x}")
            adjacency.append(tuple(sorted(canon[x] for x in nb)))
        fv = {}
        for f in np.nonzero(du["used"])[0]:
            vs = du["inc"][int(f)]
            if len(vs):
                for x in vs:
                    if x not in canon:
                        raise AssertionError(f"facet {f} lists dead vertex slot {x}")
                fv[int(f)] = tuple(sorted(canon[x] for x in vs))
        return PolyState(
            dim=self.dim,
            coords=p["data"][order] if len(order) else np.zeros((0, self.dim)),
            ideal=p["ideal"][order].astype(np.uint8) if len(order) else np.zeros(0, np.uint8),
            sltn=p["sltn"][order].astype(np.uint8) if len(order) else np.zeros(0, np.uint8),
            incidence=[inc[int(s)] for s in order],
            adjacency=adjacency,
            live_facets=tuple(sorted(fv)),
            facet_vertices=fv,
            slots=np.asarray(order, dtype=np.int64),
            n_slots=p["cnt"], n_dual_slots=du["cnt"],
            dual_used=tuple(int(x) for x in du["used"]),
        )


def compare_states(a: PolyState, b: PolyState, rtol: float = 1e-9, exact_coords: bool = False) -> None:
    """Parity gate of north_star: identical structure after canonical sorting, coords within rtol.

    Raises AssertionError with a description of the first difference."""
    assert a.dim == b.dim
    assert len(a.incidence) == len(b.incidence), f"live vertex count {len(a.incidence)} != {len(b.incidence)}"
    assert a.n_points == b.n_points and a.n_dirs == b.n_dirs, "point/direction counts differ"
    assert a.n_dual_slots == b.n_dual_slots, f"dual slot count {a.n_dual_slots} != {b.n_dual_slots}"
    assert a.live_facets == b.live_facets, (
        f"live facet sets differ: only-a={sorted(set(a.live_facets) - set(b.live_facets))[:8]} "
        f"only-b={sorted(set(b.live_facets) - set(a.live_facets))[:8]}")
    for i, (x, y) in enumerate(zip(a.incidence, b.incidence)):
        assert x == y, f"incidence of canonical vertex {i} differs: {x} vs {y}"
    assert (a.ideal == b.ideal).all(), "ideal flags differ"
    assert (a.sltn == b.sltn).all(), "sltn flags differ"
    for i, (x, y) in enumerate(zip(a.adjacency, b.adjacency)):
        assert x == y, f"adjacency of canonical vertex {i} (inc={a.incidence[i]}) differs: {x} vs {y}"
    assert a.facet_vertices == b.facet_vertices, "facet->vertex lists differ"
    if len(a.coords):
        if exact_coords:
            same = a.coords.view(np.uint64) == b.coords.view(np.uint64)
            # -0.0 vs +0.0 compare equal numerically; accept that
            same |= (a.coords == b.coords)
            assert same.all(), f"coordinates not bit-identical at {np.argwhere(~same)[:4].tolist()}"
        else:
            scale = np.maximum(1.0, np.maximum(np.abs(a.coords), np.abs(b.coords)))
            err = np.abs(a.coords - b.coords) / scale
            assert err.max() <= rtol, f"coordinate mismatch {err.max():.3e} > {rtol}"
