"""Remaining boundary functions against the reference: poly__swap, poly__plot, failure codes of
poly__intl_apprx / poly__get_vrtx."""
import ctypes as C

import numpy as np
import pytest

from bensolve_b200 import capi, polytopes as P


def _swap_state(lib, tr):
    """poly__swap feeds the vertices of `a` as halfspaces into `b` in slot / list order, which is
    order-dependent, so facets of `b` are identified by their dual points, not by id."""
    a = capi.PolyEngine(lib, tr.dim)
    P.replay(a, tr)
    b = capi.PolyEngine(lib, tr.dim)
    lib.poly__swap(C.byref(a.args), C.byref(b.args))
    r = b.raw()
    key = lambda x: tuple(np.round(x, 9) + 0.0)
    dual_pt = {f: (int(r["dual"]["ideal"][f]), key(r["dual"]["data"][f])) for f in range(r["dual"]["cnt"])}
    verts = {}
    for s in np.nonzero(r["primal"]["used"])[0]:
        verts[(int(r["primal"]["ideal"][s]), key(r["primal"]["data"][s]))] = (
            frozenset(dual_pt[f] for f in r["primal"]["inc"][int(s)]),
            frozenset((int(r["primal"]["ideal"][x]), key(r["primal"]["data"][x])) for x in r["primal"]["adj"][int(s)]))
    a.kill(); b.kill()
    return verts


def _plot_facets(lib, tr, path):
    e = capi.PolyEngine(lib, tr.dim)
    P.replay(e, tr)
    lib.poly__plot(C.byref(e.args.primal), str(path).encode())
    pts, facets, mode = [], [], None
    for line in open(path):
        line = line.strip()
        if line.startswith("#vertices"):
            mode = "v"
        elif line.startswith("#facets"):
            mode = "f"
        elif line and mode == "v":
            pts.append(tuple(round(float(x), 6) for x in line.split()))
        elif line and mode == "f":
            idx = [int(x) for x in line.split()]
            assert idx[0] == len(idx) - 1
            facets.append(frozenset(pts[i] for i in idx[1:]))
    e.kill()
    return set(pts), set(facets)


def _misc(lib_a, lib_b, tmp_path):
    for tr in (P.cube(3), P.tangent_polytope(3, 14, 3)):
        assert _swap_state(lib_a, tr) == _swap_state(lib_b, tr)
    for i, tr in enumerate((P.cube(3), P.tangent_polytope(3, 20, 5), P.cube_with_cuts(3))):
        assert _plot_facets(lib_a, tr, tmp_path / f"a{i}.off") == _plot_facets(lib_b, tr, tmp_path / f"b{i}.off")


def test_swap_and_plot_host_logic(ref_lib, emul_lib, tmp_path):
    _misc(ref_lib, emul_lib, tmp_path)


@pytest.mark.gpu
def test_swap_and_plot_gpu(ref_lib, product_lib, tmp_path):
    _misc(ref_lib, product_lib, tmp_path)


def _failure_codes(lib):
    # fewer than d halfspaces queued: EXIT_FAILURE (bslv_poly.c:158-159)
    e = capi.PolyEngine(lib, 3)
    e.add([1.0, 0.0, 0.0]); e.add([0.0, 1.0, 0.0])
    assert e.init_approx() == 1
    e.kill()
    # rank-deficient queue: EXIT_FAILURE (bslv_poly.c:174-177)
    e = capi.PolyEngine(lib, 3)
    for v in ([1.0, 0.0, 0.0], [2.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0]):
        e.add(v)
    assert e.init_approx() == 1
    e.kill()
    # nothing left to hand out: EXIT_FAILURE (bslv_poly.c:217-218)
    e = capi.PolyEngine(lib, 2)
    P.replay(e, P.tangent_polytope(2, 6, 1))
    n = 0
    while True:
        rc, idx, _, _ = e.get_vrtx()
        if rc:
            break
        e.mark_solution(idx)
        n += 1
    assert n == 6 and e.get_vrtx()[0] == 1
    e.kill()


def test_failure_codes_reference(ref_lib):
    _failure_codes(ref_lib)


def test_failure_codes_host_logic(emul_lib):
    _failure_codes(emul_lib)


@pytest.mark.gpu
def test_failure_codes_gpu(product_lib):
    _failure_codes(product_lib)
