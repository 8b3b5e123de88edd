"""Generate tests/golden/*.json from the UNMODIFIED reference engine (oracle/_ref/libref_poly.so,
built from /root/reference/bslv_poly.c by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the input trace (dual points as hex floats, ideal flags, n_init) and the
canonicalised state the reference reaches: return codes, coordinates (hex floats, so
the comparison can be bit-exact), ideal flags, incidence tuples, adjacency tuples, live facets.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from bensolve_b200 import capi, polytopes as P  # noqa: E402


def fixtures():
    return [P.cube_with_cuts(3), P.cube_with_cuts(4), P.tangent_polytope(3, 40, 5), P.tangent_polytope(4, 40, 5),
            P.tangent_polytope(5, 30, 5), P.tangent_polytope(6, 24, 5), P.lattice_polytope(4, 30, 3),
            P.lattice_polytope(5, 24, 3), P.random_cone(4, 20, 3), P.mixed_polyhedron(4, 40, 3),
            P.random_offsets(4, 50, 3)]


def dump_state(s):
    return dict(
        coords=[[float(x).hex() for x in row] for row in s.coords],
        ideal=[int(x) for x in s.ideal],
        incidence=[list(t) for t in s.incidence],
        adjacency=[list(t) for t in s.adjacency],
        live_facets=list(s.live_facets),
        n_slots=s.n_slots, n_dual_slots=s.n_dual_slots)


def main():
    ref = capi.load_lib(capi.REF_SO)
    for tr in fixtures():
        e = capi.PolyEngine(ref, tr.dim)
        rcs = P.replay(e, tr)
        st = e.state()
        out = dict(name=tr.name, dim=tr.dim, n_init=tr.n_init,
                   vals=[[float(x).hex() for x in row] for row in tr.vals],
                   ideal=[int(x) for x in tr.ideal], rcs=rcs, state=dump_state(st),
                   source="oracle/_ref/libref_poly.so (unmodified bslv_poly.c, gcc -std=c99 -O3)")
        with open(os.path.join(HERE, tr.name + ".json"), "w") as f:
            json.dump(out, f, separators=(",", ":"))
        e.kill()
        print(tr.name, st.n_points, st.n_dirs, os.path.getsize(os.path.join(HERE, tr.name + ".json")))


if __name__ == "__main__":
    main()
