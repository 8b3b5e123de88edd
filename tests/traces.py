"""The trace suite shared by the oracle, host-logic and GPU parity tests."""
from bensolve_b200 import polytopes as P


def small_traces():
    ts = [P.cube(3), P.cube(4), P.cube_zero_plus(3), P.cube_zero_plus(4), P.cube_zero_plus(5), P.pyramid(12), P.pyramid(70), P.pyramid(300), P.cube_with_cuts(3), P.cube_with_cuts(4), P.cube_with_cuts(5),
          P.tangent_polytope(2, 30), P.tangent_polytope(3, 50), P.tangent_polytope(4, 80),
          P.tangent_polytope(5, 60), P.tangent_polytope(6, 40)]
    for s in range(1, 4):
        ts += [P.lattice_polytope(3, 40, s), P.lattice_polytope(4, 30, s), P.lattice_polytope(5, 24, s),
               P.random_cone(4, 20, s), P.random_cone(5, 15, s), P.mixed_polyhedron(3, 30, s),
               P.mixed_polyhedron(4, 40, s), P.random_offsets(4, 60, s)]
    return ts


def stepwise_traces():
    """Compared after EVERY cut (trace replay)."""
    return [P.cube_with_cuts(4), P.cube_zero_plus(4), P.tangent_polytope(3, 25), P.tangent_polytope(4, 30), P.lattice_polytope(4, 30, 2),
            P.lattice_polytope(5, 20, 1), P.mixed_polyhedron(3, 30, 1), P.random_cone(4, 16, 2)]


def medium_traces():
    return [P.tangent_polytope(3, 2000, 7), P.tangent_polytope(4, 400, 7), P.tangent_polytope(5, 150, 7),
            P.tangent_polytope(6, 70, 7), P.lattice_polytope(5, 60, 11, kmax=3, bmax=4),
            P.lattice_polytope(6, 40, 12), P.random_offsets(5, 200, 3), P.mixed_polyhedron(5, 120, 9)]
