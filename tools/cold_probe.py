"""GPU probe: the unchanged caller's path (no reserve), per-call, first polytope of the process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bensolve_b200 import capi, polytopes as P

lib = capi.load_product()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6400
mode = sys.argv[2] if len(sys.argv) > 2 else "each"
tiny = P.tangent_polytope(3, 40, 1)
e0 = capi.PolyEngine(lib, 3); P.replay(e0, tiny); e0.kill()
tr = P.tangent_polytope(6, n, 20261018)
for rep in range(2):
    e = capi.PolyEngine(lib, 6)
    for i in range(6):
        e.add(tr.vals[i], 0)
    assert e.init_approx() == 0
    t0 = time.perf_counter()
    if mode == "each":
        e.add_each(tr.vals[6:])
    else:
        e.add_batch(tr.vals[6:])
    dt = time.perf_counter() - t0
    st = e.stats()
    print(f"rep {rep} {mode}: {dt*1e3:.1f} ms, {st['cuts']/dt:.0f} cuts/s", flush=True)
    e.kill()
