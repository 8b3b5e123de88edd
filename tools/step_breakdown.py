"""Where does one bench step spend its time?  (engine creation + reserve, init, cuts, mirror, kill)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bensolve_b200 import capi, polytopes as P
lib = capi.load_product()
d, n = 6, int(sys.argv[1]) if len(sys.argv) > 1 else 5000
tr = P.tangent_polytope(d, n, 20261018)
dv = torch.from_numpy(np.ascontiguousarray(tr.vals[d:])).cuda()
torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    e = capi.PolyEngine(lib, d); t.append(time.perf_counter())
    if rep: e.reserve(5300000, 5300000 * 8, 5300000 * 8)
    t.append(time.perf_counter())
    for i in range(d): e.add(tr.vals[i], 0)
    e.init_approx(); t.append(time.perf_counter())
    if rep == 2:
        rcs = [e.add(tr.vals[i], 0) for i in range(d, n)]
    else:
        rcs = e.add_batch_device(dv.data_ptr(), 0, n - d)
    t.append(time.perf_counter())
    st = e.stats(); e.kill(); t.append(time.perf_counter())
    names = ["create", "reserve", "init", "cuts(batch)" if rep < 2 else "cuts(per-call)", "stats+kill"]
    print(rep, " ".join(f"{k}={1e3*(b-a):.1f}ms" for k, a, b in zip(names, t, t[1:])), "cuts", st["cuts"], "launches", st["kernel_launches"], flush=True)
