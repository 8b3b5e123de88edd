"""Build one polytope, then time K1 alone (flushed and L2-resident) -- run once per B200_K1_IT value."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bensolve_b200 import capi, polytopes as P
lib = capi.load_product()
d, n = int(sys.argv[1]), int(sys.argv[2])
tr = P.tangent_polytope(d, n, 20261018)
dv = torch.from_numpy(np.ascontiguousarray(tr.vals[d:])).cuda()
e = capi.PolyEngine(lib, d)
for i in range(d): e.add(tr.vals[i], 0)
e.init_approx()
e.add_batch_device(dv.data_ptr(), 0, n - d)
st = e.stats()
hp = np.append(tr.vals[n // 2] * 1.0000001, -1.0)
for _ in range(2):
    msf = e.classify_bench(hp, 50, True); msl = e.classify_bench(hp, 50, False)
nb = st["live_vertices"] * (8 * d + 1)
print(json.dumps(dict(it=os.environ.get("B200_K1_IT"), d=d, live=st["live_vertices"], mb=nb / 1e6, ms_flush=msf, gbs_flush=nb / msf / 1e6, ms_l2=msl, gbs_l2=nb / msl / 1e6)))
e.kill()
