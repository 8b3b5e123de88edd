"""Multi-GPU probe of the device-resident batch path (run under torchrun): same trace on every rank, state replicated,
work split as the environment says; prints the step time as the max over ranks."""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bensolve_b200 import capi, dist as bdist, polytopes as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--n", type=int, default=6400)
    ap.add_argument("--reserve", type=int, default=7000000)
    ap.add_argument("--repeat", type=int, default=5)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load_product()
    lib.b200_set_device.argtypes = [C.c_int]
    lib.b200_set_device(local)
    bdist.init_comm(lib)
    tr = P.tangent_polytope(a.dim, a.n, 20261018)
    dev = torch.device("cuda", local)
    d_vals = torch.from_numpy(np.ascontiguousarray(tr.vals[a.dim:])).to(dev)
    out = []
    for rep in range(a.repeat):
        e = capi.PolyEngine(lib, a.dim)
        e.reserve(a.reserve, a.reserve * (a.dim + 2), a.reserve * (a.dim + 2))
        for i in range(a.dim):
            e.add(tr.vals[i], 0)
        assert e.init_approx() == 0
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        e.add_batch_device(d_vals.data_ptr(), 0, a.n - a.dim)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = e.stats()
        e.kill()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(dict(ms=1e3 * float(t[0]), cuts_per_s=st["cuts"] / float(t[0]), sharded_passes=st["sharded_passes"],
                        lookahead_passes=st["lookahead_passes"], sharded_pair_tests=st["sharded_pair_tests"], waves=st["waves"]))
    if dist.get_rank() == 0:
        print(json.dumps(dict(world=dist.get_world_size(), reps=out)), flush=True)
    dist.barrier()
    bdist.finalize_comm(lib)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
