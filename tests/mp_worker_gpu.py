"""Worker of the multi-GPU parity test (one process per GPU, NCCL): every rank replays the same
traces through the product library with K1 sharded across ranks and compares with the oracle."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

os.environ.setdefault("B200_SHARD_MIN_ROWS", "0")        # shard K1 of every cut, however small the polytope
os.environ.setdefault("B200_SHARD_MIN_ROWS_WAVE", "0")   # and every look-ahead pass of the wave path
os.environ.setdefault("B200_K4_SHARD", "1")              # and the pair test of every wave (opt-in otherwise)

import ctypes as C  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bensolve_b200 import capi, dist as bdist, polytopes as P  # noqa: E402
from traces import medium_traces, small_traces  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    lib = capi.load_product()
    lib.b200_set_device.argtypes = [C.c_int]
    lib.b200_set_device(local)
    oracle = capi.load_lib(capi.REF_SO if os.path.exists(capi.REF_SO) else capi.ORACLE_SO)
    assert bdist.init_comm(lib) == dist.get_world_size()
    traces = small_traces()[::2] + medium_traces() + [P.tangent_polytope(4, 3000, 21)]
    for tr in traces:
        a, b = capi.PolyEngine(oracle, tr.dim), capi.PolyEngine(lib, tr.dim)
        ra, rb = P.replay(a, tr), P.replay(b, tr)
        assert ra == rb, tr.name
        capi.compare_states(a.state(), b.state(), exact_coords=True)
        a.kill(); b.kill()
    # batch entry point as well
    tr = P.tangent_polytope(5, 150, 7)
    a, b = capi.PolyEngine(oracle, 5), capi.PolyEngine(lib, 5)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, 0)
    assert ra == rb
    capi.compare_states(a.state(), b.state(), exact_coords=True)
    a.kill(); b.kill()
    # wave path with the look-ahead passes sharded and exchanged over peer-mapped memory (waves from the first halfspace on)
    for tr in [P.tangent_polytope(5, 400, 3), P.tangent_polytope(6, 300, 88), P.tangent_polytope(4, 3000, 21), P.lattice_polytope(4, 60, 3),
               P.mixed_polyhedron(4, 80, 5)]:
        a, b = capi.PolyEngine(oracle, tr.dim), capi.PolyEngine(lib, tr.dim, flags=32)
        ra, rb = P.replay(a, tr), P.replay_batched(b, tr, 0)
        assert ra == rb, tr.name
        capi.compare_states(a.state(), b.state(), exact_coords=True)
        st = b.stats()
        assert st["sharded_passes"] > 0 and st["sharded_pair_tests"] > 0, st
        a.kill(); b.kill()
    torch.cuda.synchronize()
    dist.barrier()
    bdist.finalize_comm(lib)
    print(f"rank {rank}: {len(traces) + 1} traces OK on {dist.get_world_size()} GPUs", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
