"""Worker of the world_size-2 CPU test: both ranks replay the same traces through the sharded
small-cut path of the host test double (K1 over this rank's rows, exchange through a gloo
all-gather callback) and compare with the oracle."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

os.environ.setdefault("B200_SHARD_MIN_ROWS", "0")        # shard K1 of every cut, however small the polytope
os.environ.setdefault("B200_SHARD_MIN_ROWS_WAVE", "0")   # and every look-ahead pass of the wave path
os.environ.setdefault("B200_K4_SHARD", "1")              # and the pair test of every wave (opt-in otherwise)

import torch.distributed as dist  # noqa: E402

from bensolve_b200 import build, capi, dist as bdist, polytopes as P  # noqa: E402
from traces import small_traces  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    emul = capi.load_lib(build.build_emulation())
    oracle = capi.load_lib(capi.ORACLE_SO)
    assert bdist.init_comm(emul, emulate=True) == dist.get_world_size()
    traces = small_traces()[::2] + [P.tangent_polytope(4, 700, 3), P.tangent_polytope(3, 3000, 3)]
    for tr in traces:
        a = capi.PolyEngine(oracle, tr.dim)
        b = capi.PolyEngine(emul, tr.dim, flags=8 | (2 if tr.dim == 4 else 0))    # tail phases (+ eager compaction)
        ra, rb = P.replay(a, tr), P.replay(b, tr)
        assert ra == rb, tr.name
        capi.compare_states(a.state(), b.state(), exact_coords=True)
        assert b.stats()["sharded_cuts"] > 0
        a.kill(); b.kill()
    # wave path (device-resident batches): look-ahead passes split by row group, records exchanged through the callback
    n_wave = 0
    for tr, chunk in [(P.tangent_polytope(5, 120, 7), 0), (P.tangent_polytope(4, 500, 3), 0), (P.lattice_polytope(4, 60, 3), 7),
                      (P.mixed_polyhedron(4, 80, 5), 0), (P.cube_zero_plus(4), 0)]:
        a = capi.PolyEngine(oracle, tr.dim)
        b = capi.PolyEngine(emul, tr.dim, flags=32)           # waves from the first halfspace on
        ra, rb = P.replay(a, tr), P.replay_batched(b, tr, chunk)
        assert ra == rb, tr.name
        capi.compare_states(a.state(), b.state(), exact_coords=True)
        st = b.stats()
        a.kill(); b.kill()
        assert st["sharded_passes"] > 0 and st["sharded_passes"] == st["lookahead_passes"] and st["sharded_pair_tests"] > 0, st
        n_wave += 1
    # cone_vertenum's pattern on several ranks: everything queued, poly__intl_apprx re-adds it as one device-resident batch
    for tr in (P.tangent_polytope(4, 200, 9), P.random_cone(5, 40, 2)):
        q = P.Trace(tr.dim, tr.vals, tr.ideal, len(tr.vals), tr.name + "_queued")
        a, b = capi.PolyEngine(oracle, q.dim), capi.PolyEngine(emul, q.dim)
        assert P.replay(a, q) == P.replay(b, q)
        capi.compare_states(a.state(), b.state(), exact_coords=True)
        a.kill(); b.kill()
    bdist.finalize_comm(emul)
    dist.barrier()
    print(f"rank {rank}: {len(traces)} traces OK, {n_wave} wave traces OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
