"""Cut-path time per poly__add_vrtx call on recorded Benson traces (BASELINE configs 2-4): the val-level call sequence
of a real bensolve run (tests/golden/benson_*.json, recorded with the reference engine) replayed into the B200 engine
and into the unmodified reference engine, timing only the calls into the engine -- the LP time of the run that
produced the trace is reported beside it by tools/solve_time.py.

    python tools/trace_bench.py syn_q5_m40_n20 syn_q3_m120_n60 ex10        -> one JSON line per trace
"""
from __future__ import annotations

import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bensolve_b200 import capi  # noqa: E402


def replay(lib, inst):
    d = inst["dim"]
    table = {}
    for ev in inst["events"]:
        if ev[0] == "add":
            table[(tuple(float.fromhex(x) for x in ev[1]), int(ev[2]))] = [float.fromhex(x) for x in ev[3]]

    def callback(dual_point, is_dir, hp_out):
        hp = table[(tuple(dual_point[k] for k in range(d)), int(is_dir))]
        for k in range(d + 1):
            hp_out[k] = hp[k]

    e = capi.PolyEngine(lib, d, callback=callback)
    t_cut, n_cut, t_max = 0.0, 0, 0.0
    for ev in inst["events"]:
        if ev[0] == "add":
            val = [float.fromhex(x) for x in ev[1]]
            for k in range(d):
                e.args.val[k] = val[k]
            e.args.ideal = int(ev[2])
            t0 = time.perf_counter()
            rc = e.lib.poly__add_vrtx(e._args_ref)
            dt = time.perf_counter() - t0
            assert rc == ev[4]
            if e.args.init_data.intlsd:
                t_cut += dt
                n_cut += 1
                t_max = max(t_max, dt)
        else:
            if ev[1]:
                e.args.dual.ideal[0] |= 1
            else:
                e.args.dual.ideal[0] &= ~1
            for k in range(d):
                e.args.dual.data[k] = float.fromhex(ev[2][k])
            assert e.init_approx() == ev[3]
    pts = int(e.args.primal.cnt)
    e.kill()
    return {"calls": n_cut, "seconds": t_cut, "us_per_call": 1e6 * t_cut / max(1, n_cut), "max_ms": 1e3 * t_max, "primal_slots": pts}


def main():
    product = capi.load_product()
    ref = capi.load_lib(capi.REF_SO if os.path.exists(capi.REF_SO) else capi.ORACLE_SO)
    for name in sys.argv[1:]:
        fx = json.load(open(os.path.join(REPO, "tests", "golden", f"benson_{name}.json")))
        inst = max(fx["instances"], key=lambda i: len(i["events"]))          # the main (phase 2) polyhedron of the run
        replay(product, inst)                                                # warm-up (context, first allocations)
        out = {"trace": name, "dim": inst["dim"], "final": inst.get("final"), "b200": replay(product, inst), "reference": replay(ref, inst)}
        out["speedup"] = out["reference"]["seconds"] / out["b200"]["seconds"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
