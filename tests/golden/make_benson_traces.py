"""Record the poly__* call sequences of real Benson runs (build container only).

Runs the UNMODIFIED reference CLI (tools/run_bensolve.py --engine ref: reference host code +
reference engine, LPs served by the HiGHS stand-in for GLPK) on the reference's own example
problems with the trace recorder interposed, and stores one fixture per example under
tests/golden/benson_<ex>.json: for every poly_args instance the run created, its dimension and the
ordered events

    ["add", val[], ideal, hp[], rc]     one poly__add_vrtx call; hp = what the caller's callback
                                        derives from (val, ideal) -- replayed through a lookup callback
    ["apprx", slot0_ideal, slot0[], rc] poly__intl_apprx, with dual slot 0 as the caller left it
                                        (cone_vertenum patches it, bslv_algs.c:338-339)

plus the reference's final counts (points, directions, slots, dual slots).

    python tests/golden/make_benson_traces.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
EX = os.environ.get("BSLV_REF", "/root/reference") + "/ex"

RUNS = {
    "ex01": [], "ex05": [], "ex06": [], "ex08": [], "ex11": [], "ex10": [],
    "ex07": ["-e", "0.05", "-l", "primal_simplex"],
    "ex05_dual": ["-A", "dual", "-a", "dual"],
    "ex11_dual": ["-A", "dual", "-a", "dual"],
    # BASELINE config 2: the flags ex/example09.m:9-24 prescribes (61 s of LP time with the HiGHS stand-in)
    "ex09": ["-e", "1e-2", "-L", "primal_simplex", "-l", "primal_simplex"],
}
# BASELINE configs 3-4 (synthetic random VLPs, bensolve_b200/vlpgen.py), scaled to what the LP stand-in finishes in
# minutes: every image vertex costs one LP through scipy, so q=3, m=2000, n=1000 (hours of LP time) is out of reach
# here -- the cut path sees the same kind of trace, only shorter.  name -> (q, m, n, seed)
SYNTHETIC = {
    "syn_q3_m120_n60": (3, 120, 60, 1),
    "syn_q5_m40_n20": (5, 40, 20, 1),
}


def main():
    import time
    only = set(sys.argv[1:])
    todo = {**RUNS, **{k: [] for k in SYNTHETIC}}
    for name, flags in todo.items():
        if only and name not in only:
            continue
        ex = name.split("_")[0]
        with tempfile.TemporaryDirectory() as tmp:
            trace = os.path.join(tmp, "trace.jsonl")
            vlp = os.path.join(EX, ex + ".vlp")
            if name in SYNTHETIC:
                sys.path.insert(0, REPO)
                from bensolve_b200 import vlpgen
                q, m, n, seed = SYNTHETIC[name]
                os.makedirs(os.path.join(tmp, "in"), exist_ok=True)
                vlp = os.path.join(tmp, "in", name + ".vlp")
                vlpgen.write_vlp(vlp, *vlpgen.random_vlp(q, m, n, seed=seed))
            t0 = time.time()
            res = subprocess.run([sys.executable, os.path.join(REPO, "tools", "run_bensolve.py"), "--engine", "ref", "--record", trace,
                                  "--workdir", tmp, vlp] + flags, capture_output=True, text=True)
            print(name, "reference run: %.1f s" % (time.time() - t0), [l for l in res.stdout.splitlines() if "LPs" in l][-1:], flush=True)
            if not os.path.exists(trace):
                print(name, "FAILED", res.stdout[-500:], res.stderr[-500:])
                continue
            inst, order = {}, []
            for line in open(trace):
                e = json.loads(line)
                k = e["id"]
                if e["ev"] == "init":
                    k2 = f"{k}#{len(order)}"          # stack addresses are reused by later instances
                    inst[k] = dict(dim=e["dim"], events=[], key=k2)
                    order.append(inst[k])
                elif e["ev"] == "add":
                    inst[k]["events"].append(["add", e["val"], e["ideal"], e["hp"], e["rc"]])
                elif e["ev"] == "apprx":
                    inst[k]["events"].append(["apprx", e["slot0_ideal"], e["slot0"], e["rc"]])
                elif e["ev"] == "kill":
                    inst[k]["final"] = dict(points=e["points"], dirs=e["dirs"], slots=e["slots"], dual_slots=e["dual_slots"])
            for o in order:
                o.pop("key")
            out = dict(example=name, flags=flags, instances=order,
                       source="unmodified reference CLI + reference engine; LP backend: scipy HiGHS behind tools/lpshim/glpk.h")
            path = os.path.join(HERE, f"benson_{name}.json")
            json.dump(out, open(path, "w"), separators=(",", ":"))
            print(name, [(o["dim"], sum(1 for e in o["events"] if e[0] == "add"), o.get("final")) for o in order], os.path.getsize(path))


if __name__ == "__main__":
    main()
