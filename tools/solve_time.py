"""End-to-end solve time of the UNMODIFIED bensolve CLI (the "CPU time" line bslv_main.c:401 prints, interval
bslv_main.c:236 .. bslv_algs.c:1140/1572) with the reference polyhedron engine and with the B200 engine behind the
same poly__* symbols, same LP stand-in (tools/lpshim + scipy HiGHS).  Third component of BASELINE.json's metric.

    python tools/solve_time.py oracle/_ref/ex/ex10.vlp [bensolve flags]      -> one JSON line

The LP stand-in dominates every example: `lp_seconds` is the time spent inside it, `rest_seconds` what is
left for the caller and the engine."""
from __future__ import annotations

import json
import os
import re
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_time(text):
    m = re.search(r"CPU time\s*:\s*([0-9.eE+-]+)\s*(ms|s)", text)
    if not m:
        return None
    return float(m.group(1)) * (1e-3 if m.group(2) == "ms" else 1.0)


def run(engine, vlp, flags=()):
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        res = subprocess.run([sys.executable, os.path.join(REPO, "tools", "run_bensolve.py"), "--engine", engine, "--workdir", tmp, vlp, *flags],
                             capture_output=True, text=True, timeout=3600)
        wall = time.perf_counter() - t0
        m = re.search(r"LPs=(\d+) lp_seconds=([0-9.]+)", res.stdout)
        cpu = _cpu_time(res.stdout)
        if cpu is None or not m:
            raise RuntimeError((res.stdout + res.stderr)[-2000:])
        return {"engine": engine, "cpu_time_s": cpu, "lps": int(m.group(1)), "lp_seconds": float(m.group(2)),
                "rest_seconds": max(0.0, cpu - float(m.group(2))), "wall_s": wall}


def compare(vlp, flags=(), engines=("ref", "b200")):
    out = {"problem": os.path.basename(vlp), "flags": list(flags), "clock": "bensolve's own 'CPU time' line (bslv_main.c:401)",
           "lp_backend": "scipy HiGHS behind tools/lpshim/glpk.h (GLPK is not in this image)"}
    for e in engines:
        out[e] = run(e, vlp, flags)
    return out


if __name__ == "__main__":
    print(json.dumps(compare(sys.argv[1], sys.argv[2:])))
