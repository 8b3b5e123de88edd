"""Run the UNMODIFIED bensolve host code (bslv_main.c, bslv_vlp.c, bslv_algs.c, bslv_lists.c,
bslv_lp.c compiled from where they lie under /root/reference) in this process, with

  * the LP layer served by tools/lpshim (a <glpk.h> stand-in whose glp_simplex calls scipy's HiGHS), and
  * the polyhedron engine chosen at load time:  --engine ref   -> oracle/_ref/libref_poly.so
                                                --engine b200  -> bensolve_b200/libbslv_poly_b200.so
                                                --engine emul  -> tests/_emul/libbslv_poly_emul.so
  * optionally the trace recorder in front of the reference engine (--record FILE).

    python tools/run_bensolve.py --engine ref --workdir /tmp/run /root/reference/ex/ex01.vlp [bensolve flags]

Build container only for --engine ref/--record (needs /root/reference); the host library
oracle/_ref/libbensolve_host.so travels to the GPU box, so --engine b200 runs there too.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BSLV_REF", "/root/reference")
OUT = os.path.join(REPO, "oracle", "_ref")
HOST_SO = os.path.join(OUT, "libbensolve_host.so")
REC_SO = os.path.join(OUT, "librecorder.so")
SHIM = os.path.join(REPO, "tools", "lpshim")


def build_host(force=False):
    """oracle/_ref/libbensolve_host.so: the five GLPK-free TUs + bslv_lp.c against the stand-in header."""
    deps = [os.path.join(SHIM, f) for f in ("glpk.h", "glpk_shim.c", "recorder.c")]
    if os.path.exists(HOST_SO) and os.path.exists(REC_SO) and not force and all(os.path.getmtime(d) <= os.path.getmtime(REC_SO) for d in deps):
        return HOST_SO
    if not os.path.exists(os.path.join(REF, "bslv_algs.c")):
        raise RuntimeError("reference sources absent and no prebuilt " + HOST_SO)
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(REF, f) for f in ("bslv_main.c", "bslv_vlp.c", "bslv_algs.c", "bslv_lists.c", "bslv_lp.c")]
    cmd = ["gcc", "-std=c99", "-O3", "-w", "-fPIC", "-shared", "-Dmain=bensolve_main", "-I", SHIM, "-I", REF,
           "-o", HOST_SO] + srcs + [os.path.join(SHIM, "glpk_shim.c"), "-lm"]
    subprocess.check_call(cmd)
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-o", REC_SO, os.path.join(SHIM, "recorder.c"), "-ldl"])
    return HOST_SO


class Prob(C.Structure):   # mirrors struct glp_prob in tools/lpshim/glpk_shim.c
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("cap_m", C.c_int), ("cap_n", C.c_int),
                ("rtype", C.POINTER(C.c_int)), ("ctype", C.POINTER(C.c_int)),
                ("rlb", C.POINTER(C.c_double)), ("rub", C.POINTER(C.c_double)), ("clb", C.POINTER(C.c_double)),
                ("cub", C.POINTER(C.c_double)), ("obj", C.POINTER(C.c_double)),
                ("rlen", C.POINTER(C.c_int)), ("rind", C.POINTER(C.POINTER(C.c_int))), ("rval", C.POINTER(C.POINTER(C.c_double))),
                ("status", C.c_int), ("pstat", C.c_int), ("dstat", C.c_int), ("objval", C.c_double),
                ("rprim", C.POINTER(C.c_double)), ("rdual", C.POINTER(C.c_double)), ("cprim", C.POINTER(C.c_double)), ("cdual", C.POINTER(C.c_double))]


FR, LO, UP, DB, FX = 1, 2, 3, 4, 5
UNDEF, FEAS, INFEAS, NOFEAS, OPT, UNBND = 1, 2, 3, 4, 5, 6
N_LP = [0]
T_LP = [0.0]     # seconds inside the LP stand-in (the share of "CPU time" that is not the caller's or the engine's)


def solve(pp, meth):
    import numpy as np
    from scipy.optimize import linprog
    from scipy.sparse import csr_matrix
    import time as _time
    _t0 = _time.perf_counter()
    try:
        return _solve(pp, meth)
    finally:
        T_LP[0] += _time.perf_counter() - _t0


def _solve(pp, meth):
    import numpy as np
    from scipy.optimize import linprog
    from scipy.sparse import csr_matrix
    P = pp.contents
    m, n = P.m, P.n
    N_LP[0] += 1
    as_np = lambda ptr, k, dt: np.ctypeslib.as_array(ptr, shape=(k,)).astype(dt, copy=True)
    rtype, ctype = as_np(P.rtype, m + 1, np.int64), as_np(P.ctype, n + 1, np.int64)
    rlb, rub = as_np(P.rlb, m + 1, np.float64), as_np(P.rub, m + 1, np.float64)
    clb, cub = as_np(P.clb, n + 1, np.float64), as_np(P.cub, n + 1, np.float64)
    obj = as_np(P.obj, n + 1, np.float64)
    rlen = as_np(P.rlen, m + 1, np.int64)
    indptr = np.zeros(m + 1, np.int64)
    indptr[1:] = np.cumsum(rlen[1:])
    idx = np.empty(indptr[-1], np.int64)
    val = np.empty(indptr[-1], np.float64)
    for i in range(1, m + 1):
        k = int(rlen[i])
        if k:
            idx[indptr[i - 1]:indptr[i]] = np.ctypeslib.as_array(P.rind[i], shape=(k,)) - 1
            val[indptr[i - 1]:indptr[i]] = np.ctypeslib.as_array(P.rval[i], shape=(k,))
    A = csr_matrix((val, idx, indptr), shape=(m, n))
    rt = rtype[1:]
    lo_rows = np.nonzero((rt == LO) | (rt == DB))[0]
    up_rows = np.nonzero((rt == UP) | (rt == DB))[0]
    eq_rows = np.nonzero(rt == FX)[0]
    from scipy.sparse import vstack
    parts, rhs = [], []
    if len(lo_rows):
        parts.append(-A[lo_rows]); rhs.append(-rlb[1:][lo_rows])
    if len(up_rows):
        parts.append(A[up_rows]); rhs.append(rub[1:][up_rows])
    A_ub = vstack(parts).tocsr() if parts else None
    b_ub = np.concatenate(rhs) if rhs else None
    A_eq = A[eq_rows] if len(eq_rows) else None
    b_eq = rlb[1:][eq_rows] if len(eq_rows) else None
    bounds = []
    for j in range(1, n + 1):
        t = ctype[j]
        bounds.append((None, None) if t == FR else (clb[j], None) if t == LO else (None, cub[j]) if t == UP
                      else (clb[j], cub[j]) if t == DB else (clb[j], clb[j]))
    method = "highs-ds"
    res = linprog(obj[1:], A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method=method)
    if res.status == 0:
        x = res.x
        P.status, P.pstat, P.dstat = OPT, FEAS, FEAS
        P.objval = float(obj[1:] @ x + obj[0])
        ax = A @ x
        rd = np.zeros(m)
        if len(lo_rows):
            rd[lo_rows] += -res.ineqlin.marginals[:len(lo_rows)]
        if len(up_rows):
            rd[up_rows] += res.ineqlin.marginals[len(lo_rows):]
        if len(eq_rows):
            rd[eq_rows] = res.eqlin.marginals
        cd = res.lower.marginals + res.upper.marginals
        for i in range(m):
            P.rprim[i + 1] = ax[i]; P.rdual[i + 1] = rd[i]
        for j in range(n):
            P.cprim[j + 1] = x[j]; P.cdual[j + 1] = cd[j]
    elif res.status == 2:
        P.status, P.pstat, P.dstat = NOFEAS, NOFEAS, UNDEF
    elif res.status == 3:
        P.status, P.pstat, P.dstat = UNBND, FEAS, NOFEAS
    else:
        P.status, P.pstat, P.dstat = UNDEF, UNDEF, UNDEF
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="ref", choices=["ref", "b200", "emul"])
    ap.add_argument("--record", default=None, help="write a val-level trace of every poly__* call (reference engine only)")
    ap.add_argument("--workdir", default=None, help="copy the .vlp here first (bensolve writes its outputs next to the input)")
    ap.add_argument("vlp")
    ap.add_argument("flags", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    build_host()
    vlp = a.vlp
    if a.workdir:
        os.makedirs(a.workdir, exist_ok=True)
        vlp = os.path.join(a.workdir, os.path.basename(a.vlp))
        shutil.copy(a.vlp, vlp)
    mode = os.RTLD_GLOBAL | os.RTLD_NOW
    if a.record:
        assert a.engine == "ref"
        os.environ["BSLV_TRACE"] = a.record
        os.environ["BSLV_ENGINE_SO"] = os.path.join(OUT, "libref_poly.so")
        C.CDLL(REC_SO, mode=mode)
    engine = {"ref": os.path.join(OUT, "libref_poly.so"), "b200": os.path.join(REPO, "bensolve_b200", "libbslv_poly_b200.so"),
              "emul": os.path.join(REPO, "tests", "_emul", "libbslv_poly_emul.so")}[a.engine]
    C.CDLL(engine, mode=mode)
    host = C.CDLL(HOST_SO, mode=mode)
    cb_t = C.CFUNCTYPE(C.c_int, C.POINTER(Prob), C.c_int)
    cb = cb_t(solve)
    host.glp_shim_set_solver.argtypes = [cb_t]
    host.glp_shim_set_solver(cb)
    argv = [b"bensolve", vlp.encode()] + [f.encode() for f in a.flags]
    arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
    host.bensolve_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    rc = host.bensolve_main(len(argv), arr)
    sys.stdout.flush()
    print(f"[run_bensolve] engine={a.engine} rc={rc} LPs={N_LP[0]} lp_seconds={T_LP[0]:.3f}", flush=True)
    os._exit(rc)       # the host frees GLPK state at exit; skip interpreter teardown ordering issues


if __name__ == "__main__":
    main()
