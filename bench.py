#!/usr/bin/env python
"""bench.py -- cut throughput of the B200 polyhedral cut engine (BASELINE.json metric).

Workload (config 5 of BASELINE.json, the largest single-GPU configuration): pure H->V enumeration
of a random polytope in R^6 from N halfspaces tangent to the unit ball (synthetic, fixed seed).
One *step* = the whole cut sequence: one cut per halfspace after the start simplex (poly__initialise,
d queued halfspaces, poly__intl_apprx and the storage reservation are set-up, untimed), ending with a
coherent host mirror.

    value : cuts/s with the halfspaces already resident in HBM (b200_poly_add_batch_device: look-ahead
            classification + waves of commuting cuts)
    e2e   : cuts/s through the reference-facing call, one poly__add_vrtx per halfspace with HOST
            buffers; every call returns with primal.data/used/ideal/cnt current on the host.  The loop
            around poly__add_vrtx is the C caller's (b200_poly_add_each = what bslv_algs.c writes),
            not a Python loop.  e2e_unchanged_caller: the same without b200_poly_reserve and without a
            recycled host block (first polytope of the process); e2e_batch: all halfspaces handed over
            in one call from host memory (b200_poly_add_batch, an extension); e2e_vertenum: the unchanged
            API used the way bensolve enumerates vertices (queue everything, then poly__intl_apprx)
    roofline : frac = SURVEY 8(d)'s sequence figure, sum over cuts of the algorithmic bytes B_c divided
               by the step time, against MEASURED_PEAKS.json hbm_gbs; roofline.k1 = the classify kernel
               timed alone with CUDA events on its own stream with an L2 flush before each launch
    parity : the gate.  Full sequence: properties of a simple polytope on both paths' final state and
             equality of the two; prefix: both paths bit-exact against the unmodified reference engine.
             A mismatch exits non-zero
    same_prefix : CPU reference, value and e2e on the same first --ref-prefix halfspaces
    shapes : vertex-evals/s and HBM fractions for the d=3 and d=5 shapes of the metric
    cpu_baseline : the unmodified reference engine (oracle/_ref) on a bounded prefix of the trace

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libref_poly.so, built
from the unmodified bslv_poly.c; falls back to the restatement oracle/libpoly_oracle.so) on the
same bounded prefix.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

from bensolve_b200 import capi, polytopes as P  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--halfspaces", type=int, default=6400,
                    help="BASELINE config 5 names 5000 halfspaces and >=10^6 vertices; 5000 tangent halfspaces in R^6 give 7.8e5 vertices, 6400 give 1.02e6")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--ref-prefix", type=int, default=300, help="halfspaces of the trace the CPU reference is timed on")
    ap.add_argument("--classify-iters", type=int, default=30)
    ap.add_argument("--solve-time", default="ex10", help="reference example (oracle/_ref/ex/NAME.vlp) whose end-to-end solve time is measured with both engines (1 GPU only; empty = skip)")
    ap.add_argument("--shapes", default="3:200000,5:12000", help="dim:halfspaces of the extra vertex-eval shapes (1 GPU only; empty = skip)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU reference arm
def load_cpu_engine():
    from bensolve_b200 import build
    if not (os.path.exists(capi.REF_SO) or os.path.exists(capi.ORACLE_SO)):
        build.build_oracle()
    if os.path.exists(capi.REF_SO):
        return capi.load_lib(capi.REF_SO), "reference"
    return capi.load_lib(capi.ORACLE_SO), "port"


def cpu_step(lib, trace, prefix):
    e = capi.PolyEngine(lib, trace.dim)
    t0 = time.perf_counter()
    rcs = P.replay(e, trace, upto=prefix)
    dt = time.perf_counter() - t0
    cuts = len(rcs) - sum(rcs)
    live = int(np.unpackbits(np.ctypeslib.as_array(e.args.primal.used, shape=((e.args.primal.cnt + 63) // 64,)).view(np.uint8)).sum())
    e.kill()
    return cuts, dt, live


def run_reference(a, trace):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, kind = load_cpu_engine()
    prefix = min(a.ref_prefix, len(trace))
    for _ in range(a.warmup):
        cpu_step(lib, trace, prefix)
    cuts = 0
    t = 0.0
    live = 0
    for _ in range(a.steps):
        c, dt, live = cpu_step(lib, trace, prefix)
        cuts += c
        t += dt
    v = cuts / t
    line = {
        "impl": "reference", "metric": "halfspace cuts/sec", "value": v, "unit": "cuts/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pure H->V enumeration, random tangent polytope in R^{a.dim}, {a.halfspaces} halfspaces, seed {a.seed}",
                   "sample": f"first {prefix} of {a.halfspaces} halfspaces ({live} live vertices at the end of the sample)"},
        "cpu_baseline": {"value": v, "unit": "cuts/s", "cores": 1, "kind": kind,
                         "sample": f"first {prefix} of {a.halfspaces} halfspaces, single thread (the reference engine is single-threaded; {os.cpu_count()} host cores present)"},
        "e2e": {"value": v, "unit": "cuts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- B200 arm
# DRAM bytes of one K1 launch from the ncu --set full capture in profiles/ (same workload only)
PROFILED_K1_TRAFFIC = {(6, 6400, 20261018): 49.2e6}


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_prefix_state(trace, prefix):
    """State of the unmodified reference engine after the first `prefix` halfspaces (+ its timing)."""
    lib, kind = load_cpu_engine()
    e = capi.PolyEngine(lib, trace.dim)
    t0 = time.perf_counter()
    rcs = P.replay(e, trace, upto=prefix)
    dt = time.perf_counter() - t0
    st = e.state()
    e.kill()
    return st, rcs, dt, kind


def run_b200(a, trace):
    import ctypes as C
    import torch
    from bensolve_b200 import invariants as INV
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load_product()
    lib.b200_set_device.argtypes = [C.c_int]
    lib.b200_set_device(local)
    if world > 1:
        from bensolve_b200 import dist as bdist
        bdist.init_comm(lib)          # state replicated, classification sharded by row range
    d, n = trace.dim, len(trace)
    dev = torch.device("cuda", local)
    d_vals = torch.from_numpy(np.ascontiguousarray(trace.vals[d:])).to(dev)      # inputs resident in HBM
    torch.cuda.synchronize()
    sizes = {"rows": 0, "inc": 0, "adj": 0}
    prefix = min(a.ref_prefix, n)

    def fresh_engine(reserve=True):
        e = capi.PolyEngine(lib, d)
        if reserve and sizes["rows"]:
            e.reserve(sizes["rows"], sizes["inc"], sizes["adj"])
        for i in range(d):
            e.add(trace.vals[i], 0)
        assert e.init_approx() == 0
        return e

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # A step is timed from the first cut after poly__intl_apprx to the return of the last cut (host mirror
    # coherent), bracketed by barrier + synchronize.  Creating the polytope (poly__initialise, b200_poly_reserve =
    # every device and host allocation, the d start halfspaces, poly__intl_apprx) and poly__kill lie between
    # steps: cudaMalloc / cudaFree of several GB vary by 10-200 ms from call to call and are not the cut path.
    def step_value(count=None, keep=False):
        e = fresh_engine()
        barrier()
        t0 = time.perf_counter()
        e.add_batch_device(d_vals.data_ptr(), 0, (n if count is None else count) - d)
        barrier()
        dt = time.perf_counter() - t0
        st = e.stats()
        if keep:
            return e, st, dt
        e.kill()
        return None, st, dt

    def step_e2e(count=None, keep=False, reserve=True, batch=False):
        e = fresh_engine(reserve)
        upto = n if count is None else count
        barrier()
        t0 = time.perf_counter()
        if batch:
            e.add_batch(trace.vals[d:upto])   # host buffers in, one call: upload, device-resident cuts, mirror download
        else:
            e.add_each(trace.vals[d:upto])    # the C caller's loop: one poly__add_vrtx per halfspace, host buffers
        barrier()
        dt = time.perf_counter() - t0
        st = e.stats()
        if keep:
            return e, st, dt
        e.kill()
        return None, st, dt

    def step_vertenum():
        """cone_vertenum's call pattern (bslv_algs.c:331-350) through the unchanged API, no b200_* call at all: every halfspace
        queued with poly__add_vrtx, then poly__intl_apprx (which re-adds them, bslv_poly.c:190-197 -- as one device-resident batch)."""
        e = capi.PolyEngine(lib, d)
        barrier()
        t0 = time.perf_counter()
        e.add_each(trace.vals)            # (b200_poly_add_each = the caller's loop around poly__add_vrtx; before initialisation it only queues)
        assert e.init_approx() == 0
        barrier()
        dt = time.perf_counter() - t0
        st = e.stats()
        e.kill()
        return st, dt

    # ---- what an UNCHANGED caller gets: first large polytope of the process, no b200_poly_reserve (bslv_algs.c never
    # calls it), no recycled host block; only the CUDA context exists (a 3-d polytope of 40 halfspaces ran before,
    # as bensolve's own cone_vertenum does before its main loop)
    tiny = P.tangent_polytope(3, 40, 1)
    e0 = capi.PolyEngine(lib, 3)
    P.replay(e0, tiny)
    e0.kill()
    _, st_cold, dt_cold = step_e2e(reserve=False)

    # warm-up (also reveals the capacities to reserve, so the timed steps do not re-allocate)
    for w in range(max(a.warmup, 1)):
        _, st, _ = step_value()
        sizes.update(rows=int(st["slots"] * 1.25) + 65536, inc=int(st["slots"] * (d + 2)) + (1 << 20), adj=int(st["slots"] * (d + 2)) + (1 << 20))
    for _ in range(max(a.warmup - 1, 0)):
        step_e2e()

    with ClockSampler(local) as clk:
        t_value = t_e2e = t_e2e_batch = 0.0
        cuts_v = launches_v = evals_v = 0
        for k in range(a.steps):
            eng, st, dt = step_value(keep=(k == a.steps - 1))
            t_value += dt
            cuts_v += st["cuts"]; launches_v += st["kernel_launches"]; evals_v += st["vertex_evals"]
        eng_keep = eng
        cuts_e = launches_e = 0
        eng_e2e = None
        for k in range(a.steps):
            eng_e2e, st_e, dt = step_e2e(keep=(k == a.steps - 1))
            t_e2e += dt
            cuts_e += st_e["cuts"]; launches_e += st_e["kernel_launches"]
        cuts_eb = launches_eb = 0
        for k in range(a.steps):
            _, st_b, dt = step_e2e(batch=True)
            t_e2e_batch += dt
            cuts_eb += st_b["cuts"]; launches_eb += st_b["kernel_launches"]
        t_vn = 0.0
        cuts_vn = launches_vn = 0
        step_vertenum()
        for k in range(a.steps):
            st_v, dt = step_vertenum()
            t_vn += dt
            cuts_vn += st_v["cuts"]; launches_vn += st_v["kernel_launches"]
        eng = eng_keep
        # K1 alone on the final polytope of the last value step
        st_final = eng.stats()
        hp = np.append(trace.vals[n // 2] * 1.0000001, -1.0)
        ms_flush = eng.classify_bench(hp, a.classify_iters, True)
        ms_l2 = eng.classify_bench(hp, a.classify_iters, False)
        # both engines on the SAME prefix of the trace as the CPU reference
        t_pv = t_pe = 0.0
        cuts_p = 0
        step_value(count=prefix); step_e2e(count=prefix)
        for k in range(a.steps):
            _, stp, dt = step_value(count=prefix)
            t_pv += dt
            cuts_p += stp["cuts"]
            _, _, dt = step_e2e(count=prefix)
            t_pe += dt
    clocks = clk.summary()

    # max over ranks (strong scaling: every rank takes part in every cut)
    if dist is not None:
        tt = torch.tensor([t_value, t_e2e, t_e2e_batch, t_pv, t_pe, dt_cold, t_vn], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_value, t_e2e, t_e2e_batch, t_pv, t_pe, dt_cold, t_vn = (float(x) for x in tt)

    # ---------------- parity gate (before any number counts)
    parity = {"ok": False}
    failure = None
    try:
        # (1) full sequence, both paths: properties of a simple polytope + the two paths give the same polytope
        import gc
        snap = INV.Snapshot(eng)             # (one snapshot at a time: ~10 GB of host arrays at 10^7 vertices, per rank)
        inv_b, dg_b = INV.check_polytope(snap), INV.digest(snap)
        del snap
        gc.collect()
        snap = INV.Snapshot(eng_e2e)
        inv_e, dg_e = INV.check_polytope(snap), INV.digest(snap)
        del snap
        gc.collect()
        assert dg_b == dg_e, "device-resident batch path and per-call path end in different polytopes"
        assert inv_b["facets"] == n, "a tangent halfspace was lost"
        # (2) the prefix the CPU reference can do: bit-exact comparison with the unmodified bslv_poly.c, both paths
        eb, _, _ = step_value(count=prefix, keep=True)
        ee, _, _ = step_e2e(count=prefix, keep=True)
        dg_pb, dg_pe = INV.digest(INV.Snapshot(eb)), INV.digest(INV.Snapshot(ee))
        ref_kind, ref_dt, ref_cuts, ref_live = None, None, None, None
        if rank == 0:
            ref_state, ref_rcs, ref_dt, ref_kind = reference_prefix_state(trace, prefix)
            ref_cuts = len(ref_rcs) - sum(ref_rcs)
            ref_live = len(ref_state.incidence)
            capi.compare_states(ref_state, eb.state(), exact_coords=True)
            capi.compare_states(ref_state, ee.state(), exact_coords=True)
        eb.kill(); ee.kill()
        assert dg_pb == dg_pe
        ranks_identical = True
        if dist is not None:
            got = [None] * world
            dist.all_gather_object(got, (dg_b, dg_pb))
            ranks_identical = all(g == got[0] for g in got)
            assert ranks_identical, "ranks hold different polytopes"
        parity = {
            "ok": True,
            "prefix": {"halfspaces": prefix, "checker": "oracle/_ref/libref_poly.so (unmodified bslv_poly.c)" if ref_kind == "reference" else "oracle/libpoly_oracle.so",
                       "live_vertices": ref_live, "batch_path": "identical (structure after canonical sorting, coordinates bit-exact)",
                       "per_call_path": "identical (structure after canonical sorting, coordinates bit-exact)"},
            "full_sequence": {"batch_path": inv_b, "per_call_path": inv_e, "batch_vs_per_call": "identical", "sha256": dg_b,
                              "properties": "every vertex on >= d facets with >= d neighbours (equality except the counted non_simple_vertices, copies of vertices within 1e-9 of a later hyperplane); adjacency symmetric; adjacent vertices share d-1 facets; "
                                            "facet sets unique; facet lists = transpose of incidence lists; E = V*d/2; every vertex tight on its facets (1e-7), "
                                            "sampled vertices feasible for every halfspace and tight on no other"},
            "ranks_identical": ranks_identical,
        }
    except AssertionError as ex:
        failure = str(ex) or repr(ex)
        parity = {"ok": False, "error": failure}
        ref_dt = ref_cuts = ref_live = ref_kind = None
    eng.kill()
    eng_e2e.kill()

    peak, peak_src = measured_peak()
    n_live = st_final["live_vertices"]
    alg_bytes = n_live * (8 * d + 1)
    k1_achieved = alg_bytes / (ms_flush * 1e-3) / 1e9
    per_step = st_final
    seq_bytes = int(per_step["algorithmic_bytes"])
    seq_achieved = seq_bytes / (t_value / a.steps) / 1e9
    # bytes crossing PCIe per e2e step: d doubles in per call; the packed delta back per call
    h2d = (n - d) * 8 * (d + 1)
    d2h_step = int(128 * (n - d) + per_step["slots"] * (8 * d + 5) + 4 * per_step["slots"])
    evals_per_s = evals_v / t_value

    line = {
        "metric": "halfspace cuts/sec", "value": cuts_v / t_value, "unit": "cuts/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_value / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"pure H->V enumeration, random tangent polytope in R^{d}, {n} halfspaces, seed {a.seed}",
            "live_vertices": int(n_live), "slots": int(per_step["slots"]), "facets": int(per_step["facets"]),
            "l2": "step: coordinates (%.0f MB) stay L2-resident across cuts, inherent to the workload; roofline.k1: L2 flushed (read sweep over 256 MB) before every timed K1 launch" % (n_live * 8 * d / 1e6),
            "multi_gpu": (f"{world} ranks, one process per GPU: state replicated; look-ahead passes split by row group and exchanged over peer-mapped memory from 3M rows, "
                          f"K1 of a per-call cut split + NCCL all-gather from 4M rows, replicated (nothing exchanged) below; rest of the cut replicated" if world > 1 else "single"),
            "timed_region": "first cut after poly__intl_apprx .. last cut returned with a coherent host mirror; polytope creation (poly__initialise, b200_poly_reserve with the warm-up's counts = all device and host allocation, start simplex) and poly__kill lie between steps, untimed; e2e_unchanged_caller is the figure without any of that help",
        },
        "vertex_evals_per_s": evals_per_s,
        "vertex_evals_frac": evals_per_s * (8 * d + 1) / 1e9 / peak,
        "e2e": {"value": cuts_e / t_e2e, "unit": "cuts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_step,
                "ms_per_step": 1e3 * t_e2e / a.steps, "call": "poly__add_vrtx per halfspace (host buffers in, coherent host mirror after every call)"},
        "e2e_unchanged_caller": {"value": st_cold["cuts"] / dt_cold, "unit": "cuts/s", "ms_per_step": 1e3 * dt_cold,
                                 "what": "first polytope of the process through poly__add_vrtx: no b200_poly_reserve, no recycled host block, all growth inside the timed region"},
        "e2e_batch": {"value": cuts_eb / t_e2e_batch, "unit": "cuts/s", "ms_per_step": 1e3 * t_e2e_batch / a.steps,
                      "h2d_bytes_per_step": (n - d) * 8 * d, "d2h_bytes_per_step": int(per_step["slots"] * (8 * d + 9) + 4 * (n - d)),
                      "call": "b200_poly_add_batch (extension): all halfspaces in one call from host memory, host mirror coherent at return"},
        "e2e_vertenum": {"value": cuts_vn / t_vn, "unit": "cuts/s", "ms_per_step": 1e3 * t_vn / a.steps,
                         "h2d_bytes_per_step": n * 8 * d, "d2h_bytes_per_step": int(per_step["slots"] * (8 * d + 9) + 4 * n),
                         "call": "the UNCHANGED API the way bensolve enumerates vertices (cone_vertenum, bslv_algs.c:331-350): every halfspace queued with poly__add_vrtx, then poly__intl_apprx; "
                                 "no b200_* extension, no reserve; timed from the first poly__add_vrtx to the return of poly__intl_apprx (growth included)"},
        "gpu_launches": int(launches_v + launches_e + launches_eb + launches_vn),
        "roofline": {"bound": "hbm", "kernel": "whole cut sequence (SURVEY 8(d): sum over cuts of B_c / T_total); dominant kernels by the launch list in profiles/: see roofline.launch_list",
                     "achieved": seq_achieved, "peak": peak, "unit": "GB/s", "frac": seq_achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "sequence_algorithmic_bytes": seq_bytes, "bytes_per_cut": seq_bytes / max(1, per_step["cuts"]),
                     "rows_scanned_over_vertex_evals": per_step["rows_scanned"] / max(1, per_step["vertex_evals"]),
                     "launch_list": "profiles/r02_launches_wave.csv",
                     "k1": {"kernel": f"k_classify_lists<{d},false>", "achieved": k1_achieved, "frac": k1_achieved / peak,
                            "algorithmic_bytes_per_launch": int(alg_bytes), "ms_per_launch": ms_flush, "ms_per_launch_l2_resident": ms_l2,
                            "achieved_l2_resident": alg_bytes / (ms_l2 * 1e-3) / 1e9 if ms_l2 > 0 else None,
                            "traffic": PROFILED_K1_TRAFFIC.get((d, n, a.seed)),
                            "traffic_source": "profiles/r01_k1_summary.md (ncu dram__bytes_read.sum + dram__bytes_write.sum)" if (d, n, a.seed) in PROFILED_K1_TRAFFIC else None}},
        "parity": parity,
        "clocks": clocks,
    }
    if rank == 0:
        # CPU baseline beside it: unmodified reference engine on a bounded prefix, rank 0, 1 thread; the GPU on the same prefix
        if ref_dt is None:
            cpu_lib, ref_kind = load_cpu_engine()
            ref_cuts, ref_dt, ref_live = cpu_step(cpu_lib, trace, prefix)
        line["cpu_baseline"] = {"value": ref_cuts / ref_dt, "unit": "cuts/s", "cores": 1, "kind": ref_kind,
                                "sample": f"first {prefix} of {n} halfspaces ({ref_live} live vertices at the end), {ref_dt:.1f} s, single thread of {os.cpu_count()}"}
        line["same_prefix"] = {"halfspaces": prefix, "cpu_reference_cuts_per_s": ref_cuts / ref_dt,
                               "gpu_value_cuts_per_s": cuts_p / t_pv, "gpu_e2e_cuts_per_s": cuts_p / t_pe,
                               "note": "all three on the first %d halfspaces of the trace (identical inputs, identical result: parity.prefix)" % prefix}
        if a.shapes and world == 1:
            line["shapes"] = run_shapes(lib, a, peak)
        ex = os.path.join(REPO, "oracle", "_ref", "ex", a.solve_time + ".vlp")
        if a.solve_time and world == 1 and os.path.exists(ex) and os.path.exists(os.path.join(REPO, "oracle", "_ref", "libbensolve_host.so")):
            # third component of the metric: bensolve's own "CPU time" for a whole solve, reference engine vs B200 engine
            sys.path.insert(0, os.path.join(REPO, "tools"))
            import solve_time
            try:
                line["solve_time"] = solve_time.compare(ex)
            except Exception as exc:      # (the LP stand-in needs scipy; never let it take the bench line down)
                line["solve_time"] = {"error": str(exc)[-300:]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if failure is not None:
        sys.stderr.write("bench.py: PARITY GATE FAILED: " + failure + "\n")
        sys.exit(1)


def run_shapes(lib, a, peak):
    """vertex-evals/s and its HBM fraction on the d=3 and d=5 shapes of the metric (25 and 41 bytes per
    vertex-eval): one device-resident sequence each + K1 alone on the final polytope."""
    out = []
    for spec in a.shapes.split(","):
        dd, nn = (int(x) for x in spec.split(":"))
        tr = P.tangent_polytope(dd, nn, a.seed)
        best = None
        for rep in range(2):
            e = capi.PolyEngine(lib, dd)
            if best:
                e.reserve(int(best["slots"] * 1.25) + 65536, int(best["slots"] * (dd + 2)) + (1 << 20), int(best["slots"] * (dd + 2)) + (1 << 20))
            for i in range(dd):
                e.add(tr.vals[i], 0)
            assert e.init_approx() == 0
            t0 = time.perf_counter()
            e.add_batch(tr.vals[dd:])
            dt = time.perf_counter() - t0
            st = e.stats()
            st["dt"] = dt
            if rep == 1:
                hp = np.append(tr.vals[nn // 2] * 1.0000001, -1.0)
                ms = e.classify_bench(hp, a.classify_iters, True)
                st["k1_ms"] = ms
            e.kill()
            best = st
        bpe = 8 * dd + 1
        out.append({"dim": dd, "halfspaces": nn, "live_vertices": int(best["live_vertices"]), "bytes_per_vertex_eval": bpe,
                    "cuts_per_s": best["cuts"] / best["dt"], "vertex_evals_per_s": best["vertex_evals"] / best["dt"],
                    "vertex_evals_frac": best["vertex_evals"] / best["dt"] * bpe / 1e9 / peak,
                    "sequence_frac": best["algorithmic_bytes"] / best["dt"] / 1e9 / peak,
                    "k1_alone_frac": best["live_vertices"] * bpe / (best["k1_ms"] * 1e-3) / 1e9 / peak})
    return out


def main():
    a = parse()
    trace = P.tangent_polytope(a.dim, a.halfspaces, a.seed)
    if a.impl == "reference":
        run_reference(a, trace)
    else:
        run_b200(a, trace)


if __name__ == "__main__":
    main()
