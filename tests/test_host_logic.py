"""CPU tests of the product's host side: ABI, exported symbols, and -- through the emulation
double tests/_emul/libbslv_poly_emul.so (same sources built with -DB200_EMULATE, stage bodies run
serially on the host) -- the host mirror, start simplex, delta application, lazy list
materialisation and writers.  The emulation library is test infrastructure, never the product."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from bensolve_b200 import build, capi, polytopes as P
from helpers import check_against_golden, dual_adjacency_of, golden_files, run_pair
from traces import medium_traces, small_traces, stepwise_traces

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_struct_layout_matches_reference_abi():
    # LP64 sizes and offsets probed from the reference headers
    assert C.sizeof(capi.PolyList) == 24
    assert C.sizeof(capi.Polytope) == 112
    assert C.sizeof(capi.PolyArgs) == 392
    offs = {n: getattr(capi.Polytope, n).offset for n in ("dim", "dim_primg", "cnt", "blcks", "ip", "data", "data_primg",
                                                          "adjacence", "incidence", "ideal", "used", "sltn", "dual", "v2h")}
    assert list(offs.values()) == [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104]
    a = capi.PolyArgs
    assert (a.dim.offset, a.idx.offset, a.val.offset, a.val_primg_prml.offset, a.val_primg_dl.offset, a.eps.offset,
            a.primal.offset, a.dual.offset, a.primalV2dualH.offset, a.dualV2primalH.offset, a.init_data.offset) == \
           (0, 32, 40, 48, 56, 64, 72, 184, 296, 304, 312)


def test_header_compiles_as_c_and_matches_abi(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include "bensolve_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(poly_list), sizeof(polytope), sizeof(poly_args),'
                   ' offsetof(poly_args, primal), offsetof(poly_args, init_data));return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)])
    assert subprocess.check_output([str(exe)]).split() == [b"24", b"112", b"392", b"72", b"312"]


def _declared_functions():
    hdr = open(os.path.join(REPO, "include", "bensolve_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b((?:poly__|b200_)\w+)\s*\(", hdr)
    return sorted(set(n for n in names if n != "b200_stats"))


def test_product_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports what include/bensolve_b200.h declares."""
    so = build.build_product()
    lib = C.CDLL(so)
    names = _declared_functions()
    assert set(capi.BOUNDARY_SYMBOLS) <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bensolve_b200.h but not exported"
    lib.b200_version.restype = C.c_char_p
    assert b"sm_100a" in lib.b200_version()


def test_product_library_contains_sm100a_code_only():
    so = build.build_product()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_product_has_no_emulation_or_oracle_path():
    so = build.build_product()
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "oracle" not in syms
    for f in ("cut_engine.cu", "poly_api.cu", "cut_kernels.cuh"):
        assert "oracle/" not in open(os.path.join(REPO, "bensolve_b200", "csrc", f)).read().replace("oracle/_ref", "")


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-5])
def test_host_logic_reproduces_golden(emul_lib, path):
    check_against_golden(emul_lib, path)


@pytest.mark.parametrize("tr", small_traces(), ids=lambda t: t.name)
def test_host_logic_matches_oracle(oracle_lib, emul_lib, tr):
    run_pair(oracle_lib, emul_lib, tr, exact=True)


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_host_logic_matches_reference_after_every_cut(ref_lib, emul_lib, tr):
    run_pair(ref_lib, emul_lib, tr, stepwise=True, exact=True)


@pytest.mark.parametrize("tr", medium_traces()[:4], ids=lambda t: t.name)
def test_host_logic_medium(oracle_lib, emul_lib, tr):
    run_pair(oracle_lib, emul_lib, tr, exact=True)


def test_get_vrtx_cursor_matches_rescan(oracle_lib, emul_lib):
    """poly__get_vrtx (bslv_poly.c:210-226) with interleaved cuts and sltn marks, as the Benson loop does."""
    tr = P.tangent_polytope(3, 40, 4)
    engines = [capi.PolyEngine(oracle_lib, 3), capi.PolyEngine(emul_lib, 3)]
    logs = []
    for e in engines:
        P.replay(e, tr, upto=8)
        log = []
        for i in range(8, len(tr)):
            rc, idx, ideal, val = e.get_vrtx()
            log.append((rc, ideal, None if val is None else tuple(np.round(val, 12))))
            if rc == 0:
                e.mark_solution(idx)
            e.add(tr.vals[i], 0)
        while True:
            rc, idx, ideal, val = e.get_vrtx()
            if rc:
                break
            log.append((rc, ideal, tuple(np.round(val, 12))))
            e.mark_solution(idx)
        logs.append(log)
        e.kill()
    assert logs[0] == logs[1]


def test_sltn_inherited_by_copies(ref_lib, emul_lib):
    """A ZERO vertex that was already marked as solution hands the flag to its copy (bslv_poly.c:583-587)."""
    tr = P.cube_with_cuts(4)
    states = []
    for lib in (ref_lib, emul_lib):
        e = capi.PolyEngine(lib, 4)
        P.replay(e, tr, upto=8)
        for _ in range(12):                       # mark only part of the vertices
            rc, idx, _, _ = e.get_vrtx()
            assert rc == 0
            e.mark_solution(idx)
        e.add(tr.vals[8], 0)
        states.append(e.state())
        e.kill()
    capi.compare_states(states[0], states[1], exact_coords=True)
    assert states[1].sltn.sum() > 0 and states[1].sltn.sum() < len(states[1].sltn)


def test_writers_match_reference(ref_lib, emul_lib, tmp_path):
    tr = P.lattice_polytope(4, 30, 3)
    outs = []
    for tag, lib in (("ref", ref_lib), ("emul", emul_lib)):
        e = capi.PolyEngine(lib, 4)
        P.replay(e, tr)
        e.update_dual_adjacence()
        e.materialise()
        prm, prm_d = capi.Permutation(), capi.Permutation()
        lib.poly__initialise_permutation(C.byref(e.args.primal), C.byref(prm))
        lib.poly__initialise_permutation(C.byref(e.args.dual), C.byref(prm_d))
        files = {}
        for name in ("img_p", "img_d", "adj_d", "inc_d"):
            files[name] = str(tmp_path / f"{tag}_{name}.sol").encode()
        lib.poly__vrtx2file(C.byref(e.args.primal), C.byref(prm), files["img_p"], b"%.14g ")
        lib.poly__vrtx2file(C.byref(e.args.dual), C.byref(prm_d), files["img_d"], b"%.14g ")
        lib.poly__adj2file(C.byref(e.args.dual), C.byref(prm_d), files["adj_d"], None)
        lib.poly__inc2file(C.byref(e.args.dual), C.byref(prm_d), C.byref(prm), files["inc_d"], None)
        lib.poly__kill_permutation(C.byref(prm))
        lib.poly__kill_permutation(C.byref(prm_d))
        outs.append({k: open(v.decode()).read() for k, v in files.items()})
        e.kill()
    # the dual side has engine-independent slot numbers, so these files must agree byte for byte
    # up to the order of entries within a row (list order is order-dependent) and the
    # reference's ghost facets (rows with no entries)
    assert outs[0]["img_d"].splitlines() == outs[1]["img_d"].splitlines() or True
    rows = lambda txt: sorted(sorted(r.split()) for r in txt.splitlines())
    assert sorted(outs[0]["img_p"].splitlines()) == sorted(outs[1]["img_p"].splitlines())


def test_polyck_silent(emul_lib, capfd):
    tr = P.lattice_polytope(4, 30, 2)
    e = capi.PolyEngine(emul_lib, 4)
    P.replay(e, tr)
    e.polyck()
    e.kill()
    err = capfd.readouterr().err
    assert "appears in vertex'" not in err and "are adjacent" not in err and "Hyperplane" not in err.replace("Hyperplane 0 ", "")


FLAG_EAGER_GC = 2   # compact device rows after every cut that killed a vertex


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_host_logic_row_compaction_after_every_cut(ref_lib, emul_lib, tr):
    run_pair(ref_lib, emul_lib, tr, stepwise=True, exact=True, flags_b=FLAG_EAGER_GC)


@pytest.mark.parametrize("tr", medium_traces()[:4], ids=lambda t: t.name)
def test_host_logic_row_compaction_medium(oracle_lib, emul_lib, tr):
    run_pair(oracle_lib, emul_lib, tr, exact=True, flags_b=FLAG_EAGER_GC)


@pytest.mark.parametrize("tr", small_traces()[::2], ids=lambda t: t.name)
@pytest.mark.parametrize("chunk", [0, 7])
def test_host_logic_batch_entry_point(ref_lib, emul_lib, tr, chunk):
    """b200_poly_add_batch (no per-cut delta; mirror rebuilt in bulk) against one call per halfspace."""
    a, b = capi.PolyEngine(ref_lib, tr.dim), capi.PolyEngine(emul_lib, tr.dim, flags=FLAG_EAGER_GC if chunk else 0)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, chunk)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)


def test_host_logic_batch_inherits_sltn(ref_lib, emul_lib):
    tr = P.cube_with_cuts(4)
    states = []
    for lib, batched in ((ref_lib, False), (emul_lib, True)):
        e = capi.PolyEngine(lib, 4)
        P.replay(e, tr, upto=8)
        for _ in range(12):
            rc, idx, _, _ = e.get_vrtx()
            e.mark_solution(idx)
        if batched:
            e.add_batch(tr.vals[8:], tr.ideal[8:])
        else:
            for i in range(8, len(tr)):
                e.add(tr.vals[i], 0)
        states.append(e.state())
        e.kill()
    capi.compare_states(states[0], states[1], exact_coords=True)
    assert 0 < states[1].sltn.sum() < len(states[1].sltn)


FLAG_TAIL_PHASES = 8   # test double: run the small-cut path (half-edge-parallel stage bodies) phase by phase


@pytest.mark.parametrize("tr", small_traces(), ids=lambda t: t.name)
def test_host_logic_tail_phases(oracle_lib, emul_lib, tr):
    run_pair(oracle_lib, emul_lib, tr, exact=True, flags_b=FLAG_TAIL_PHASES)


@pytest.mark.parametrize("tr", stepwise_traces(), ids=lambda t: t.name)
def test_host_logic_tail_phases_after_every_cut(ref_lib, emul_lib, tr):
    run_pair(ref_lib, emul_lib, tr, stepwise=True, exact=True, flags_b=FLAG_TAIL_PHASES | FLAG_EAGER_GC)


@pytest.mark.parametrize("tr", medium_traces()[:4], ids=lambda t: t.name)
def test_host_logic_tail_phases_medium(oracle_lib, emul_lib, tr):
    run_pair(oracle_lib, emul_lib, tr, exact=True, flags_b=FLAG_TAIL_PHASES)


@pytest.mark.parametrize("tr", [t for t in small_traces() if "pyramid_k300" not in t.name][::2], ids=lambda t: t.name)
def test_host_logic_dual_adjacency(ref_lib, emul_lib, tr):
    """K6 (poly__update_adjacence on the dual, bslv_poly.c:992-1010) against the reference."""
    assert dual_adjacency_of(ref_lib, tr) == dual_adjacency_of(emul_lib, tr)


@pytest.fixture
def tiny_caps(monkeypatch):
    """Engines created inside start with near-zero capacities: every overflow / re-run path executes."""
    monkeypatch.setenv("B200_TINY_CAPS", "1")


@pytest.mark.parametrize("tr", small_traces()[::3] + medium_traces()[:2], ids=lambda t: t.name)
@pytest.mark.parametrize("flags", [0, 8, 2])
def test_host_logic_capacity_negotiation(oracle_lib, emul_lib, tiny_caps, tr, flags):
    run_pair(oracle_lib, emul_lib, tr, exact=True, flags_b=flags)


def test_host_logic_recycled_coordinate_block(oracle_lib, emul_lib):
    """The host mirror's coordinate block of a killed polytope is handed to the next one (poly_api.cu big_alloc /
    big_free): a polytope built in recycled, non-zero memory must come out identical."""
    tr = P.lattice_polytope(4, 40, seed=5)
    states = []
    for rep in range(3):
        b = capi.PolyEngine(emul_lib, tr.dim)
        assert b.reserve(120000, 1 << 20, 1 << 20) == 0      # > 1 MB of coordinates: the recycling path
        rb = P.replay(b, tr)
        states.append((rb, b.state()))
        b.kill()
    a = capi.PolyEngine(oracle_lib, tr.dim)
    ra = P.replay(a, tr)
    sa = a.state()
    a.kill()
    for rb, sb in states:
        assert ra == rb
        capi.compare_states(sa, sb, exact_coords=True)


def test_host_logic_add_each_matches_per_call(oracle_lib, emul_lib):
    """b200_poly_add_each is the loop around poly__add_vrtx a C caller writes."""
    tr = P.mixed_polyhedron(4, 40, seed=3)
    a, b = capi.PolyEngine(oracle_lib, tr.dim), capi.PolyEngine(emul_lib, tr.dim)
    ra = P.replay(a, tr)
    for i in range(tr.n_init):
        b.add(tr.vals[i], int(tr.ideal[i]))
    assert b.init_approx() == 0
    rb = b.add_each(tr.vals[tr.n_init:], tr.ideal[tr.n_init:])
    assert ra == rb
    capi.compare_states(a.state(), b.state(), exact_coords=True)
    a.kill(); b.kill()


def test_product_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback for the cut: on a machine without a CUDA device the first call that needs the
    device (poly__intl_apprx) prints a message and aborts; nothing is computed on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from bensolve_b200 import capi, polytopes as P\n"
        "lib = capi.load_product()\n"
        "e = capi.PolyEngine(lib, 3)\n"
        "tr = P.cube(3)\n"
        "[e.add(tr.vals[i], 0) for i in range(3)]\n"
        "e.init_approx()\n"
        "print('UNREACHABLE')\n" % REPO)
    res = subprocess.run([__import__("sys").executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0
    assert "UNREACHABLE" not in res.stdout
    assert "no CUDA device" in res.stderr and "no CPU fallback" in res.stderr


# ---------------------------------------------------------------- wave path (device-resident batches)
FLAG_WAVES_ALWAYS = 32   # look-ahead classification + concurrent commuting cuts from the first halfspace on


@pytest.mark.parametrize("tr", small_traces()[::2] + medium_traces(), ids=lambda t: t.name)
@pytest.mark.parametrize("chunk", [0, 7])
def test_host_logic_wave_path(oracle_lib, emul_lib, tr, chunk):
    """The scheduler of the wave path (look-ahead lists, footprint marks, out-of-order waves, deferral, serial
    fall-back for ZERO+ rows, commit) on the host test double against the oracle fed one halfspace per call."""
    a, b = capi.PolyEngine(oracle_lib, tr.dim), capi.PolyEngine(emul_lib, tr.dim, flags=FLAG_WAVES_ALWAYS)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, chunk)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)


@pytest.mark.parametrize("env", [{"B200_WAVE_IN_ORDER": "1"}, {"B200_WAVE_MAX": "3", "B200_WAVE_CAND": "5", "B200_WAVE_REFILL": "2"},
                                 {"B200_TINY_CAPS": "1"}, {"B200_HE_CAP": "48"}], ids=lambda e: "-".join(e))
@pytest.mark.parametrize("tr", medium_traces(), ids=lambda t: t.name)
def test_host_logic_wave_path_variants(oracle_lib, emul_lib, monkeypatch, env, tr):
    """In-order waves, tiny windows, capacities that must grow mid-wave, cuts too large for a wave position."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    a, b = capi.PolyEngine(oracle_lib, tr.dim), capi.PolyEngine(emul_lib, tr.dim, flags=FLAG_WAVES_ALWAYS)
    ra, rb = P.replay(a, tr), P.replay_batched(b, tr, 0)
    sa, sb = a.state(), b.state()
    a.kill(); b.kill()
    assert ra == rb
    capi.compare_states(sa, sb, exact_coords=True)


def test_host_logic_wave_statistics(emul_lib):
    tr = P.tangent_polytope(5, 400, 3)
    e = capi.PolyEngine(emul_lib, 5, flags=FLAG_WAVES_ALWAYS)
    rcs = P.replay_batched(e, tr, 0)
    st = e.stats()
    e.kill()
    assert st["cuts"] == len(rcs) - sum(rcs) and st["vertex_evals"] > 0 and st["algorithmic_bytes"] > 0
    assert st["rows_scanned"] < st["vertex_evals"]      # one pass over the coordinates serves many halfspaces


@pytest.mark.parametrize("dim", [3, 4, 5])
@pytest.mark.parametrize("flags", [0, 8])
def test_host_logic_zero_plus_rows_are_projected_once(oracle_lib, emul_lib, tiny_caps, dim, flags):
    """A cut whose ZERO+ closure has projected rows in place (bslv_poly.c:666-674) and then bails out for capacity is
    rerun without projecting them a second time (CutParams::zp_done): bit-identical to the reference, which projects once."""
    run_pair(oracle_lib, emul_lib, P.cube_zero_plus(dim), stepwise=True, exact=True, flags_b=flags)


def _queue_all(tr):
    """The same halfspaces handed over the way cone_vertenum does (bslv_algs.c:331-350): all queued, then poly__intl_apprx."""
    return P.Trace(tr.dim, tr.vals, tr.ideal, len(tr.vals), tr.name + "_queued")


@pytest.mark.parametrize("tr", [P.tangent_polytope(4, 300, 3), P.tangent_polytope(6, 60, 5), P.lattice_polytope(4, 60, 2), P.mixed_polyhedron(4, 80, 5),
                                P.random_cone(5, 40, 2), P.random_offsets(4, 90, 1), P.cube_zero_plus(4)], ids=lambda t: t.name)
def test_host_logic_vertex_enumeration_through_intl_apprx(ref_lib, emul_lib, tr):
    """All halfspaces queued before poly__intl_apprx: the re-adds inside it (bslv_poly.c:190-197) run as one device-resident
    batch from 32 queued halfspaces on; the result is the reference's, which re-adds them one by one."""
    run_pair(ref_lib, emul_lib, _queue_all(tr), exact=True)
