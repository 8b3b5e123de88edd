"""Build recipes.

* :func:`build_product`  -- nvcc, sm_100a, in-tree ``bensolve_b200/libbslv_poly_b200.so``
  (cross-compiles on a machine without a GPU; the built file travels to the GPU box).
* :func:`build_oracle`   -- ``oracle/libpoly_oracle.so`` and, when the reference sources are
  present, ``oracle/_ref/libref_poly.so`` (test infrastructure).
* :func:`build_emulation` -- ``tests/_emul/libbslv_poly_emul.so``: the host-side test double of the
  device data-layout logic (g++ -DB200_EMULATE).  Only ``tests/`` loads it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
PRODUCT_SO = os.path.join(HERE, "libbslv_poly_b200.so")
EMUL_SO = os.path.join(REPO, "tests", "_emul", "libbslv_poly_emul.so")
SOURCES = ["poly_api.cu", "cut_engine.cu"]
HEADERS = ["cut_types.h", "cut_bodies.h", "cut_kernels.cuh", "cut_engine.h", "wave_bodies.h", "wave_kernels.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # the reference is built without FMA contraction
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-cudart", "static",
]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _deps():
    return ([os.path.join(CSRC, f) for f in SOURCES + HEADERS]
            + [os.path.join(REPO, "include", "bensolve_b200.h"), os.path.abspath(__file__)])


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_product(force: bool = False, verbose: bool = False) -> str:
    if not force and _newer(PRODUCT_SO, _deps()):
        return PRODUCT_SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", PRODUCT_SO] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return PRODUCT_SO


def build_emulation(force: bool = False) -> str:
    if not force and _newer(EMUL_SO, _deps()):
        return EMUL_SO
    os.makedirs(os.path.dirname(EMUL_SO), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
           "-DB200_EMULATE", "-Wl,-Bsymbolic", "-o", EMUL_SO] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ (emulation build) failed")
    return EMUL_SO


def build_oracle() -> None:
    res = subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "all"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("oracle build failed")


if __name__ == "__main__":
    build_oracle()
    print(build_product(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_emulation(force="--force" in sys.argv))
