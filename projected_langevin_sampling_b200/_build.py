"""Builds libpls_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Called by `__graft_entry__.build()`.  Objects are compiled in parallel (one translation unit per exponent depth of
the hot kernel) and cached by source hash under csrc/build/.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libpls_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
MAX_NKD = 7


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpls_b200.so cannot be built (there is no CPU fallback)")


def _source_hash() -> str:
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    for name in names:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(INCLUDE, "pls_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _units():
    units = [("pls_api.cu", [], "pls_api.o"), ("pls_aux.cu", [], "pls_aux.o"), ("pls_selector.cu", [], "pls_selector.o"), ("pls_step.cu", [], "pls_step.o"),
             ("pls_gen_gemm.cu", [], "pls_gen_gemm.o")]
    for k in range(1, MAX_NKD + 1):
        for role in range(5):  # forward epilogues PLS_EPI_* (0..3), backward (4)
            units.append(("pls_gen_gemm_inst.cu", [f"-DPLS_NKD={k}", f"-DPLS_ROLE={role}"], f"pls_gen_gemm_nkd{k}_role{role}.o"))
    return units


def _unit_hash(src: str, defs) -> str:
    """Hash of what one object depends on: its source, every header of csrc/ and include/, the flags."""
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [src]
    for name in names:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(INCLUDE, "pls_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + list(defs)).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of libpls_b200.so.  Objects are cached one by one under csrc/build/."""
    stamp = LIB + ".sha256"  # next to the library (csrc/build/ does not travel to the GPU box)
    digest = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)

    def compile_one(unit):
        src, defs, obj = unit
        out = os.path.join(BUILD, obj)
        unit_digest = _unit_hash(src, defs)
        if not force and not verbose and os.path.exists(out) and os.path.exists(out + ".sha256") and open(out + ".sha256").read() == unit_digest:
            return ""
        cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", "-o", out, os.path.join(CSRC, src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {defs}:\n{r.stdout}\n{r.stderr}")
        with open(out + ".sha256", "w") as f:
            f.write(unit_digest)
        return r.stderr

    units = _units()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
        logs = list(pool.map(compile_one, units))
    if verbose:
        print("\n".join(logs))
    objs = [os.path.join(BUILD, u[2]) for u in units]
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
