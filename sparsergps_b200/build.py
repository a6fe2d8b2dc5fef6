"""Build libsrgp.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

    python -m sparsergps_b200.build [--force] [--verbose]

Objects go to build/obj (git-ignored); the library lands next to this file so it travels with `gpurun`.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libsrgp.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                     "-I" + os.path.join(ROOT, "include")]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]


def _sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cu") or f.endswith(".cpp"):
            out.append(os.path.join(CSRC, f))
    return out


def _headers_mtime():
    t = os.path.getmtime(os.path.join(ROOT, "include", "srgp.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h", ".inc")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def _compile(src, obj, verbose):
    if src.endswith(".cu"):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    else:
        cmd = ["g++"] + CXX_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("compile failed: " + src)
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _headers_mtime()
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda so: _compile(so[0], so[1], verbose), jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl", "-Xlinker", "-z", "-Xlinker", "defs"]   # undefined symbols fail the link, not the dlopen
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
