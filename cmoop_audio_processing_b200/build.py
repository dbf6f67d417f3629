"""In-tree build of libcmoop_b200.so (sm_100a only) with plain nvcc.

``python -m cmoop_audio_processing_b200.build`` or ``build_library()``.  The shared
library is written next to this file so it travels with the repo snapshot to the
GPU box; objects are cached under csrc/_build/ and rebuilt when a source or header
is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "_build")
LIB_PATH = os.path.join(HERE, "libcmoop_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-DCMOOP_BUILD", "-Xptxas", "-v",
]
# translation units whose fp64 arithmetic must round exactly like CPython (no FMA contraction)
NO_FMAD = {"nds.cu", "quality.cu"}


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; libcmoop_b200 cannot be built")
    return path


def _sources() -> list[str]:
    out = []
    for root, _dirs, files in os.walk(CSRC):
        if os.path.basename(root) == "_build":
            continue
        out += [os.path.join(root, f) for f in sorted(files) if f.endswith(".cu")]
    return sorted(out)


def _headers_mtime() -> float:
    newest = 0.0
    for root, _dirs, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, f)))
    inc = os.path.join(os.path.dirname(HERE), "include", "cmoop_b200.h")
    if os.path.exists(inc):
        newest = max(newest, os.path.getmtime(inc))
    return newest


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_time = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
        obj = os.path.join(OBJ_DIR, rel[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)
        if stale:
            flags = list(NVCC_FLAGS)
            if os.path.basename(src) in NO_FMAD:
                flags.append("-fmad=false")
            jobs.append((src, [nvcc, *flags, "-c", src, "-o", obj]))

    def run(job):
        src, cmd = job
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return src, res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for src, log in pool.map(run, jobs):
                if verbose:
                    print(f"--- {os.path.relpath(src, HERE)}\n{log}")
                with open(os.path.join(OBJ_DIR, os.path.basename(src) + ".ptxas.log"), "w") as fh:
                    fh.write(log)
    need_link = bool(jobs) or not os.path.exists(LIB_PATH) or \
        any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)
    if need_link:
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True)
    print("built", path)
