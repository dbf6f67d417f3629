"""Compile the plain-C oracle (oracle/c/*.c) into oracle/liboracle_c.so with gcc.  TEST INFRASTRUCTURE.

The reference is pure Python (no C/C++ sources to compile into oracle/_ref), so this library is the
'port' CPU baseline: bench.py's cpu_baseline / --impl reference legs and tests/test_oracle_c.py use it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_c.so")
SRCS = [os.path.join(HERE, "c", f) for f in ("mfcc_oracle.c", "nds_oracle.c")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRCS):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB, *SRCS, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stderr}")
    return LIB


def load() -> C.CDLL:
    lib = C.CDLL(build())
    lib.cmoop_oracle_mfcc.restype = C.c_long
    lib.cmoop_oracle_mfcc.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.cmoop_oracle_num_threads.restype = C.c_int
    lib.cmoop_oracle_nds.restype = C.c_int
    lib.cmoop_oracle_nds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
    lib.cmoop_oracle_crowding.restype = None
    lib.cmoop_oracle_crowding.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    return lib


def mfcc(wave, *, sample_rate=16000, frame_length=640, hop=320, n_fft=1024, n_mels=40, n_mfcc=40, f_min=0.0,
         f_max=8000.0, log_floor=1e-10):
    import numpy as np
    lib = load()
    wave = np.ascontiguousarray(wave, np.float32)
    n_clips, n_samples = wave.shape
    frames = 0 if n_samples < frame_length else 1 + (n_samples - frame_length) // hop
    n_out = n_mfcc if n_mfcc > 0 else n_mels
    out = np.empty((n_clips, frames, n_out), np.float64)
    rc = lib.cmoop_oracle_mfcc(wave.ctypes.data, n_clips, n_samples, sample_rate, frame_length, hop, n_fft, n_mels,
                               n_mfcc, f_min, f_max, log_floor, out.ctypes.data)
    if rc < 0:
        raise ValueError("bad MFCC configuration")
    return out


def nds(objs, cv, lam):
    import numpy as np
    lib = load()
    objs = np.ascontiguousarray(objs, np.float64)
    cv = np.ascontiguousarray(cv, np.float64)
    n, m = objs.shape
    rank = np.empty(n, np.int32)
    order = np.empty(n, np.int32)
    foff = np.empty(n + 1, np.int32)
    nf = lib.cmoop_oracle_nds(objs.ctypes.data, cv.ctypes.data, n, m, float(lam), rank.ctypes.data, order.ctypes.data,
                              foff.ctypes.data)
    return [order[foff[i]:foff[i + 1]].tolist() for i in range(nf)], rank


def crowding(objs, front, eps=1e-6, mode=0):
    import numpy as np
    lib = load()
    objs = np.ascontiguousarray(objs, np.float64)
    idx = np.ascontiguousarray(front, np.int32)
    out = np.empty(len(idx), np.float64)
    lib.cmoop_oracle_crowding(objs.ctypes.data, objs.shape[1], idx.ctypes.data, len(idx), eps, mode, out.ctypes.data)
    return out


if __name__ == "__main__":
    print(build(force=True))
