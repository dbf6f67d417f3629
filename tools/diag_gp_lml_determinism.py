"""Bitwise repeatability of cmoop_gp_lml_eval: the same (target, theta) problems evaluated repeatedly, alone and inside
batches, in different scratch slots."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmoop_audio_processing_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
bad = 0
for n in (33, 64, 144, 288):
    x = np.ascontiguousarray(rng.integers(0, 4, (n, 8)).astype(np.float64) + 0.01 * rng.standard_normal((n, 8)))
    y = np.ascontiguousarray(rng.standard_normal((4, n)))
    slots = 44
    h = C.c_void_p()
    _lib.check(lib.cmoop_gp_lml_create(_lib.ptr(x), n, 8, _lib.ptr(y), 4, 0, 1.5, 1e-10, slots, C.byref(h)), "create")
    th = np.ascontiguousarray(rng.uniform([-4, -2, -9], [6, 6, 2], size=(slots, 3)))
    tg = np.ascontiguousarray(np.arange(slots) % 4, np.int32)

    def run(slot, count, first):
        lml, grad = np.empty(count), np.empty((count, 3))
        _lib.check(lib.cmoop_gp_lml_eval(h, slot, count, _lib.ptr(th[first:first + count]), _lib.ptr(tg[first:first + count]),
                                         _lib.ptr(lml), _lib.ptr(grad)), "eval")
        return np.concatenate([lml[:, None], grad], axis=1)

    ref = run(0, slots, 0)
    for rep in range(20):
        if not np.array_equal(run(0, slots, 0).view(np.uint64), ref.view(np.uint64)):
            bad += 1
    for k in range(slots):                                  # alone, in a different slot
        one = run((k * 7) % slots, 1, k)
        if not np.array_equal(one.view(np.uint64), ref[k:k + 1].view(np.uint64)):
            bad += 1
    print(f"n={n}: mismatches so far {bad}", flush=True)
    lib.cmoop_gp_lml_destroy(h)
sys.exit(1 if bad else 0)
