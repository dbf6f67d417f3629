"""Latency of one cmoop_gp_lml_eval call and throughput of concurrent slots (threads), per training-set size."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmoop_audio_processing_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in ([int(a) for a in sys.argv[1:]] or [64, 144, 288]):
        x = np.ascontiguousarray(rng.integers(0, 4, (n, 8)).astype(np.float64) + 0.01 * rng.standard_normal((n, 8)))
        y = np.ascontiguousarray(rng.standard_normal((4, n)))
        slots = 44
        h = C.c_void_p()
        _lib.check(lib.cmoop_gp_lml_create(_lib.ptr(x), n, 8, _lib.ptr(y), 4, 0, 1.5, 1e-10, slots, C.byref(h)), "create")
        th = np.tile(np.array([0.0, 0.0, np.log(0.1)]), (slots, 1))
        tg = np.ascontiguousarray(np.arange(slots) % 4, np.int32)
        lml, grad = np.empty(slots), np.empty((slots, 3))

        def one(slot, reps):
            l, g = np.empty(1), np.empty(3)
            for _ in range(reps):
                lib.cmoop_gp_lml_eval(h, slot, 1, _lib.ptr(th[slot]), _lib.ptr(tg[slot:slot + 1]), _lib.ptr(l), _lib.ptr(g))

        one(0, 3)
        t0 = time.perf_counter(); one(0, 20); t1 = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(5):
            lib.cmoop_gp_lml_eval(h, 0, slots, _lib.ptr(th), _lib.ptr(tg), _lib.ptr(lml), _lib.ptr(grad))
        tb = (time.perf_counter() - t0) / 5
        ts = [threading.Thread(target=one, args=(s, 20)) for s in range(slots)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        tt = (time.perf_counter() - t0) / 20
        print(f"n={n}: one eval {t1 * 1e3:.3f} ms; batch of {slots} in one launch {tb * 1e3:.3f} ms; "
              f"{slots} threads x 1 eval {tt * 1e3:.3f} ms per round", flush=True)
        lib.cmoop_gp_lml_destroy(h)


if __name__ == "__main__":
    main()
