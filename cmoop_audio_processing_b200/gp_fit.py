"""Concurrent multi-start fit of scikit-learn Gaussian-process regressors (host side of the surrogate).

scikit-learn's ``GaussianProcessRegressor.fit`` (sklearn:_gpr.py:299-340) maximises the log-marginal likelihood from the
kernel's initial theta and from ``n_restarts_optimizer`` log-uniform draws, one start after the other.  In
``SurrogateManager.update`` (sa_nsga_local.py:180-181,209: 4 models x 11 starts, every generation) that costs 10-25 s on
the host once the training set approaches the 288 genotypes of the space -- more than the GPU needs for the true
evaluations of the generation.  The starts are independent, so this module keeps the arithmetic (scikit-learn's own
``log_marginal_likelihood``, ``_constrained_optimization`` = SciPy L-BFGS-B, selection of the best start, final Cholesky)
and only changes the schedule: the initial points are drawn first, in scikit-learn's order and from the same RandomState
(the global NumPy stream for ``random_state=None``, as in the reference), then every (model, start) pair is optimised by a
pool of worker PROCESSES, each with BLAS pinned to one thread (threads do not help: the objective is GIL-bound
Python/NumPy glue around small LAPACK calls).  The workers are plain ``python -m cmoop_audio_processing_b200.gp_fit``
subprocesses fed pickles over pipes -- not ``multiprocessing`` children -- so nothing of the host program is forked or
re-imported (the reference's drivers run at import time, and the parent owns a CUDA context).  With a fixed
``random_state`` the fitted models equal those of ``GaussianProcessRegressor.fit`` (same log-marginal likelihood and
predictions; tests/test_host_logic.py).  Small problems stay in-process.
"""
from __future__ import annotations

import atexit
import os
import pickle
import struct
import subprocess
import sys
import threading
from operator import itemgetter

import numpy as np


def _optimise_start(payload):
    """One L-BFGS-B run of scikit-learn's objective from theta0; returns (theta_opt, -lml)."""
    from sklearn.gaussian_process import GaussianProcessRegressor

    kernel, x, y, normalize_y, theta0 = payload
    # a private regressor per start: log_marginal_likelihood(clone_kernel=False) mutates kernel_.theta
    w = GaussianProcessRegressor(kernel=kernel, optimizer=None, normalize_y=normalize_y).fit(x, y)
    w.optimizer = "fmin_l_bfgs_b"

    def obj_func(theta, eval_gradient=True):
        if eval_gradient:
            lml, grad = w.log_marginal_likelihood(theta, eval_gradient=True, clone_kernel=False)
            return -lml, -grad
        return -w.log_marginal_likelihood(theta, clone_kernel=False)

    theta_opt, fval = w._constrained_optimization(obj_func, theta0, w.kernel_.bounds)
    return np.asarray(theta_opt, np.float64), float(fval)


# ---- worker side: length-prefixed pickles on stdin / stdout -------------------------------------------------------
def _serve() -> None:
    import warnings

    warnings.filterwarnings("ignore")
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:                                   # pragma: no cover
        pass
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr                             # stray prints must not corrupt the result stream
    while True:
        head = inp.read(8)
        if len(head) < 8:
            return
        payload = pickle.loads(inp.read(struct.unpack("<Q", head)[0]))
        try:
            result = ("ok", _optimise_start(payload))
        except Exception as exc:                        # reported to the parent, which falls back in-process
            result = ("err", repr(exc))
        blob = pickle.dumps(result, protocol=pickle.HIGHEST_PROTOCOL)
        out.write(struct.pack("<Q", len(blob)))
        out.write(blob)
        out.flush()


# ---- parent side ------------------------------------------------------------------------------------------------------
class _Worker:
    def __init__(self):
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        self.proc = subprocess.Popen([sys.executable, "-m", "cmoop_audio_processing_b200.gp_fit"], stdin=subprocess.PIPE,
                                     stdout=subprocess.PIPE, env=env)

    def call(self, payload):
        blob = pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL)
        self.proc.stdin.write(struct.pack("<Q", len(blob)))
        self.proc.stdin.write(blob)
        self.proc.stdin.flush()
        head = self.proc.stdout.read(8)
        if len(head) < 8:
            raise RuntimeError("gp_fit worker died")
        status, value = pickle.loads(self.proc.stdout.read(struct.unpack("<Q", head)[0]))
        if status != "ok":
            raise RuntimeError(f"gp_fit worker: {value}")
        return value

    def close(self):
        try:
            self.proc.stdin.close()
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


_WORKERS: list[_Worker] = []


def _shutdown():
    while _WORKERS:
        _WORKERS.pop().close()


atexit.register(_shutdown)


def _map_on_workers(payloads, n_workers):
    while len(_WORKERS) < n_workers:
        _WORKERS.append(_Worker())
    results = [None] * len(payloads)
    errors = []
    lock = threading.Lock()
    nxt = [0]

    def drive(worker):
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(payloads):
                return
            try:
                results[i] = worker.call(payloads[i])
            except Exception as exc:
                errors.append(exc)
                return

    threads = [threading.Thread(target=drive, args=(w,), daemon=True) for w in _WORKERS[:n_workers]]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        _shutdown()                                     # a broken worker must not be reused
        raise errors[0]
    return results


def _initialised_dist():
    try:
        import torch.distributed as dist
    except ImportError:                                 # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def default_workers(shared: bool = True) -> int:
    """Worker processes for the optimiser starts: all host cores when this process is the only one fitting (rank 0 of a
    distributed job broadcasts its optima), otherwise the cores divided by the local world size."""
    env = os.environ.get("CMOOP_GP_FIT_WORKERS")
    if env:
        return max(1, int(env))
    world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))) if shared else 1
    return max(1, min(32, (os.cpu_count() or 1) // world))


def fit_gprs_parallel(kernels, x, ys, n_restarts_optimizer=0, normalize_y=False, random_state=None, max_workers=None,
                      min_rows_for_pool=64):
    """Fitted ``GaussianProcessRegressor`` per (kernel, y) pair; see the module docstring."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.utils import check_random_state

    rng = check_random_state(random_state)
    x = np.asarray(x, np.float64)
    jobs, probes = [], []
    for kernel, y in zip(kernels, ys):
        probe = GaussianProcessRegressor(kernel=kernel, optimizer=None, normalize_y=normalize_y).fit(x, y)
        probes.append(probe)
        bounds = probe.kernel_.bounds
        starts = [probe.kernel_.theta.copy()]
        if n_restarts_optimizer > 0:
            if not np.isfinite(bounds).all():
                raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all bounds are finite.")
            starts += [rng.uniform(bounds[:, 0], bounds[:, 1]) for _ in range(n_restarts_optimizer)]
        jobs += [(len(probes) - 1, theta0) for theta0 in starts]
    payloads = [(kernels[m], x, ys[m], normalize_y, theta0) for m, theta0 in jobs]

    # Under torch.distributed the surrogate is replicated on every rank.  Every rank has drawn the
    # same starts above (identically seeded streams stay in step), but only rank 0 optimises -- with all host cores -- and
    # broadcasts the optima, instead of world_size ranks fitting the same models on a slice of the cores each.
    dist = _initialised_dist()
    rank0 = dist is None or dist.get_rank() == 0
    results = None
    if rank0:
        workers = max_workers if max_workers is not None else default_workers(shared=dist is None)
        workers = min(workers, len(jobs))
        if workers > 1 and len(x) >= min_rows_for_pool:
            try:
                results = _map_on_workers(payloads, workers)
            except Exception:                           # no subprocesses here (sandbox, frozen app): same maths in-process
                results = None
        if results is None:
            results = [_optimise_start(pl) for pl in payloads]
    if dist is not None:
        box = [results]
        dist.broadcast_object_list(box, src=0)
        results = box[0]

    fitted = []
    for m, probe in enumerate(probes):
        optima = [r for (mm, _), r in zip(jobs, results) if mm == m]      # scikit-learn's order: initial theta first
        lml_values = list(map(itemgetter(1), optima))
        best = optima[int(np.argmin(lml_values))][0]
        gpr = GaussianProcessRegressor(kernel=probe.kernel_.clone_with_theta(best), optimizer=None, normalize_y=normalize_y)
        gpr.fit(x, ys[m])
        gpr.kernel_._check_bounds_params()
        gpr.log_marginal_likelihood_value_ = -float(np.min(lml_values))
        gpr.n_restarts_optimizer = n_restarts_optimizer
        fitted.append(gpr)
    return fitted


if __name__ == "__main__":
    _serve()
