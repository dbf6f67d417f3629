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

``backend="device"`` (or ``CMOOP_GP_FIT_BACKEND=device``) keeps SciPy's L-BFGS-B and the start order but evaluates the
objective -- log-marginal likelihood and gradient -- with ``cmoop_gp_lml_eval`` (csrc/gp_lml.cu): the chains advance in
lock step so that each round of objective requests (44 for a surrogate update) is one kernel launch with one CTA per
request.  One host thread steps all chains through SciPy's reverse-communication ``setulb``
(``minimize_lbfgsb_multiplexed``, per chain bit-identical to ``scipy.optimize.minimize``); SciPy builds without that entry
point fall back to one optimiser thread per chain meeting in ``_LockStepObjective``.  The
device objective agrees with scikit-learn's to ~1e-10 relative, not bit for bit, so the optima can differ in the last
digits: it is opt-in, and kernels other than the reference's two forms stay on the host.
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


def device_kernel_spec(kernel):
    """(kind, nu) when gp_lml.cu evaluates this kernel -- ``C * Matern + WhiteKernel`` (kind 0, theta = log c, log l, log
    noise) or a bare ``Matern`` (kind 1), isotropic, nothing fixed -- else None."""
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, Product, Sum, WhiteKernel

    def matern_ok(k):
        return (isinstance(k, Matern) and np.ndim(k.length_scale) == 0 and k.nu in (0.5, 1.5, 2.5)
                and not isinstance(k.length_scale_bounds, str))

    if matern_ok(kernel):
        return 1, float(kernel.nu)
    if (isinstance(kernel, Sum) and isinstance(kernel.k1, Product) and isinstance(kernel.k2, WhiteKernel)
            and isinstance(kernel.k1.k1, ConstantKernel) and matern_ok(kernel.k1.k2)
            and not isinstance(kernel.k1.k1.constant_value_bounds, str)
            and not isinstance(kernel.k2.noise_level_bounds, str)):
        return 0, float(kernel.k1.k2.nu)
    return None


def _multiplexed_lbfgsb_available() -> bool:
    """True when SciPy exposes the reverse-communication L-BFGS-B step this module can drive for many problems at once
    (scipy.optimize._lbfgsb_py._lbfgsb.setulb with the 17-argument form of SciPy >= 1.15)."""
    try:
        from scipy.optimize import _lbfgsb_py as lb

        doc = lb._lbfgsb.setulb.__doc__ or ""
        return doc.lstrip().startswith("setulb(m, x, l, u, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave, "
                                       "maxls, ln_task)") and hasattr(lb, "HAS_ILP64")
    except Exception:
        return False


def minimize_lbfgsb_multiplexed(evaluate_batch, starts, bounds_list):
    """``scipy.optimize.minimize(fun, x0, method="L-BFGS-B", jac=True, bounds=bounds)`` with default options -- what
    scikit-learn's ``_constrained_optimization`` runs (sklearn:_gpr.py) -- for several independent problems at once, in
    one thread: SciPy's own driver loop (scipy/optimize/_lbfgsb_py.py ``_minimize_lbfgsb``) around the same compiled
    ``setulb`` step, with the objective requests of all live problems collected and handed to
    ``evaluate_batch([(problem, x), ...]) -> [(f, g), ...]`` together.  Every problem sees exactly the sequence of points,
    values and decisions it would see alone, so the optima are bit-identical to SciPy's for the same objective.
    Returns [OptimizeResult] in problem order."""
    import inspect

    from scipy.optimize import OptimizeResult
    from scipy.optimize import _lbfgsb_py as lb

    defaults = {k: v.default for k, v in inspect.signature(lb._minimize_lbfgsb).parameters.items()}
    m, maxls = defaults["maxcor"], defaults["maxls"]
    maxiter, maxfun = defaults["maxiter"], defaults["maxfun"]
    factr, pgtol = defaults["ftol"] / np.finfo(float).eps, defaults["gtol"]
    int_dtype = np.int64 if lb.HAS_ILP64 else np.int32
    bounds_map = {(-np.inf, np.inf): 0, (1, np.inf): 1, (1, 1): 2, (-np.inf, 1): 3}
    setulb = lb._lbfgsb.setulb

    class Chain:
        pass

    chains = []
    for x0, bounds in zip(starts, bounds_list):
        c = Chain()
        b = np.array(lb.old_bound_to_new([tuple(r) for r in np.asarray(bounds, np.float64)]))
        x0 = np.clip(np.asarray(x0, np.float64).ravel(), b[0], b[1])
        n = x0.shape[0]
        c.nbd, c.low, c.up = np.zeros(n, int_dtype), np.zeros(n), np.zeros(n)
        for i in range(n):
            lo, hi = b[0, i], b[1, i]
            if not np.isinf(lo):
                c.low[i], lo = lo, 1
            if not np.isinf(hi):
                c.up[i], hi = hi, 1
            c.nbd[i] = bounds_map[lo, hi]
        c.x, c.f, c.g = np.array(x0, np.float64), 0.0, np.zeros(n)
        c.wa, c.iwa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m), np.zeros(3 * n, int_dtype)
        c.task, c.ln_task, c.lsave = np.zeros(2, int_dtype), np.zeros(2, int_dtype), np.zeros(4, int_dtype)
        c.isave, c.dsave = np.zeros(44, int_dtype), np.zeros(29)
        c.nit = c.nfev = 0
        c.last_x = None
        chains.append(c)

    live = list(range(len(chains)))
    while live:
        asking = []
        for ci in live:
            c = chains[ci]
            while True:
                setulb(m, c.x, c.low, c.up, c.nbd, c.f, c.g, factr, pgtol, c.wa, c.iwa, c.task, c.lsave, c.isave, c.dsave,
                       maxls, c.ln_task)
                if c.task[0] == 3:                       # wants f and g at the current x
                    if c.last_x is not None and np.array_equal(c.x, c.last_x):
                        c.f, c.g = c.last_f, c.last_g.copy()      # SciPy's ScalarFunction does not re-evaluate an unchanged x
                        continue
                    asking.append(ci)
                    break
                if c.task[0] == 1:                       # new iteration
                    c.nit += 1
                    if c.nit >= maxiter:
                        c.task[0], c.task[1] = 5, 504
                    elif c.nfev > maxfun:
                        c.task[0], c.task[1] = 5, 502
                    continue
                break                                    # converged / stopped / error
        live = asking
        if asking:
            values = evaluate_batch([(ci, chains[ci].x.copy()) for ci in asking])
            for ci, (f, g) in zip(asking, values):
                c = chains[ci]
                c.f, c.g = float(f), np.asarray(g, np.float64).copy()
                c.last_x, c.last_f, c.last_g = c.x.copy(), c.f, c.g.copy()
                c.nfev += 1
    results = []
    for c in chains:
        warnflag = 0 if c.task[0] == 4 else (1 if c.nfev > maxfun or c.nit >= maxiter else 2)
        results.append(OptimizeResult(fun=c.f, jac=c.g, nfev=c.nfev, njev=c.nfev, nit=c.nit, status=warnflag,
                                      message=lb.status_messages[c.task[0]] + ": " + lb.task_messages[c.task[1]],
                                      x=c.x, success=warnflag == 0))
    return results


class _LockStepObjective:
    """One objective request per live optimiser chain, evaluated together: the chains (threads running SciPy's L-BFGS-B)
    block in ``evaluate`` until every live chain has asked, then ONE ``cmoop_gp_lml_eval`` launch serves them all (one CTA
    per request).  A chain that has converged retires and no longer counts."""

    def __init__(self, lib, handle, targets, n_theta):
        self.lib, self.handle, self.targets, self.n_theta = lib, handle, targets, n_theta
        self.cv = threading.Condition()
        self.live = len(targets)
        self.pending, self.results = {}, {}
        self.error = None
        self.rounds = self.requests = 0

    def _flush(self):
        from . import _lib

        slots = sorted(self.pending)
        self.rounds += 1
        self.requests += len(slots)
        thetas = np.ascontiguousarray([self.pending[s] for s in slots], np.float64)
        target = np.ascontiguousarray([self.targets[s] for s in slots], np.int32)
        lml, grad = np.empty(len(slots)), np.empty((len(slots), self.n_theta))
        try:        # scratch slots are interchangeable: request k of this round uses slot k
            _lib.check(self.lib.cmoop_gp_lml_eval(self.handle, 0, len(slots), _lib.ptr(thetas), _lib.ptr(target),
                                                  _lib.ptr(lml), _lib.ptr(grad)), "cmoop_gp_lml_eval")
        except Exception as exc:
            self.error = exc
        for k, s in enumerate(slots):
            self.results[s] = (float(lml[k]), grad[k].copy())
        self.pending.clear()
        self.cv.notify_all()

    def evaluate(self, slot, theta):
        with self.cv:
            self.pending[slot] = np.array(theta, np.float64)
            if len(self.pending) == self.live:
                self._flush()
            while slot not in self.results:
                self.cv.wait()
            if self.error is not None:
                raise self.error
            return self.results.pop(slot)

    def retire(self):
        with self.cv:
            self.live -= 1
            if self.pending and len(self.pending) == self.live:
                self._flush()


def _optimise_on_device(probes, jobs):
    """Every (model, start) job optimised by SciPy L-BFGS-B (scikit-learn's ``_constrained_optimization``) on the device
    objective; returns [(theta_opt, -lml)] in job order, or None when a kernel has no device form."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor

    from sklearn.gaussian_process import GaussianProcessRegressor

    from . import _lib

    specs = {device_kernel_spec(p.kernel_) for p in probes}
    targets = [np.asarray(p.y_train_, np.float64) for p in probes]
    if len(specs) != 1 or None in specs or any(t.size != t.shape[0] for t in targets):      # one output column per model
        return None
    kind, nu = next(iter(specs))
    LAST_DEVICE_FIT.clear()
    lib = _lib.load()
    _lib.bind_device()
    x = np.ascontiguousarray(probes[0].X_train_, np.float64)
    ys = np.ascontiguousarray(np.stack([t.reshape(-1) for t in targets]), np.float64)
    handle = C.c_void_p()
    _lib.check(lib.cmoop_gp_lml_create(_lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(ys), ys.shape[0], kind, nu,
                                       float(probes[0].alpha), len(jobs), C.byref(handle)), "cmoop_gp_lml_create")
    n_theta = lib.cmoop_gp_lml_n_theta(handle)
    if _multiplexed_lbfgsb_available() and not os.environ.get("CMOOP_GP_FIT_THREADS"):
        # one thread drives every chain through SciPy's own L-BFGS-B step; each round of requests is one launch
        from sklearn.utils.optimize import _check_optimize_result

        stats = {"rounds": 0, "requests": 0}

        def evaluate_batch(requests):
            thetas = np.ascontiguousarray([th for _, th in requests], np.float64)
            target = np.ascontiguousarray([jobs[ci][0] for ci, _ in requests], np.int32)
            lml, grad = np.empty(len(requests)), np.empty((len(requests), n_theta))
            _lib.check(lib.cmoop_gp_lml_eval(handle, 0, len(requests), _lib.ptr(thetas), _lib.ptr(target), _lib.ptr(lml),
                                             _lib.ptr(grad)), "cmoop_gp_lml_eval")
            stats["rounds"] += 1
            stats["requests"] += len(requests)
            return [(-lml[k], -grad[k]) for k in range(len(requests))]

        try:
            optima = minimize_lbfgsb_multiplexed(evaluate_batch, [theta0 for _, theta0 in jobs],
                                                 [probes[m].kernel_.bounds for m, _ in jobs])
        finally:
            lib.cmoop_gp_lml_destroy(handle)
        for res in optima:
            _check_optimize_result("lbfgs", res)                      # scikit-learn's convergence warnings
        LAST_DEVICE_FIT.update(rounds=stats["rounds"], requests=stats["requests"], chains=len(jobs), driver="multiplexed")
        return [(np.asarray(res.x, np.float64), float(res.fun)) for res in optima]

    objective = _LockStepObjective(lib, handle, [m for m, _ in jobs], n_theta)
    driver = GaussianProcessRegressor(optimizer="fmin_l_bfgs_b")      # only its _constrained_optimization is used

    def run(slot):
        m, theta0 = jobs[slot]

        def obj_func(theta, eval_gradient=True):
            lml, grad = objective.evaluate(slot, theta)
            return (-lml, -grad) if eval_gradient else -lml

        try:
            theta_opt, fval = driver._constrained_optimization(obj_func, theta0, probes[m].kernel_.bounds)
        finally:
            objective.retire()
        return np.asarray(theta_opt, np.float64), float(fval)

    try:
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:      # every chain needs its own thread: they meet in evaluate
            return list(pool.map(run, range(len(jobs))))
    finally:
        lib.cmoop_gp_lml_destroy(handle)
        LAST_DEVICE_FIT.update(rounds=objective.rounds, requests=objective.requests, chains=len(jobs), driver="threads")


LAST_DEVICE_FIT: dict = {}       # launches / objective requests of the most recent device-backed fit (diagnostics)


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
                      min_rows_for_pool=64, backend=None):
    """Fitted ``GaussianProcessRegressor`` per (kernel, y) pair; see the module docstring."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.utils import check_random_state

    backend = backend or os.environ.get("CMOOP_GP_FIT_BACKEND", "host")
    if backend not in ("host", "device"):                # validated on EVERY rank, before any rank-dependent branch
        raise ValueError(f"unknown GP fit backend {backend!r} (host | device)")
    rng = check_random_state(random_state)
    x = np.asarray(x, np.float64)
    # Under torch.distributed the surrogate is replicated on every rank.  Every rank draws the
    # same starts (identically seeded streams stay in step), but only rank 0 optimises -- with all host cores -- and
    # broadcasts the optima, instead of world_size ranks fitting the same models on a slice of the cores each.
    # Anything that can raise (non-finite bounds, a CUDA error in the device objective, a dead worker) is caught per
    # rank and exchanged, so that an error on one rank is re-raised on ALL of them instead of leaving the others
    # blocked in the broadcast until the collective times out.
    dist = _initialised_dist()
    rank0 = dist is None or dist.get_rank() == 0
    jobs, probes, results, error = [], [], None, None
    try:
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
        if rank0 and backend == "device":
            results = _optimise_on_device(probes, jobs)      # None: a kernel without a device form -> host schedule below
        if rank0 and results is None:
            workers = max_workers if max_workers is not None else default_workers(shared=dist is None)
            workers = min(workers, len(jobs))
            if workers > 1 and len(x) >= min_rows_for_pool:
                try:
                    results = _map_on_workers(payloads, workers)
                except Exception:                           # no subprocesses here (sandbox, frozen app): same maths in-process
                    results = None
            if results is None:
                results = [_optimise_start(pl) for pl in payloads]
    except Exception as exc:                                # noqa: BLE001 -- exchanged below, then re-raised
        if dist is None:
            raise
        error = exc
    if dist is not None:
        states = [None] * dist.get_world_size()
        dist.all_gather_object(states, None if error is None else f"rank {dist.get_rank()}: {error!r}")
        failed = [s for s in states if s is not None]
        if failed:
            if error is not None:
                raise error
            raise RuntimeError("GP hyper-parameter fit failed on another rank -- " + "; ".join(failed))
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
