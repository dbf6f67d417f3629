"""Host-side logic of the drop-in (no kernels): operators keep the reference's random stream,
infill selection, MOBO encoding, synthetic data.  CPU only."""
import random

import numpy as np
import pytest

from conftest import records
from cmoop_audio_processing_b200 import nsga, surrogate, synth
from oracle import extract as ex


def test_select_infill_points_matches_reference_golden(golden):
    for case in golden("infill")["cases"]:
        idx, _ = surrogate.select_infill_points(records(case["objs"], case["cv"]), case["k"])
        assert idx == case["indices"]


def test_mobo_encoding_matches_reference_golden(golden):
    g = golden("mobo")
    for item in g["encoding"]:
        np.testing.assert_array_equal(surrogate.hparams_to_vector(item["hp"]), item["vec"])
        assert surrogate.vector_to_hparams(np.array(item["vec"])) == item["hp"]
    for case in g["cases"]:
        best = np.array(case["candidates"])[case["argmax"]]
        assert surrogate.vector_to_hparams(best) == case["decoded"]


def test_lambda_schedule(golden):
    g = golden("nsga")["lambda_schedule"]
    assert [nsga.get_lambda(i, g["max_gen"]) for i in range(g["max_gen"])] == g["values"]


@pytest.mark.skipif(not ex.reference_available(), reason="/root/reference not present")
def test_operators_consume_random_stream_like_the_reference():
    ref = ex.extract("ablation_study/sa_nsga_local.py",
                     ["initialize_population", "dominates", "tournament_selection", "crossover", "mutate",
                      "perturb_hparams"])
    for seed in range(5):
        random.seed(seed)
        pop_r = ref["initialize_population"](9)
        a = ref["crossover"](pop_r[0], pop_r[1])
        b = [ref["mutate"](p) for p in pop_r]
        c = [ref["perturb_hparams"](p) for p in pop_r]
        recs = [{"hparams": p, "objs": [random.random() for _ in range(3)], "CV": random.choice([0.0, 0.3])} for p in pop_r]
        d = [ref["tournament_selection"](recs, 3.0) for _ in range(20)]
        tail_r = random.random()
        random.seed(seed)
        pop_o = nsga.initialize_population(9)
        a2 = nsga.crossover(pop_o[0], pop_o[1])
        b2 = [nsga.mutate(p) for p in pop_o]
        c2 = [surrogate.perturb_hparams(p) for p in pop_o]
        recs2 = [{"hparams": p, "objs": [random.random() for _ in range(3)], "CV": random.choice([0.0, 0.3])} for p in pop_o]
        d2 = [nsga.tournament_selection(recs2, 3.0) for _ in range(20)]
        tail_o = random.random()
        assert (pop_r, a, b, c, d, tail_r) == (pop_o, a2, b2, c2, d2, tail_o)


def test_synthetic_clips_are_deterministic_and_bounded():
    w1, l1 = synth.make_clips(24, 12, seed=1234)
    w2, l2 = synth.make_clips(24, 12, seed=1234)
    assert w1.shape == (24, 16000) and w1.dtype == np.float32
    np.testing.assert_array_equal(w1, w2)
    np.testing.assert_array_equal(l1, l2)
    assert np.abs(w1).max() <= 1.0 and sorted(set(l1.tolist())) == list(range(12))
    u = synth.uniform_clips(4)
    assert u.shape == (4, 16000) and u.min() >= -1 and u.max() <= 1


def test_parallel_gp_fit_matches_sklearn():
    """fit_gprs_parallel = scikit-learn's own objective / L-BFGS-B / best-start selection on a different schedule: with the
    same RandomState the log-marginal likelihood and the predictions of GaussianProcessRegressor.fit are reproduced, both
    in-process (small n) and through the worker subprocesses (n >= 64)."""
    import warnings

    import numpy as np
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    from cmoop_audio_processing_b200.gp_fit import fit_gprs_parallel

    warnings.filterwarnings("ignore")
    rng = np.random.default_rng(3)

    def kernel():
        return ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1)

    for n, workers in ((24, None), (72, 2)):
        x = np.column_stack([rng.choice([16, 32, 64], n), rng.choice([3, 5], n), rng.choice([1, 2, 3], n),
                             rng.choice([1, 2, 3, 4], n), rng.integers(0, 2, n), rng.integers(0, 2, n)]).astype(float)
        ys = [np.sin(x[:, 0] / 20.0 + j) + 0.1 * rng.standard_normal(n) for j in range(2)]
        state = np.random.RandomState(11)
        ref = [GaussianProcessRegressor(kernel=kernel(), n_restarts_optimizer=3, random_state=state).fit(x, y) for y in ys]
        got = fit_gprs_parallel([kernel() for _ in ys], x, ys, n_restarts_optimizer=3, random_state=np.random.RandomState(11),
                                max_workers=workers)
        xq = rng.random((40, 6)) * np.array([64, 5, 3, 4, 1, 1])
        for a, b in zip(ref, got):
            assert abs(a.log_marginal_likelihood_value_ - b.log_marginal_likelihood_value_) < 1e-8
            ma, sa = a.predict(xq, return_std=True)
            mb, sb = b.predict(xq, return_std=True)
            np.testing.assert_allclose(mb, ma, rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(sb, sa, rtol=1e-6, atol=1e-7)
    # the draws come from the same stream in the same order: the next number of both generators agrees
    s1, s2 = np.random.RandomState(5), np.random.RandomState(5)
    GaussianProcessRegressor(kernel=kernel(), n_restarts_optimizer=2, random_state=s1).fit(x[:24], ys[0][:24])
    fit_gprs_parallel([kernel()], x[:24], [ys[0][:24]], n_restarts_optimizer=2, random_state=s2)
    assert s1.uniform() == s2.uniform()


def test_run_log_formats_round_trip(tmp_path):
    """persistence.py: the reference's per-generation / Pareto table columns (nsga_penalty.py:700-763,785-820) and the
    seeded initial population of psi_sa_nsga_local.py:255-269 (CV recomputed from the thresholds)."""
    import random

    from cmoop_audio_processing_b200 import persistence
    from cmoop_audio_processing_b200.nsga import HPARAM_SPACE

    rnd = random.Random(4)
    pops = []
    for gen in range(3):
        pop = []
        for _ in range(5):
            hp = {k: rnd.choice(v) for k, v in HPARAM_SPACE.items()}
            acc, size, fpr = rnd.uniform(0.7, 0.99), rnd.uniform(0.05, 4.0), rnd.uniform(0.0, 0.2)
            cv = max(0, 0.9 - acc) + max(0, size - 2.5) + max(0, fpr - 0.09)
            pop.append({"hparams": hp, "objs": [-acc, size, fpr], "CV": cv})
        pops.append(pop)
    frames = [persistence.generation_frame(g, p) for g, p in enumerate(pops)]
    assert list(frames[0].columns) == ["Generation", "Accuracy", "Size_MB", "FPR", "CV"] + list(HPARAM_SPACE)
    written = persistence.save_generations(frames, str(tmp_path / "all_generations.xlsx"))
    back = persistence.load_generations(written)
    assert len(back) == 3
    for a, b in zip(frames, back):
        assert (a["Accuracy"].to_numpy() - b["Accuracy"].to_numpy()).__abs__().max() < 1e-12
        assert a["filters"].tolist() == b["filters"].tolist() and a["use_bn"].tolist() == [bool(v) for v in b["use_bn"]]
    path = persistence.save_pareto_csv(pops[2], str(tmp_path / "final_pareto.csv"))
    seeded = persistence.load_seed_population(path, 0.9, 2.5, 0.09)
    for got, want in zip(seeded, pops[2]):
        assert got["hparams"] == want["hparams"]
        assert all(abs(x - y) < 1e-12 for x, y in zip(got["objs"], want["objs"]))
        assert abs(got["CV"] - want["CV"]) < 1e-12


def test_device_kernel_spec_recognises_the_reference_kernels_only():
    """gp_fit.device_kernel_spec: the two kernel forms the reference fits (sa_nsga_local.py:180, mobo_penalty.py:259) have
    a device objective; anything else keeps scikit-learn's objective on the host."""
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

    from cmoop_audio_processing_b200.gp_fit import device_kernel_spec

    assert device_kernel_spec(ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(0.1)) == (0, 1.5)
    assert device_kernel_spec(Matern(nu=2.5)) == (1, 2.5)
    assert device_kernel_spec(Matern(nu=0.5)) == (1, 0.5)
    assert device_kernel_spec(RBF()) is None
    assert device_kernel_spec(Matern(nu=3.5)) is None
    assert device_kernel_spec(Matern(length_scale=[1.0, 2.0], nu=1.5)) is None                      # anisotropic
    assert device_kernel_spec(Matern(nu=1.5, length_scale_bounds="fixed")) is None
    assert device_kernel_spec(WhiteKernel(0.1) + ConstantKernel(1.0) * Matern(nu=1.5)) is None     # other theta order
    assert device_kernel_spec(ConstantKernel(1.0) * Matern(nu=1.5)) is None


def test_lock_step_objective_batches_live_chains():
    """gp_fit._LockStepObjective with a stand-in library: every round is one call carrying one request per live chain,
    chains that finish early retire without stalling the others, and each chain sees its own values."""
    import ctypes as C
    import threading

    from scipy.optimize import minimize

    from cmoop_audio_processing_b200.gp_fit import _LockStepObjective

    calls = []

    class FakeLib:
        def cmoop_gp_lml_eval(self, handle, slot, count, theta, target, lml, grad):
            th = np.ctypeslib.as_array(C.cast(theta, C.POINTER(C.c_double)), (count, 2))
            tg = np.ctypeslib.as_array(C.cast(target, C.POINTER(C.c_int32)), (count,))
            out = np.ctypeslib.as_array(C.cast(lml, C.POINTER(C.c_double)), (count,))
            g = np.ctypeslib.as_array(C.cast(grad, C.POINTER(C.c_double)), (count, 2))
            centre = tg[:, None].astype(np.float64)                  # chain with target t minimises |theta - t|^2
            out[:] = -((th - centre) ** 2).sum(axis=1)
            g[:] = -2.0 * (th - centre)
            calls.append((slot, count))
            return 0

    n_chains = 5
    obj = _LockStepObjective(FakeLib(), None, list(range(n_chains)), 2)
    found = {}

    def run(s):
        try:
            def f(th):
                lml, grad = obj.evaluate(s, th)
                return -lml, -grad
            found[s] = minimize(f, np.array([3.0, -2.0]), jac=True, method="L-BFGS-B", options={"maxiter": 2 + 3 * s}).x
        finally:
            obj.retire()

    threads = [threading.Thread(target=run, args=(s,)) for s in range(n_chains)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(30)
    assert not any(t.is_alive() for t in threads)
    assert calls[0] == (0, n_chains) and all(slot == 0 for slot, _ in calls)
    assert [c for _, c in calls] == sorted((c for _, c in calls), reverse=True)      # chains only ever retire
    assert obj.rounds == len(calls) and obj.requests == sum(c for _, c in calls)
    for s in range(1, n_chains):                                     # chain 0 is capped at two iterations
        np.testing.assert_allclose(found[s], [s, s], atol=1e-6)


def test_multiplexed_lbfgsb_is_bit_identical_to_scipy():
    """gp_fit.minimize_lbfgsb_multiplexed drives SciPy's own L-BFGS-B step for many problems in one thread (batching their
    objective requests); for every problem the optimum, value, iteration / evaluation counts and termination message must
    equal scipy.optimize.minimize(method="L-BFGS-B", jac=True, bounds=...) -- scikit-learn's _constrained_optimization."""
    import warnings

    from scipy.optimize import minimize
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    from cmoop_audio_processing_b200 import gp_fit

    if not gp_fit._multiplexed_lbfgsb_available():
        pytest.skip("this SciPy does not expose the reverse-communication L-BFGS-B step")
    rng = np.random.default_rng(0)
    x = np.unique(rng.integers(0, 4, (60, 8)).astype(np.float64), axis=0)
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(len(x))
    y = (y - y.mean()) / y.std()
    kernel = ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(0.1)
    gpr = GaussianProcessRegressor(kernel=kernel, optimizer=None).fit(x, y)
    bounds = gpr.kernel_.bounds
    starts = [gpr.kernel_.theta] + [rng.uniform(bounds[:, 0], bounds[:, 1]) for _ in range(10)]
    starts.append(bounds[:, 1] + 1.0)                                  # outside the box: clipped like SciPy does

    def objective(theta):
        lml, grad = gpr.log_marginal_likelihood(theta, eval_gradient=True, clone_kernel=False)
        return -lml, -grad

    batches = []

    def evaluate_batch(requests):
        batches.append(len(requests))
        return [objective(theta) for _, theta in requests]

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = [minimize(objective, s0, method="L-BFGS-B", jac=True, bounds=bounds) for s0 in starts]
        got = gp_fit.minimize_lbfgsb_multiplexed(evaluate_batch, starts, [bounds] * len(starts))
    assert batches[0] == len(starts) and batches == sorted(batches, reverse=True)
    for w, g in zip(want, got):
        assert np.array_equal(w.x, g.x) and w.fun == g.fun and np.array_equal(w.jac, g.jac)
        assert (w.nit, w.nfev, w.status, w.message, w.success) == (g.nit, g.nfev, g.status, g.message, g.success)
    assert sum(batches) == sum(w.nfev for w in want)
