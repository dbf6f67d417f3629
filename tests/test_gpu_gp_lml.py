"""Device log-marginal likelihood + gradient (csrc/gp_lml.cu) against scikit-learn's own
``GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)`` (the objective of the hyper-parameter fit in
SurrogateManager.update, sa_nsga_local.py:195-210, and train_gps, mobo_penalty.py:252-263), and the device-backed fit
against the host fit."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _genotype_rows(n, rng):
    """De-duplicated 8-column surrogate features as SurrogateManager builds them (sa_nsga_local.py:183-193)."""
    space = [(f, k, r, fc, bn, 1 - bn, dr, 1 - dr) for f in (16, 32, 64, 128) for k in (3, 5) for r in (1, 2, 3)
             for fc in (1, 2, 3) for bn in (0, 1) for dr in (0, 1)]
    idx = rng.permutation(len(space))[:n]
    return np.asarray([space[i] for i in idx], np.float64)


def _targets(x, rng):
    f = x[:, 0] / 128.0
    y0 = -0.9 + 0.2 * np.exp(-f) + 0.02 * rng.standard_normal(len(x))
    y1 = 0.1 * x[:, 0] * x[:, 1] / 50.0 + 0.3 * x[:, 2]
    return [y0, y1]


def _device_lml(x, ys, kind, nu, thetas, targets, alpha=1e-10):
    from cmoop_audio_processing_b200 import _lib
    lib = _lib.load()
    x = np.ascontiguousarray(x, np.float64)
    ys = np.ascontiguousarray(ys, np.float64)
    thetas = np.ascontiguousarray(thetas, np.float64)
    targets = np.ascontiguousarray(targets, np.int32)
    b = len(targets)
    h = C.c_void_p()
    _lib.check(lib.cmoop_gp_lml_create(_lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(ys), ys.shape[0], kind, nu, alpha, b,
                                       C.byref(h)), "create")
    try:
        nt = lib.cmoop_gp_lml_n_theta(h)
        lml, grad = np.empty(b), np.empty((b, nt))
        _lib.check(lib.cmoop_gp_lml_eval(h, 0, b, _lib.ptr(thetas), _lib.ptr(targets), _lib.ptr(lml), _lib.ptr(grad)), "eval")
    finally:
        lib.cmoop_gp_lml_destroy(h)
    return lml, grad


@pytest.mark.parametrize("n", [1, 15, 97, 288])
def test_lml_and_gradient_match_sklearn_surrogate_kernel(n):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    rng = np.random.default_rng(n)
    x = _genotype_rows(n, rng)
    ys = _targets(x, rng)
    ys = [(y - y.mean()) / (y.std() if y.std() > 0 else 1.0) for y in ys]       # SurrogateManager z-scores its targets
    kernel = ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1)
    thetas = np.concatenate([[kernel.theta], rng.uniform([-4, -2, -9], [6, 6, 2], size=(9, 3))])
    for m, y in enumerate(ys):
        gpr = GaussianProcessRegressor(kernel=kernel, optimizer=None).fit(x, y)
        want = [gpr.log_marginal_likelihood(t, eval_gradient=True) for t in thetas]
        lml, grad = _device_lml(x, np.stack(ys), 0, 1.5, thetas, np.full(len(thetas), m))
        for i, (wl, wg) in enumerate(want):
            assert lml[i] == pytest.approx(wl, rel=1e-8, abs=1e-8), (n, m, i)
            np.testing.assert_allclose(grad[i], wg, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(wg).max()))


@pytest.mark.parametrize("nu", [0.5, 1.5, 2.5])
def test_lml_matches_sklearn_mobo_kernel(nu):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Matern

    rng = np.random.default_rng(7)
    x = rng.uniform(size=(45, 6))                                  # hparams_to_vector: [0, 1]^6 (mobo_penalty.py:305-318)
    y = np.sin(3 * x[:, 0]) + x[:, 1] * x[:, 2]
    gpr = GaussianProcessRegressor(kernel=Matern(nu=nu), optimizer=None, normalize_y=True).fit(x, y)
    thetas = np.linspace(-3.0, 3.0, 7)[:, None]
    want = [gpr.log_marginal_likelihood(t, eval_gradient=True) for t in thetas]
    lml, grad = _device_lml(x, gpr.y_train_[None], 1, nu, thetas, np.zeros(len(thetas)))
    for i, (wl, wg) in enumerate(want):
        assert lml[i] == pytest.approx(wl, rel=1e-8, abs=1e-8)
        np.testing.assert_allclose(grad[i], wg, rtol=1e-6, atol=1e-8)


def test_not_positive_definite_reports_minus_infinity():
    """Duplicate rows with (almost) no noise: scikit-learn returns (-inf, 0) when the Cholesky fails (_gpr.py)."""
    x = np.zeros((6, 3))
    y = np.arange(6, dtype=np.float64)[None]
    lml, grad = _device_lml(x, y, 1, 1.5, np.zeros((1, 1)), np.zeros(1), alpha=0.0)
    assert np.isneginf(lml[0]) and np.all(grad == 0.0)


def test_bad_arguments_are_reported():
    from cmoop_audio_processing_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    x = np.zeros((4, 2))
    y = np.zeros((1, 4))
    assert lib.cmoop_gp_lml_create(_lib.ptr(x), 4, 2, _lib.ptr(y), 1, 7, 1.5, 1e-10, 1, C.byref(h)) != 0
    assert lib.cmoop_gp_lml_create(_lib.ptr(x), 4, 2, _lib.ptr(y), 1, 0, 2.0, 1e-10, 1, C.byref(h)) != 0
    assert lib.cmoop_gp_lml_create(_lib.ptr(x), 2000, 2, _lib.ptr(y), 1, 0, 1.5, 1e-10, 1, C.byref(h)) != 0
    _lib.check(lib.cmoop_gp_lml_create(_lib.ptr(x), 4, 2, _lib.ptr(y), 1, 0, 1.5, 1e-10, 2, C.byref(h)), "create")
    out = np.zeros(4)
    th = np.zeros(3)
    tg = np.array([3], np.int32)
    assert lib.cmoop_gp_lml_eval(h, 0, 1, _lib.ptr(th), _lib.ptr(tg), _lib.ptr(out), _lib.ptr(out)) != 0      # target
    assert lib.cmoop_gp_lml_eval(h, 2, 1, _lib.ptr(th), _lib.ptr(tg), _lib.ptr(out), _lib.ptr(out)) != 0      # slot
    lib.cmoop_gp_lml_destroy(h)


def test_device_backed_fit_agrees_with_host_fit():
    """Same starts, same optimiser, device objective: fitted hyper-parameters and predictions agree with the host fit."""
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    from cmoop_audio_processing_b200 import gp_fit
    from cmoop_audio_processing_b200.gp_fit import fit_gprs_parallel

    rng = np.random.default_rng(3)
    x = _genotype_rows(120, rng)
    ys = _targets(x, rng)
    ys = [(y - y.mean()) / y.std() for y in ys]
    kernels = [ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1) for _ in ys]
    host = fit_gprs_parallel(kernels, x, ys, n_restarts_optimizer=4, random_state=11, max_workers=1, backend="host")
    dev = fit_gprs_parallel(kernels, x, ys, n_restarts_optimizer=4, random_state=11, backend="device")
    assert gp_fit.LAST_DEVICE_FIT["chains"] == 10 and gp_fit.LAST_DEVICE_FIT["rounds"] > 0
    xq = _genotype_rows(60, np.random.default_rng(5))
    for a, b in zip(host, dev):
        assert b.log_marginal_likelihood_value_ == pytest.approx(a.log_marginal_likelihood_value_, rel=1e-6, abs=1e-6)
        ma, sa = a.predict(xq, return_std=True)
        mb, sb = b.predict(xq, return_std=True)
        np.testing.assert_allclose(mb, ma, rtol=0, atol=1e-4)
        np.testing.assert_allclose(sb, sa, rtol=0, atol=1e-4)
    # the MOBO form (bare Matern, normalize_y) through the same entry point
    xm = rng.uniform(size=(40, 6))
    ym = [np.sin(3 * xm[:, 0]) + xm[:, 1], xm[:, 2] ** 2]
    host = fit_gprs_parallel([Matern(nu=2.5)] * 2, xm, ym, normalize_y=True, max_workers=1, backend="host")
    dev = fit_gprs_parallel([Matern(nu=2.5)] * 2, xm, ym, normalize_y=True, backend="device")
    for a, b in zip(host, dev):
        np.testing.assert_allclose(b.kernel_.theta, a.kernel_.theta, rtol=0, atol=1e-5)


def test_device_fitted_surrogate_reproduces_reference_predictions(golden):
    """SurrogateManager(fit_backend="device") on the golden cases the reference's own class produced (oracle/make_golden.py):
    same RNG stream for the optimiser starts, objective on the GPU; predictions within 1e-4 of the reference's (the host
    backend, which runs scikit-learn's objective itself, is held to 1e-5 in test_gpu_gp.py)."""
    import random

    import sklearn

    from cmoop_audio_processing_b200 import gp_fit
    from cmoop_audio_processing_b200.surrogate import SurrogateManager

    g = golden("surrogate")
    if sklearn.__version__ != g["sklearn"]:
        pytest.skip("fixtures were fitted with a different scikit-learn")
    for case in g["cases"]:
        np.random.seed(case["seed"])
        random.seed(case["seed"])
        recs = [{"hparams": hp, "objs": o, "CV": c}
                for hp, o, c in zip(case["train_hparams"], case["train_objs"], case["train_cv"])]
        sm = SurrogateManager(fit_backend="device")
        gp_fit.LAST_DEVICE_FIT.clear()
        sm.update(case["train_hparams"], recs)
        assert gp_fit.LAST_DEVICE_FIT.get("rounds", 0) > 0 and gp_fit.LAST_DEVICE_FIT["chains"] == 44      # no host fallback
        preds, stds = sm.predict(case["queries"], return_std=True)
        for k in preds:
            np.testing.assert_allclose(preds[k], case["pred"][k], rtol=0, atol=1e-4)
            np.testing.assert_allclose(stds[k], case["std"][k], rtol=0, atol=1e-4)


@pytest.mark.parametrize("n", [64, 65, 160, 161, 300, 512])
def test_lml_kernel_instantiation_boundaries(n):
    """Sizes either side of the rows-per-thread buckets of gp_lml_kernel<MR> (64 / 160 / 288) and the largest n."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    rng = np.random.default_rng(1000 + n)
    x = rng.uniform(0.0, 4.0, size=(n, 5))
    y = np.sin(x[:, 0]) + 0.3 * x[:, 1] + 0.05 * rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    kernel = ConstantKernel(1.0) * Matern(length_scale=1.0, nu=2.5) + WhiteKernel(noise_level=0.1)
    gpr = GaussianProcessRegressor(kernel=kernel, optimizer=None).fit(x, y)
    thetas = np.array([[0.0, 0.0, np.log(0.1)], [1.5, 0.7, -4.0], [-1.0, -0.5, -1.0]])
    lml, grad = _device_lml(x, y[None], 0, 2.5, thetas, np.zeros(3))
    for i, t in enumerate(thetas):
        wl, wg = gpr.log_marginal_likelihood(t, eval_gradient=True)
        assert lml[i] == pytest.approx(wl, rel=1e-8, abs=1e-8)
        np.testing.assert_allclose(grad[i], wg, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(wg).max()))
