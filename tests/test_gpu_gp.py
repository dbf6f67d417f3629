"""GPU parity: GP posterior mean/std (1e-5, north_star) vs scikit-learn and the reference-made goldens;
SurrogateManager / local search / MOBO drop-ins end to end."""
import random

import numpy as np
import pytest

from oracle import gp_ref

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _specs_from_golden(models):
    out = []
    for m in models:
        out.append(dict(x_train=np.array(m["x_train"]), alpha=np.array(m["alpha"]), chol_lower=np.array(m["chol_lower"]),
                        amplitude=m["amplitude"], length_scale=m["length_scale"], nu=m["nu"], noise=m["noise"],
                        y_scale=m.get("scaler_scale", m.get("y_scale", 1.0)),
                        y_shift=m.get("scaler_mean", m.get("y_shift", 0.0))))
    return out


def test_posterior_vs_reference_golden_surrogate(golden):
    from cmoop_audio_processing_b200.surrogate import DeviceGPGroup
    for case in golden("surrogate")["cases"]:
        cols = case["feature_columns"]
        cats = {k: sorted({c.split("_")[-1] for c in cols if c.startswith(f"cat__{k}_")}) for k in ("use_bn", "use_dropout")}
        xq = np.array([[float(q[k]) for k in ("filters", "kernel_size", "residual_blocks", "fc_layers")]
                       + [1.0 if str(q[k]) == c else 0.0 for k in ("use_bn", "use_dropout") for c in cats[k]]
                       for q in case["queries"]])
        keys = list(case["models"].keys())
        grp = DeviceGPGroup(_specs_from_golden([case["models"][k] for k in keys]))
        mean, std = grp.predict(xq, return_std=True)
        for j, k in enumerate(keys):
            np.testing.assert_allclose(mean[j], case["pred"][k], rtol=TOL, atol=TOL)
            if case["models"][k]["scaler_var"] > 0:
                np.testing.assert_allclose(std[j], case["std"][k], rtol=TOL, atol=TOL)
        grp.close()


def test_posterior_vs_reference_golden_mobo(golden):
    from cmoop_audio_processing_b200.surrogate import DeviceGPGroup
    for case in golden("mobo")["cases"]:
        grp = DeviceGPGroup(_specs_from_golden(case["models"]))
        mean, _ = grp.predict(np.array(case["candidates"]), return_std=False)
        np.testing.assert_allclose(mean.T, np.array(case["mu"]), rtol=TOL, atol=TOL)
        acq = -np.sum(mean.T[:, :3] + case["lam"] * mean.T[:, 3:4], axis=1)
        assert int(np.argmax(acq)) == case["argmax"]
        grp.close()


@pytest.mark.parametrize("n,nu", [(5, 1.5), (33, 2.5), (128, 1.5), (288, 1.5), (300, 0.5), (700, 2.5)])
def test_posterior_vs_sklearn_live(n, nu):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
    from cmoop_audio_processing_b200.surrogate import DeviceGPGroup
    rng = np.random.default_rng(n)
    x = rng.random((n, 8)) * np.array([64, 5, 3, 4, 1, 1, 1, 1])
    y = np.sin(x[:, 0] / 20) + 0.1 * x[:, 2] + 0.05 * rng.standard_normal(n)
    kernel = ConstantKernel(1.3) * Matern(length_scale=7.0, nu=nu) + WhiteKernel(0.05)
    gpr = GaussianProcessRegressor(kernel=kernel, optimizer=None, normalize_y=True).fit(x, y)
    xq = rng.random((77, 8)) * np.array([64, 5, 3, 4, 1, 1, 1, 1])
    xq[:5] = x[:5]                                           # exact training points: variance ~ noise level
    mu_ref, sd_ref = gpr.predict(xq, return_std=True)
    grp = DeviceGPGroup.from_sklearn([gpr])
    mu, sd = grp.predict(xq, return_std=True)
    np.testing.assert_allclose(mu[0], mu_ref, rtol=TOL, atol=TOL)
    np.testing.assert_allclose(sd[0], sd_ref, rtol=TOL, atol=TOL)
    # and the numpy oracle agrees with both
    u = gp_ref.unpack_sklearn(gpr)
    mu_o, sd_o = gp_ref.posterior(xq, u["x_train"], u["alpha"], u["chol_lower"], amplitude=u["amplitude"],
                                  length_scale=u["length_scale"], nu=u["nu"], noise=u["noise"], y_scale=u["y_scale"],
                                  y_shift=u["y_shift"])
    np.testing.assert_allclose(mu[0], mu_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sd[0], sd_o, rtol=1e-7, atol=1e-9)
    grp.close()


def test_surrogate_manager_matches_reference_end_to_end(golden):
    """Same seeds as oracle/make_golden.py: update (sklearn fit, identical RNG stream) then predict /
    predict_and_structure / perform_local_search must reproduce what the reference's own classes returned."""
    import sklearn
    from cmoop_audio_processing_b200.surrogate import SurrogateManager, perform_local_search
    g = golden("surrogate")
    if sklearn.__version__ != g["sklearn"]:
        pytest.skip("fixtures were fitted with a different scikit-learn")
    for case in g["cases"]:
        np.random.seed(case["seed"])
        random.seed(case["seed"])
        recs = [{"hparams": hp, "objs": o, "CV": c} for hp, o, c in zip(case["train_hparams"], case["train_objs"], case["train_cv"])]
        sm = SurrogateManager()
        with pytest.raises(RuntimeError):
            sm.predict(case["queries"])
        sm.update(case["train_hparams"], recs)
        assert sm.is_fitted and len(sm.training_data) == len(case["dedup_rows"])
        preds, stds = sm.predict(case["queries"], return_std=True)
        for k in preds:
            np.testing.assert_allclose(preds[k], case["pred"][k], rtol=TOL, atol=TOL)
            np.testing.assert_allclose(stds[k], case["std"][k], rtol=TOL, atol=TOL)
        structured = sm.predict_and_structure(case["queries"])
        np.testing.assert_allclose([r["CV"] for r in structured], case["structured_cv"], rtol=TOL, atol=TOL)
        # table-served and direct predictions are the same numbers
        direct = SurrogateManager(use_table=False)
        direct.__dict__.update({k: v for k, v in sm.__dict__.items() if k not in ("_use_table", "_table")})
        p2, s2 = direct.predict(case["queries"], return_std=True)
        for k in preds:
            np.testing.assert_array_equal(preds[k], p2[k])
            np.testing.assert_array_equal(stds[k], s2[k])
        off = [{"hparams": dict(q), "objs": [preds["neg_acc"][i], preds["size"][i], preds["fpr"][i]],
                "stds": [stds["neg_acc"][i], stds["size"][i], stds["fpr"][i]], "CV": max(0, preds["cv"][i])}
               for i, q in enumerate(case["queries"])]
        random.seed(case["ls_seed"])
        improved = perform_local_search(off, sm)
        assert improved == case["ls_improved"]
        assert random.random() == case["ls_rand_after"]


def test_mobo_helpers_end_to_end(golden):
    import sklearn
    from cmoop_audio_processing_b200 import surrogate
    g = golden("mobo")
    if sklearn.__version__ != g["sklearn"]:
        pytest.skip("fixtures were fitted with a different scikit-learn")
    for case in g["cases"]:
        x, y, cv = np.array(case["x"]), np.array(case["y"]), np.array(case["cv"])
        gp_objs = surrogate.train_gps(x, y)
        gp_cv = surrogate.train_gps(x, cv[:, None])[0]
        cand = np.array(case["candidates"])
        mu = surrogate.predict_gps(gp_objs + [gp_cv], cand)
        np.testing.assert_allclose(mu, np.array(case["mu"]), rtol=TOL, atol=TOL)
        acq = surrogate.penalized_acquisition(cand, gp_objs, gp_cv, case["lam"])
        np.testing.assert_allclose(acq, case["acq"], rtol=TOL, atol=TOL)
        assert surrogate.vector_to_hparams(cand[int(np.argmax(acq))]) == case["decoded"]
