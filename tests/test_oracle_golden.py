"""The oracle restatements against golden vectors produced by the REFERENCE'S OWN code
(oracle/make_golden.py, AST-extracted functions).  CPU only."""
import math

import numpy as np
import pytest

from conftest import records
from oracle import gp_ref, nsga_ref


def test_nds_fronts_and_crowding_bit_exact(golden):
    g = golden("nsga")
    assert len(g["cases"]) >= 60
    for case in g["cases"]:
        recs = records(case["objs"], case["cv"])
        fronts = nsga_ref.fast_non_dominated_sort(recs, case["lam"])
        assert fronts == case["fronts"]
        for f, gt, lt in zip(fronts, case["crowd_gt"], case["crowd_lt"]):
            d0 = nsga_ref.crowding_distance(f, recs, skip_on_equal=True)
            d1 = nsga_ref.crowding_distance(f, recs, skip_on_equal=False)
            assert [d0[i] for i in f] == gt          # == on floats: bit-exact (inf == inf)
            assert [d1[i] for i in f] == lt
        assert nsga_ref.environmental_selection(recs, case["lam"], case["pop"]) == case["keep"]


def test_crowding_epsilon_boundary(golden):
    for case in golden("nsga")["eps_cases"]:
        recs = records(case["objs"], [0.0] * len(case["objs"]))
        d0 = nsga_ref.crowding_distance(case["front"], recs, skip_on_equal=True)
        d1 = nsga_ref.crowding_distance(case["front"], recs, skip_on_equal=False)
        assert [d0[i] for i in case["front"]] == case["gt"]
        assert [d1[i] for i in case["front"]] == case["lt"]


def test_lambda_schedule(golden):
    g = golden("nsga")["lambda_schedule"]
    assert [nsga_ref.get_lambda(i, g["max_gen"]) for i in range(g["max_gen"])] == g["values"]
    assert nsga_ref.get_lambda(0, 1) == 50.0
    with pytest.raises(ZeroDivisionError):
        nsga_ref.get_lambda(0, 1, guarded=False)       # nsga_penalty.py:218 divides by MAX_GEN-1


def test_fpr_variants(golden):
    for case in golden("fpr")["cases"]:
        cm = nsga_ref.confusion(case["y_true"], case["y_pred"], case["classes"])
        assert nsga_ref.fpr_macro(cm, "all") == case["fpr_all"]                  # bit-exact: same np.mean of the same floats
        assert nsga_ref.fpr_macro(cm, "filtered") == case["fpr_filtered"]
        assert nsga_ref.fpr_macro(cm, "all") == case["fpr_vectorised"]


def test_fpr_all_zero_labels_quirk():
    # nsga_penalty.py:387 feeds argmax over an (N,1) array = all zeros: FPR == (1 - frac_pred0)/C
    rng = np.random.default_rng(0)
    c, n = 10, 1000
    pred = rng.integers(0, c, n)
    cm = nsga_ref.confusion(np.zeros(n, int), pred, c)
    assert nsga_ref.fpr_macro(cm, "all") == pytest.approx((1 - np.mean(pred == 0)) / c, rel=1e-12)


def test_infill_selection(golden):
    for case in golden("infill")["cases"]:
        recs = records(case["objs"], case["cv"])
        idx, _ = nsga_ref.select_infill_points(recs, case["k"])
        assert idx == case["indices"]


def test_quality_metrics(golden):
    for case in golden("quality")["cases"]:
        fronts = [np.array(f) for f in case["fronts"]]
        allp = np.vstack(fronts)
        mask = nsga_ref.nondominated_mask(allp)
        assert mask.tolist() == case["true_mask"]
        true = allp[mask]
        for i, f in enumerate(fronts):
            assert nsga_ref.generational_distance(f, true) == pytest.approx(case["gd"][i], rel=1e-12, abs=1e-15)
            assert nsga_ref.inverted_gd(f, true) == pytest.approx(case["igd"][i], rel=1e-12, abs=1e-15)
            s = nsga_ref.spread_metric(f, true)
            if math.isnan(case["spread"][i]):
                assert math.isnan(s)
            else:
                assert s == pytest.approx(case["spread"][i], rel=1e-12)
            for j, b in enumerate(fronts):
                assert nsga_ref.coverage_metric(f, b) == pytest.approx(case["coverage"][i][j])


def test_model_size_closed_form():
    # hand count quoted in SURVEY.md section 2.2: variant A, (16,3,no BN,1 block,1 FC), 10 classes
    hp = dict(filters=16, kernel_size=3, use_bn=False, residual_blocks=1, fc_layers=1, use_dropout=False)
    assert nsga_ref.param_count(hp, 10, "A") == 19674
    assert nsga_ref.model_size_mb(hp, 10, "A") == pytest.approx(0.0751, abs=5e-5)
    sizes_a, sizes_b = [], []
    for f in (16, 32, 64):
        for k in (3, 5):
            for bn in (True, False):
                for r in (1, 2, 3):
                    for fc in (1, 2, 3, 4):
                        hp = dict(filters=f, kernel_size=k, use_bn=bn, residual_blocks=r, fc_layers=fc, use_dropout=False)
                        sizes_a.append(nsga_ref.model_size_mb(hp, 10, "A"))
                        sizes_b.append(nsga_ref.model_size_mb(hp, 10, "B"))
    assert min(sizes_a) == pytest.approx(0.075, abs=1e-3) and max(sizes_a) == pytest.approx(51.97, abs=1e-2)
    assert min(sizes_b) == pytest.approx(0.031, abs=1e-3) and max(sizes_b) == pytest.approx(18.75, abs=1e-2)
    assert sum(s <= 2.5 for s in sizes_a) == 72 and sum(s <= 2.5 for s in sizes_b) == 108


def test_gp_posterior_restatement_vs_reference_surrogate(golden):
    """oracle/gp_ref.posterior with the stored (theta, X, alpha, L) must reproduce what the reference's
    SurrogateManager.predict(return_std=True) returned (sa_nsga_local.py:212-223)."""
    g = golden("surrogate")
    for case in g["cases"]:
        cols = case["feature_columns"]
        cats = {k: sorted({c.split("_")[-1] for c in cols if c.startswith(f"cat__{k}_")}) for k in ("use_bn", "use_dropout")}
        def encode(hp):
            row = [float(hp[k]) for k in ("filters", "kernel_size", "residual_blocks", "fc_layers")]
            for k in ("use_bn", "use_dropout"):
                row += [1.0 if str(hp[k]) == c else 0.0 for c in cats[k]]
            return row
        xq = np.array([encode(q) for q in case["queries"]])
        for key, mdl in case["models"].items():
            mean, std = gp_ref.posterior(xq, np.array(mdl["x_train"]), np.array(mdl["alpha"]), np.array(mdl["chol_lower"]),
                                         amplitude=mdl["amplitude"], length_scale=mdl["length_scale"], nu=mdl["nu"],
                                         noise=mdl["noise"], y_scale=mdl["scaler_scale"], y_shift=mdl["scaler_mean"])
            np.testing.assert_allclose(mean, case["pred"][key], rtol=1e-9, atol=1e-10)
            ref_std = np.array(case["std"][key])
            if mdl["scaler_var"] > 0:
                np.testing.assert_allclose(std, ref_std, rtol=1e-7, atol=1e-10)


def test_gp_posterior_restatement_vs_reference_mobo(golden):
    g = golden("mobo")
    for case in g["cases"]:
        cand = np.array(case["candidates"])
        mu = np.array(case["mu"])
        for j, mdl in enumerate(case["models"]):
            mean, _ = gp_ref.posterior(cand, np.array(mdl["x_train"]), np.array(mdl["alpha"]), np.array(mdl["chol_lower"]),
                                       amplitude=mdl["amplitude"], length_scale=mdl["length_scale"], nu=mdl["nu"],
                                       noise=mdl["noise"], y_scale=mdl["y_scale"], y_shift=mdl["y_shift"])
            np.testing.assert_allclose(mean, mu[:, j], rtol=1e-8, atol=1e-9)
        acq = -np.sum(mu[:, :3] + case["lam"] * mu[:, 3:4], axis=1)
        np.testing.assert_allclose(acq, case["acq"], rtol=1e-12)
        assert int(np.argmax(acq)) == case["argmax"]


def test_live_reference_matches_golden_when_available(golden):
    """In the build container the reference is present: re-run its own code and compare with the
    committed fixtures (guards against stale goldens).  Skipped on the GPU box."""
    from oracle import extract as ex
    if not ex.reference_available():
        pytest.skip("/root/reference not present")
    sa = ex.extract("ablation_study/sa_nsga_local.py", ["dominates", "fast_non_dominated_sort", "crowding_distance"])
    for case in golden("nsga")["cases"][:24]:
        recs = records(case["objs"], case["cv"])
        assert sa["fast_non_dominated_sort"](recs, case["lam"]) == case["fronts"]


def test_gp_objective_restatement_matches_sklearn():
    """oracle/gp_ref.log_marginal_likelihood (the blocked-sweep algorithm of csrc/gp_lml.cu in NumPy) against
    GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True), up to cond(K) ~ 1e9."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    from oracle import gp_ref

    rng = np.random.default_rng(5)
    x = rng.integers(0, 4, size=(70, 8)).astype(np.float64)
    x = np.unique(x, axis=0)
    y = np.sin(x[:, 0]) + 0.2 * x[:, 1] + 0.05 * rng.standard_normal(len(x))
    y = (y - y.mean()) / y.std()
    kernel = ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1)
    gpr = GaussianProcessRegressor(kernel=kernel, optimizer=None).fit(x, y)
    for theta in ([0.0, 0.0, np.log(0.1)], [2.0, 1.0, -6.0], [-2.0, -1.0, 0.5], [5.0, 3.0, -11.0]):
        want, want_grad = gpr.log_marginal_likelihood(np.array(theta), eval_gradient=True)
        for block in (1, 16):
            got, grad = gp_ref.log_marginal_likelihood(theta, x, y, 0, 1.5, block=block)
            assert got == pytest.approx(want, rel=1e-8, abs=1e-8)
            np.testing.assert_allclose(grad, want_grad, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(want_grad).max()))
    # the MOBO form, smooth kernel with only the 1e-10 jitter: cond(K) grows to ~3e9 at log l = 3
    xm = np.random.default_rng(7).uniform(size=(45, 6))
    ym = np.sin(3 * xm[:, 0]) + xm[:, 1] * xm[:, 2]
    g2 = GaussianProcessRegressor(kernel=Matern(nu=2.5), optimizer=None, normalize_y=True).fit(xm, ym)
    for t in np.linspace(-3.0, 3.0, 7):
        want, want_grad = g2.log_marginal_likelihood(np.array([t]), eval_gradient=True)
        got, grad = gp_ref.log_marginal_likelihood([t], xm, g2.y_train_, 1, 2.5)
        assert got == pytest.approx(want, rel=1e-8, abs=1e-8)
        np.testing.assert_allclose(grad, want_grad, rtol=1e-5, atol=1e-8)
    # not positive definite: duplicate rows without noise
    assert gp_ref.log_marginal_likelihood([0.0], np.zeros((5, 2)), np.arange(5.0), 1, 1.5, jitter=0.0)[0] == -np.inf
