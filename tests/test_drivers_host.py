"""Host driver logic (drivers.py) against the reference's own loops, AST-extracted and run with the same fake
evaluator, the same `random` seed and the same (oracle) selection functions.  CPU only; needs /root/reference."""
import random
from types import SimpleNamespace

import numpy as np
import pytest

from cmoop_audio_processing_b200 import drivers, nsga, surrogate
from oracle import extract as ex
from oracle import nsga_ref

pytestmark = pytest.mark.skipif(not ex.reference_available(), reason="/root/reference not present")


def fake_eval(hp):
    acc = 0.70 + 0.003 * hp["filters"] + 0.02 * hp["residual_blocks"] - 0.01 * hp["fc_layers"] + 0.02 * hp["use_bn"]
    size = nsga_ref.model_size_mb(hp, 10, "B")
    fpr = 0.13 - 0.0008 * hp["filters"] - 0.01 * hp["use_dropout"] + 0.002 * hp["kernel_size"]
    return acc, size, fpr


def fake_compute(thr):
    def compute(population):
        out = []
        for hp in population:
            acc, size, fpr = fake_eval(hp)
            cv = max(0.0, thr[0] - acc) + max(0.0, size - thr[1]) + max(0.0, fpr - thr[2])
            out.append({"hparams": hp, "objs": [-acc, size, fpr], "CV": cv})
        return out
    return compute


def strip(pop):
    return [(r["hparams"], r["objs"], r["CV"]) for r in pop]


def test_nsga2_matches_reference_loop(monkeypatch):
    import pandas as pd
    monkeypatch.setattr(pd.DataFrame, "to_csv", lambda self, *a, **k: None)
    ref = ex.extract("nsga_penalty.py", ["get_lambda", "initialize_population", "dominates", "fast_non_dominated_sort",
                                         "crowding_distance", "tournament_selection", "crossover", "mutate", "nsga2"])
    ref["MAX_GEN"] = 4
    ref["compute_objectives_and_constraints"] = fake_compute((0.9, 2.5, 0.1))
    for seed in (0, 1):
        random.seed(seed)
        pareto_ref, dfs = ref["nsga2"](9, 4)
        ops = SimpleNamespace(compute_objectives_and_constraints=fake_compute((0.9, 2.5, 0.1)),
                              fast_non_dominated_sort=nsga_ref.fast_non_dominated_sort,
                              crowding_distance=lambda f, r: nsga_ref.crowding_distance(f, r, skip_on_equal=False),
                              tournament_selection=nsga.tournament_selection, crossover=nsga.crossover,
                              mutate=nsga.mutate, initialize_population=nsga.initialize_population)
        random.seed(seed)
        pareto, history, timings = drivers.nsga2(9, 4, ops)
        assert strip(pareto) == strip(pareto_ref)
        last = dfs[-1]
        assert [-r["objs"][0] for r in history[-1]] == last["Accuracy"].tolist()
        assert [r["CV"] for r in history[-1]] == last["CV"].tolist()
        assert len(timings) == 4 and all(t["true_evals"] == 9 for t in timings)


def test_sa_nsga2_with_local_search_matches_reference_loop():
    ref = ex.extract("ablation_study/sa_nsga_local.py", None)
    ref["MAX_GEN"] = 3
    ref["compute_objectives_and_constraints"] = fake_compute((0.90, 2.5, 0.09))
    for seed in (3,):
        random.seed(seed)
        np.random.seed(seed)
        pareto_ref, dfs = ref["nsga2"](8, 3, 0.334)
        ops = SimpleNamespace(compute_objectives_and_constraints=fake_compute((0.90, 2.5, 0.09)),
                              fast_non_dominated_sort=nsga_ref.fast_non_dominated_sort,
                              crowding_distance=nsga_ref.crowding_distance,
                              tournament_selection=nsga.tournament_selection, crossover=nsga.crossover,
                              mutate=nsga.mutate, initialize_population=nsga.initialize_population,
                              SurrogateManager=ref["SurrogateManager"],          # CPU sklearn class of the reference
                              perform_local_search=surrogate.perform_local_search,
                              select_infill_points=surrogate.select_infill_points)
        random.seed(seed)
        np.random.seed(seed)
        pareto, history, timings = drivers.sa_nsga2(8, 3, 0.334, ops, local_search=True)
        assert strip(pareto) == strip(pareto_ref)
        assert [r["CV"] for r in history[-1]] == dfs[-1]["CV"].tolist()
        assert all(t["true_evals"] == max(1, int(8 * 0.334)) for t in timings)


class _HostGPGroup:
    """CPU stand-in for surrogate.DeviceGPGroup (no GPU in the CPU suite): the same (mean, std) contract computed by the
    fitted scikit-learn models, so the HOST logic of SurrogateManager / SurrogateManagerPenalty (encoding, table,
    record structure, CV clamp) is what the replay exercises."""

    def __init__(self, gprs, y_affine):
        self.gprs, self.y_affine = gprs, y_affine

    @classmethod
    def from_sklearn(cls, gprs, y_affine=None):
        return cls(gprs, y_affine)

    def predict(self, xq, return_std=True):
        mean, std = [], []
        for g, (scale, shift) in zip(self.gprs, self.y_affine):
            m, s = g.predict(np.atleast_2d(xq), return_std=True)
            mean.append(np.ravel(m) * scale + shift)
            std.append(np.ravel(s) * scale)
        return np.array(mean), np.array(std)

    def close(self):
        pass


def test_sa_nsga_penalty_loop_with_rebound_manager(monkeypatch):
    """sa_nsga_penalty.py:522-637 (AST-extracted) run twice with the same seeds: once as it is, once with
    ``SurrogateManager`` rebound to the repo's SurrogateManagerPenalty exactly as INTEGRATION.md section 3 does --
    ``predict(hparams)`` must hand records to select_infill_points (call site :563) -- and against the driver mirror
    drivers.sa_nsga2(local_search=False)."""
    monkeypatch.setattr(surrogate, "DeviceGPGroup", _HostGPGroup)
    monkeypatch.setattr(surrogate, "fit_gprs_parallel",
                        lambda kernels, x, ys, n_restarts_optimizer=10, backend=None, **kw:
                        _fit_sequential(kernels, x, ys, n_restarts_optimizer))
    thr = (0.75, 2.5, 0.09)
    seed = 5

    def run_reference(manager_cls):
        ref = ex.extract("sa_nsga_penalty.py", None)
        ref["MAX_GEN"] = 3
        ref["compute_objectives_and_constraints"] = fake_compute(thr)
        if manager_cls is not None:
            ref["SurrogateManager"] = manager_cls                    # the INTEGRATION.md rebinding
        random.seed(seed)
        np.random.seed(seed)
        return ref["nsga2"](8, 3, 0.25)

    pareto_ref, dfs_ref = run_reference(None)
    pareto_reb, dfs_reb = run_reference(surrogate.SurrogateManagerPenalty)
    assert [r["hparams"] for r in pareto_reb] == [r["hparams"] for r in pareto_ref]
    for a, b in zip(dfs_reb, dfs_ref):
        assert a[list(nsga.HPARAM_SPACE)].to_dict("records") == b[list(nsga.HPARAM_SPACE)].to_dict("records")
        np.testing.assert_allclose(a[["Accuracy", "Size_MB", "FPR", "CV"]].to_numpy(float),
                                   b[["Accuracy", "Size_MB", "FPR", "CV"]].to_numpy(float), rtol=1e-9, atol=1e-12)

    ref = ex.extract("sa_nsga_penalty.py", ["SurrogateManager"])
    ops = SimpleNamespace(compute_objectives_and_constraints=fake_compute(thr),
                          fast_non_dominated_sort=nsga_ref.fast_non_dominated_sort,
                          crowding_distance=nsga_ref.crowding_distance,
                          tournament_selection=nsga.tournament_selection, crossover=nsga.crossover,
                          mutate=nsga.mutate, initialize_population=nsga.initialize_population,
                          SurrogateManager=ref["SurrogateManager"],
                          select_infill_points=surrogate.select_infill_points)
    random.seed(seed)
    np.random.seed(seed)
    pareto, history, timings = drivers.sa_nsga2(8, 3, 0.25, ops, local_search=False)
    assert strip(pareto) == strip(pareto_ref)
    assert [r["CV"] for r in history[-1]] == dfs_ref[-1]["CV"].tolist()
    assert all(t["true_evals"] == 2 for t in timings)


def _fit_sequential(kernels, x, ys, n_restarts):
    """The reference's own fit order (one GaussianProcessRegressor after the other, global NumPy RNG)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    return [GaussianProcessRegressor(kernel=k, n_restarts_optimizer=n_restarts).fit(x, y) for k, y in zip(kernels, ys)]


def test_penalty_manager_predict_contract(monkeypatch):
    """SurrogateManagerPenalty.predict: records with Python-level 'objs' triples and CV clamped at 0; RuntimeError
    before the first update (sa_nsga_penalty.py:344-345)."""
    monkeypatch.setattr(surrogate, "DeviceGPGroup", _HostGPGroup)
    monkeypatch.setattr(surrogate, "fit_gprs_parallel",
                        lambda kernels, x, ys, n_restarts_optimizer=10, backend=None, **kw:
                        _fit_sequential(kernels, x, ys, 0))
    sm = surrogate.SurrogateManagerPenalty()
    with pytest.raises(RuntimeError):
        sm.predict([{}])
    random.seed(1)
    pop = nsga.initialize_population(12)
    sm.update(pop, fake_compute((0.9, 2.5, 0.09))(pop))
    recs = sm.predict(pop[:5])
    assert isinstance(recs, list) and len(recs) == 5
    for hp, r in zip(pop[:5], recs):
        assert set(r) == {"hparams", "objs", "CV"} and r["hparams"] is hp and len(r["objs"]) == 3 and r["CV"] >= 0
    arrays = sm.predict_arrays(pop[:5])
    assert set(arrays) == {"neg_acc", "size", "fpr", "cv"}
    chosen, hps = surrogate.select_infill_points(recs, 2)             # the call the reference makes next (:568)
    assert len(chosen) == 2 and all(h in pop[:5] for h in hps)


def test_default_ops_crowding_mode_per_script():
    """nsga_penalty.py:518 skips an objective on (range < EPS); every other script on not (range > EPS)."""
    class _P:
        compute_objectives_and_constraints = evaluate_individual = None
    assert drivers.default_ops(_P(), surrogate=False, script="nsga_penalty").crowding_distance.keywords == {
        "crowd_mode": nsga.CROWD_RANGE_LT}
    assert drivers.default_ops(_P(), surrogate=False).crowding_distance.keywords == {"crowd_mode": nsga.CROWD_RANGE_GT}
    ops = drivers.default_ops(_P(), script="sa_nsga_penalty")
    assert ops.SurrogateManager is surrogate.SurrogateManagerPenalty
