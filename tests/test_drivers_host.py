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
