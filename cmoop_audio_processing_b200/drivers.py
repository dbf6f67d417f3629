"""The reference's search loops with the hot functions injected.

The reference drivers stay what they are -- host-side Python that owns the ``random`` stream and the
record lists.  These mirrors exist because the reference scripts cannot be imported (they load data
from hard-coded paths and run the search at import, nsga_penalty.py:157-167,783); they follow the
reference control flow statement by statement so that, with the same ``random`` seed and the same
evaluation results, they visit the same populations:

* ``nsga2``        nsga_penalty.py:610-776          (constrained NSGA-II, adaptive penalty)
* ``sa_nsga2``     ablation_study/sa_nsga_local.py:436-554 (surrogate + Lamarckian local search);
                   ``local_search=False`` gives sa_nsga_penalty.py:522-637
* ``run_mobo``     mobo_penalty.py:343-487

Every hot function is a parameter (``ops``) defaulting to the CUDA-backed drop-ins, so the host logic
can be checked on a CPU-only box against the AST-extracted reference loop with the same fake evaluator.
Quality indicators (HV / IGD / Spread, compare.ipynb) can be recorded per generation.
"""
from __future__ import annotations

import random
import time
from copy import deepcopy
from types import SimpleNamespace

import numpy as np

from . import nsga as _nsga

CROSSOVER_PROB = 0.9          # nsga_penalty.py:201
LAMBDA_INITIAL, LAMBDA_FINAL = 1.0, 50.0


def default_ops(problem, *, surrogate=True, script="sa_nsga_local"):
    """Bundle of the drop-in functions a driver needs; override any of them for testing.

    ``script`` names the reference file whose copies are mirrored where the copies differ:
    * crowding_distance: nsga_penalty.py:518 skips an objective when ``f_max - f_min < EPSILON`` (CROWD_RANGE_LT);
      every other script (sa_nsga_penalty.py:437, sa_nsga_local.py:270) skips unless ``> EPSILON`` (CROWD_RANGE_GT).
      They differ only when the range equals EPSILON exactly.
    * SurrogateManager: sa_nsga_penalty.py:342-363 ``predict`` returns records (SurrogateManagerPenalty); the
      ablation_study/*local* scripts return dicts of arrays (+ stds)."""
    from functools import partial
    crowd_mode = _nsga.CROWD_RANGE_LT if script == "nsga_penalty" else _nsga.CROWD_RANGE_GT
    ops = SimpleNamespace(
        compute_objectives_and_constraints=problem.compute_objectives_and_constraints,
        evaluate_individual=problem.evaluate_individual,
        fast_non_dominated_sort=_nsga.fast_non_dominated_sort,
        crowding_distance=partial(_nsga.crowding_distance, crowd_mode=crowd_mode),
        dominates=_nsga.dominates,
        tournament_selection=_nsga.tournament_selection,
        crossover=_nsga.crossover,
        mutate=_nsga.mutate,
        initialize_population=_nsga.initialize_population,
    )
    if surrogate:
        from . import surrogate as s
        ops.SurrogateManager = s.SurrogateManagerPenalty if script == "sa_nsga_penalty" else s.SurrogateManager
        ops.perform_local_search = s.perform_local_search
        ops.select_infill_points = s.select_infill_points
    return ops


def _truncate(ops, combined, lam, pop_size):
    """Environmental selection, nsga_penalty.py:676-690."""
    new_pop = []
    for front in ops.fast_non_dominated_sort(combined, lam):
        if len(new_pop) + len(front) <= pop_size:
            new_pop.extend(combined[i] for i in front)
        else:
            remaining = pop_size - len(new_pop)
            dist = ops.crowding_distance(front, combined)
            ranked = sorted(front, key=lambda i: dist.get(i, 0), reverse=True)
            new_pop.extend(combined[i] for i in ranked[:remaining])
            break
    return new_pop


def _final_front(ops, pop_data):
    feas = [ind for ind in pop_data if ind["CV"] == 0]
    if not feas:
        return []
    fronts = ops.fast_non_dominated_sort(feas, LAMBDA_FINAL)
    return [feas[i] for i in fronts[0]] if fronts else []


def nsga2(pop_size, max_gen, ops, *, on_generation=None):
    """nsga_penalty.py:610-776.  Returns (pareto_set, per-generation populations, timings).  ``on_generation(gen,
    pop_data)`` runs INSIDE the generation's timed region (the reference times the loop body, sa_nsga_penalty.py:540,602)
    and its return value is kept under timings[gen]["indicators"]."""
    get_lambda = lambda g: _nsga.get_lambda(g, max_gen, LAMBDA_INITIAL, LAMBDA_FINAL)   # noqa: E731
    population = ops.initialize_population(pop_size)
    pop_data = ops.compute_objectives_and_constraints(population)
    history, timings = [], []
    for gen in range(max_gen):
        t0 = time.perf_counter()
        lam = get_lambda(gen)
        fronts = ops.fast_non_dominated_sort(pop_data, lam)
        for f in fronts:                                   # computed and unused in the reference too (626-629)
            ops.crowding_distance(f, pop_data)
        parents = [ops.tournament_selection(pop_data, lam, k=2) for _ in range(pop_size)]
        offspring = []
        for i1, i2 in zip(parents[0::2], parents[1::2]):
            p1, p2 = pop_data[i1]["hparams"], pop_data[i2]["hparams"]
            if random.random() < CROSSOVER_PROB:
                c1, c2 = ops.crossover(p1, p2)
            else:
                c1, c2 = deepcopy(p1), deepcopy(p2)
            offspring.append(ops.mutate(c1))
            offspring.append(ops.mutate(c2))
        if pop_size % 2 == 1:
            offspring.append(ops.mutate(deepcopy(pop_data[parents[-1]]["hparams"])))
        offspring = offspring[:pop_size]
        off_data = ops.compute_objectives_and_constraints(offspring)
        pop_data = _truncate(ops, pop_data + off_data, lam, pop_size)
        extra = on_generation(gen, pop_data) if on_generation else None     # e.g. HV / IGD / Spread of the generation
        timings.append({"generation": gen, "seconds": time.perf_counter() - t0, "true_evals": len(offspring),
                        "indicators": extra})
        history.append(list(pop_data))
    return _final_front(ops, pop_data), history, timings


def sa_nsga2(pop_size, max_gen, infill_percent, ops, *, local_search=True, on_generation=None):
    """ablation_study/sa_nsga_local.py:436-554 (local_search=False: sa_nsga_penalty.py:522-637)."""
    get_lambda = lambda g: _nsga.get_lambda(g, max_gen, LAMBDA_INITIAL, LAMBDA_FINAL)   # noqa: E731
    population = ops.initialize_population(pop_size)
    pop_data = ops.compute_objectives_and_constraints(population)
    sm = ops.SurrogateManager()
    sm.update([d["hparams"] for d in pop_data], pop_data)
    history, timings = [], []
    for gen in range(max_gen):
        t0 = time.perf_counter()
        lam = get_lambda(gen)
        if not local_search:
            ops.fast_non_dominated_sort(pop_data, lam)     # computed and unused in the reference too (sa_nsga_penalty.py:547)
        parents = [ops.tournament_selection(pop_data, lam) for _ in range(pop_size)]
        parent_hparams = [pop_data[i]["hparams"] for i in parents]
        offspring = []
        while len(offspring) < pop_size:
            p1, p2 = random.sample(parent_hparams, 2)
            c1, c2 = ops.crossover(p1, p2) if random.random() < CROSSOVER_PROB else (deepcopy(p1), deepcopy(p2))
            offspring.extend([ops.mutate(c1), ops.mutate(c2)])
        offspring = offspring[:pop_size]
        if local_search:
            preds, stds = sm.predict(offspring, return_std=True)
            predicted = [{"hparams": hp, "objs": [preds["neg_acc"][i], preds["size"][i], preds["fpr"][i]],
                          "stds": [stds["neg_acc"][i], stds["size"][i], stds["fpr"][i]], "CV": max(0, preds["cv"][i])}
                         for i, hp in enumerate(offspring)]
            offspring = ops.perform_local_search(predicted, sm)
            final_pred = sm.predict_and_structure(offspring)
        else:
            final_pred = sm.predict(offspring)             # sa_nsga_penalty.py:563: predict() returns the records
            if isinstance(final_pred, dict):               # a dict-returning (sa_nsga_local-form) manager was bound
                final_pred = sm.predict_and_structure(offspring)
        num_infill = max(1, int(pop_size * infill_percent))
        infill_idx, infill_hp = ops.select_infill_points(final_pred, num_infill)
        t_eval = time.perf_counter()
        infill_true = ops.compute_objectives_and_constraints(infill_hp)
        eval_s = time.perf_counter() - t_eval
        t_upd = time.perf_counter()
        sm.update(infill_hp, infill_true)
        upd_s = time.perf_counter() - t_upd
        off_data = list(final_pred)
        for i, true_res in enumerate(infill_true):
            off_data[infill_idx[i]] = true_res
        pop_data = _truncate(ops, pop_data + off_data, lam, pop_size)
        extra = on_generation(gen, pop_data) if on_generation else None     # e.g. HV / IGD / Spread of the generation
        timings.append({"generation": gen, "seconds": time.perf_counter() - t0, "true_evals": len(infill_hp),
                        "eval_seconds": eval_s, "update_seconds": upd_s, "indicators": extra})
        history.append(list(pop_data))
    return _final_front(ops, pop_data), history, timings


def run_mobo(initial_samples, max_iterations, candidate_batch, problem, *, min_accuracy=0.90, max_model_size=2.5,
             max_fpr=0.09, ops=None):
    """mobo_penalty.py:343-487: 4 independent GPs, 500 uniform candidates, penalised-sum acquisition."""
    from . import surrogate as s
    evaluate = ops.evaluate_individual if ops else problem.evaluate_individual
    dim = 6
    x_vec = np.zeros((initial_samples, dim))
    y_objs = np.zeros((initial_samples, 3))
    y_cv = np.zeros((initial_samples, 1))

    def cv_of(acc, size_mb, fpr):
        return max(0.0, min_accuracy - acc) + max(0.0, size_mb - max_model_size) + max(0.0, fpr - max_fpr)

    for i in range(initial_samples):
        hp = {k: random.choice(v) for k, v in _nsga.HPARAM_SPACE.items()}
        acc, size_mb, fpr = evaluate(hp)
        x_vec[i] = s.hparams_to_vector(hp)
        y_objs[i] = [-acc, size_mb, fpr]
        y_cv[i, 0] = cv_of(acc, size_mb, fpr)
    all_hp = [s.vector_to_hparams(x_vec[i]) for i in range(initial_samples)]
    for it in range(max_iterations):
        lam = LAMBDA_INITIAL + it / float(max_iterations - 1) * (LAMBDA_FINAL - LAMBDA_INITIAL)
        gp_objs = s.train_gps(x_vec, y_objs)
        gp_cv = s.train_gps(x_vec, y_cv)[0]
        candidates = np.random.rand(candidate_batch, dim)
        acq = s.penalized_acquisition(candidates, gp_objs, gp_cv, lam)
        x_next = candidates[int(np.argmax(acq))]
        hp_next = s.vector_to_hparams(x_next)
        acc, size_mb, fpr = evaluate(hp_next)
        x_vec = np.vstack([x_vec, x_next.reshape(1, -1)])           # un-rounded, as mobo_penalty.py:401
        y_objs = np.vstack([y_objs, [-acc, size_mb, fpr]])
        y_cv = np.vstack([y_cv, [[cv_of(acc, size_mb, fpr)]]])
        all_hp.append(hp_next)
    feas = [i for i in range(len(y_cv)) if y_cv[i, 0] <= 1e-8]
    from .quality import nondominated_mask
    if not feas:
        return [], (x_vec, y_objs, y_cv)
    mask = nondominated_mask(y_objs[feas])
    return [(all_hp[i], y_objs[i], y_cv[i, 0]) for i, m in zip(feas, mask) if m], (x_vec, y_objs, y_cv)


def front_indicators(pop_data, reference_front=None, ref_point=None):
    """HV / GD / IGD / Spread of the feasible non-dominated records (compare.ipynb conventions)."""
    from . import quality
    feas = np.array([r["objs"] for r in pop_data if r["CV"] == 0], dtype=np.float64)
    if len(feas) == 0:
        return {"hv": 0.0, "gd": float("nan"), "igd": float("nan"), "spread": float("nan"), "n": 0}
    front = feas[quality.nondominated_mask(feas)]
    ref = quality.reference_point(feas) if ref_point is None else np.asarray(ref_point, np.float64)
    out = {"hv": quality.hypervolume(front, ref), "n": int(len(front))}
    if reference_front is not None and len(reference_front):
        out.update(quality.front_metrics(front, np.asarray(reference_front, np.float64)))
    return out
