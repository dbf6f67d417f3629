"""Generate tests/golden/*.json by running the REFERENCE'S OWN functions.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference has no tests or golden vectors of its own, so the
pins are constructed here: functions are AST-extracted from the read-only sources
(oracle/extract.py), executed on seeded synthetic inputs, and the inputs/outputs
are committed as small JSON fixtures.  Floats are written with repr() round-trip
precision, so comparisons against them can be bit-exact.
"""
from __future__ import annotations

import json
import os
import random

import numpy as np

from oracle import extract as ex
from oracle import gp_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

GENES = {
    "filters": [16, 32, 64], "kernel_size": [3, 5], "use_bn": [True, False],
    "residual_blocks": [1, 2, 3], "fc_layers": [1, 2, 3, 4], "use_dropout": [True, False],
}


def rand_hp(rng: random.Random) -> dict:
    return {k: rng.choice(v) for k, v in GENES.items()}


def make_records(rng: np.random.Generator, n: int, m: int, style: str) -> list[dict]:
    """Synthetic {'hparams','objs','CV'} records with heavy ties / duplicates."""
    pyr = random.Random(int(rng.integers(1 << 30)))
    if style == "grid":            # few distinct values -> many ties
        objs = rng.integers(0, 4, size=(n, m)).astype(float) / 4.0
    elif style == "dup":           # duplicated rows
        base = rng.random((max(1, n // 3), m))
        objs = base[rng.integers(0, len(base), size=n)]
    elif style == "chain":         # totally ordered -> n fronts
        t = np.sort(rng.random(n))
        objs = np.stack([t + 0.01 * j for j in range(m)], axis=1)
    elif style == "kws":           # objective-shaped: (-acc, size_mb, fpr)
        cols = [-(0.5 + 0.5 * rng.random(n)), rng.choice([0.075, 0.31, 1.1, 2.81, 9.4, 51.97], size=n),
                0.1 * rng.random(n)]
        objs = np.stack(cols[:m], axis=1)
    else:
        objs = rng.random((n, m))
    if style in ("grid", "dup"):
        cv = rng.choice([0.0, 0.0, 0.25, 0.5], size=n)
    elif style == "allinf":
        cv = 0.1 + rng.random(n)
    else:
        cv = np.where(rng.random(n) < 0.5, 0.0, rng.random(n))
    return [{"hparams": rand_hp(pyr), "objs": [float(v) for v in objs[i]], "CV": float(cv[i])} for i in range(n)]


def golden_nsga() -> dict:
    sa = ex.extract("ablation_study/sa_nsga_local.py",
                    ["dominates", "fast_non_dominated_sort", "crowding_distance", "get_lambda"])
    ns = ex.extract("nsga_penalty.py", ["dominates", "fast_non_dominated_sort", "crowding_distance"])
    rng = np.random.default_rng(20251018)
    cases = []
    sizes = [1, 2, 3, 5, 8, 13, 16, 31, 32, 33, 47, 64, 79, 128]
    styles = ["rand", "grid", "dup", "chain", "kws", "allinf"]
    for ci in range(72):
        n = sizes[ci % len(sizes)]
        m = 2 if ci % 3 == 0 else 3
        style = styles[ci % len(styles)]
        lam = [1.0, 17.896551724137932, 50.0, 0.0][ci % 4]
        recs = make_records(rng, n, m, style)
        fronts = sa["fast_non_dominated_sort"](recs, lam)
        fronts_ns = ns["fast_non_dominated_sort"](recs, lam)
        assert fronts == fronts_ns, "reference copies disagree"
        crowd_gt = [sa["crowding_distance"](f, recs) for f in fronts]      # range > EPSILON form
        crowd_lt = [ns["crowding_distance"](f, recs) for f in fronts]      # range < EPSILON skip form
        # (mu + lambda) truncation exactly as sa_nsga_local.py:506-515 with pop = n // 2
        pop = max(1, n // 2)
        keep = []
        for f in fronts:
            if len(keep) + len(f) <= pop:
                keep.extend(f)
            else:
                d = sa["crowding_distance"](f, recs)
                keep.extend(sorted(f, key=lambda i: d.get(i, 0), reverse=True)[: pop - len(keep)])
                break
        cases.append({
            "n": n, "m": m, "style": style, "lam": lam,
            "objs": [r["objs"] for r in recs], "cv": [r["CV"] for r in recs],
            "fronts": fronts,
            "crowd_gt": [[d[i] for i in f] for d, f in zip(crowd_gt, fronts)],
            "crowd_lt": [[d[i] for i in f] for d, f in zip(crowd_lt, fronts)],
            "pop": pop, "keep": keep,
        })
    # near-EPSILON range cases for the two crowding variants
    eps_cases = []
    for span in [1e-6, 9.999999e-7, 1.0000001e-6, 0.0]:
        recs = [{"hparams": {}, "objs": [0.25 + span * t, 1.0 - t], "CV": 0.0} for t in (0.0, 0.5, 1.0)]
        f = [0, 1, 2]
        eps_cases.append({"objs": [r["objs"] for r in recs], "front": f,
                          "gt": [sa["crowding_distance"](f, recs)[i] for i in f],
                          "lt": [ns["crowding_distance"](f, recs)[i] for i in f]})
    lam_sched = {"max_gen": 30, "values": [sa["get_lambda"](g) for g in range(30)]}
    return {"source": "ablation_study/sa_nsga_local.py:240-277 + nsga_penalty.py:448-524",
            "cases": cases, "eps_cases": eps_cases, "lambda_schedule": lam_sched}


def golden_fpr() -> dict:
    full = ex.extract("nsga_penalty.py", ["calculate_fpr"])["calculate_fpr"]
    filt = ex.extract("ablation_study/sa_nsga_local.py", ["calculate_fpr"])["calculate_fpr"]
    vect = ex.extract("ablation_study/init_sa_nsga_local.py", ["calculate_fpr"])["calculate_fpr"]
    rng = np.random.default_rng(7)
    cases = []
    for ci in range(24):
        c = [2, 10, 11, 12, 397][ci % 5]
        n = [1, 17, 64, 768, 3000][ci % 5]
        y_true = rng.integers(0, c, size=n)
        if ci % 4 == 0:
            y_true[:] = 0                              # the nsga_penalty.py:387 quirk (argmax of (N,1))
        if ci % 6 == 1:
            y_true[:] = c - 1                          # one class only -> denom 0 for it
        y_pred = np.where(rng.random(n) < 0.7, y_true, rng.integers(0, c, size=n))
        cases.append({"classes": c, "y_true": y_true.tolist(), "y_pred": y_pred.tolist(),
                      "fpr_all": float(full(y_true, y_pred, c)),
                      "fpr_filtered": float(filt(y_true, y_pred, c)),
                      "fpr_vectorised": float(vect(y_true, y_pred, c))})
    return {"source": "nsga_penalty.py:351-364, sa_nsga_local.py:138-141, init_sa_nsga_local.py:137-143",
            "cases": cases}


def golden_infill() -> dict:
    fn = ex.extract("ablation_study/sa_nsga_local.py", ["select_infill_points"])["select_infill_points"]
    fn2 = ex.extract("sa_nsga_penalty.py", ["select_infill_points"])["select_infill_points"]
    rng = np.random.default_rng(11)
    cases = []
    for ci in range(16):
        n = [1, 4, 15, 64][ci % 4]
        recs = make_records(rng, n, 3, ["rand", "kws", "allinf", "grid"][ci % 4])
        k = max(1, int(n * [0.2, 0.334][ci % 2]))
        idx, hps = fn(recs, k)
        idx2, _ = fn2(recs, k)
        assert list(idx) == list(idx2)
        cases.append({"objs": [r["objs"] for r in recs], "cv": [r["CV"] for r in recs], "k": k,
                      "indices": [int(i) for i in idx]})
    return {"source": "sa_nsga_local.py:303-349 == sa_nsga_penalty.py:472-518", "cases": cases}


def golden_quality() -> dict:
    nb = ex.extract_notebook("compare.ipynb", ["dominates_min", "generational_distance", "inverted_gd",
                                                "spread_metric", "coverage_metric"])
    rng = np.random.default_rng(5)
    cases = []
    for ci in range(10):
        fronts = []
        for _ in range(3):
            n = int(rng.integers(1, 24))
            pts = np.stack([-(0.6 + 0.4 * rng.random(n)), 0.05 + 3.0 * rng.random(n), 0.1 * rng.random(n)], axis=1)
            if ci % 3 == 0:
                pts = np.round(pts, 1)
            fronts.append(pts)
        allp = np.vstack(fronts)
        mask = np.ones(len(allp), dtype=bool)
        for i in range(len(allp)):                       # compare.ipynb section 6, via its dominates_min
            for j in range(len(allp)):
                if i != j and nb["dominates_min"](allp[j], allp[i]):
                    mask[i] = False
                    break
        true_front = allp[mask]
        cases.append({
            "fronts": [f.tolist() for f in fronts], "true_mask": mask.tolist(),
            "gd": [float(nb["generational_distance"](f, true_front)) for f in fronts],
            "igd": [float(nb["inverted_gd"](f, true_front)) for f in fronts],
            "spread": [float(nb["spread_metric"](f, true_front)) for f in fronts],
            "coverage": [[float(nb["coverage_metric"](a, b)) for b in fronts] for a in fronts],
        })
    return {"source": "compare.ipynb cell 0 sections 6-9", "cases": cases}


def _dump_gpr(gpr) -> dict:
    u = gp_ref.unpack_sklearn(gpr)
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in u.items()}


def golden_surrogate() -> dict:
    """SurrogateManager.update / predict(return_std) / predict_and_structure and
    perform_local_search from ablation_study/sa_nsga_local.py, fixed seeds."""
    import sklearn
    ns = ex.extract("ablation_study/sa_nsga_local.py",
                    ["SurrogateManager", "perturb_hparams", "lcb_dominates", "perform_local_search"])
    cases = []
    for seed, n_train, n_query in [(0, 15, 15), (1, 40, 24), (2, 6, 9)]:
        np.random.seed(seed)          # sklearn restarts draw from the global NumPy RNG (_gpr.py:330)
        random.seed(seed)
        pyr = random.Random(100 + seed)
        rng = np.random.default_rng(200 + seed)
        hps = [rand_hp(pyr) for _ in range(n_train)]
        recs = []
        for hp in hps:                 # smooth synthetic response surface + noise
            acc = 0.7 + 0.002 * hp["filters"] + 0.02 * hp["residual_blocks"] - 0.01 * hp["fc_layers"] \
                + 0.01 * hp["use_bn"] + 0.01 * rng.standard_normal()
            size = 0.002 * hp["filters"] ** 1.5 * hp["residual_blocks"] * (hp["kernel_size"] / 3.0) ** 2 \
                + 0.05 * hp["fc_layers"]
            fpr = max(0.0, 0.12 - 0.001 * hp["filters"] + 0.005 * rng.standard_normal())
            cv = max(0.0, 0.9 - acc) + max(0.0, size - 2.5) + max(0.0, fpr - 0.09)
            recs.append({"hparams": hp, "objs": [-acc, size, fpr], "CV": cv})
        sm = ns["SurrogateManager"]()
        sm.update(hps, recs)
        queries = [rand_hp(pyr) for _ in range(n_query)]
        preds, stds = sm.predict(queries, return_std=True)
        structured = sm.predict_and_structure(queries)
        # local search (consumes the global `random` stream)
        off = [{"hparams": dict(q), "objs": [preds["neg_acc"][i], preds["size"][i], preds["fpr"][i]],
                "stds": [stds["neg_acc"][i], stds["size"][i], stds["fpr"][i]], "CV": max(0, preds["cv"][i])}
               for i, q in enumerate(queries)]
        random.seed(1000 + seed)
        improved = ns["perform_local_search"](off, sm)
        rand_after = random.random()
        td = sm.training_data
        cases.append({
            "seed": seed, "train_hparams": hps,
            "train_objs": [r["objs"] for r in recs], "train_cv": [r["CV"] for r in recs],
            "dedup_rows": td[["filters", "kernel_size", "residual_blocks", "fc_layers", "use_bn", "use_dropout"]]
            .astype(object).values.tolist(),
            "feature_columns": [str(c) for c in sm.preprocessor.get_feature_names_out()],
            "models": {k: dict(_dump_gpr(m), scaler_mean=float(sm.scalers[k].mean_[0]),
                               scaler_var=float(sm.scalers[k].var_[0]),
                               scaler_scale=float(sm.scalers[k].scale_[0])) for k, m in sm.models.items()},
            "queries": queries,
            "pred": {k: v.tolist() for k, v in preds.items()},
            "std": {k: np.asarray(v).tolist() for k, v in stds.items()},
            "structured_cv": [float(r["CV"]) for r in structured],
            "ls_seed": 1000 + seed, "ls_improved": improved, "ls_rand_after": rand_after,
        })
    return {"source": "ablation_study/sa_nsga_local.py:169-234, 351-433", "sklearn": sklearn.__version__,
            "cases": cases}


def golden_mobo() -> dict:
    import sklearn
    ns = ex.extract("mobo_penalty.py", ["train_gps", "predict_gps", "penalized_acquisition",
                                        "hparams_to_vector", "vector_to_hparams", "get_lambda_it"])
    cases = []
    for seed, n in [(0, 15), (1, 30)]:
        rng = np.random.default_rng(300 + seed)
        pyr = random.Random(400 + seed)
        hps = [rand_hp(pyr) for _ in range(n)]
        x = np.stack([ns["hparams_to_vector"](hp) for hp in hps])
        if seed == 1:                                     # mobo_penalty.py:401 appends un-rounded candidates
            x[-5:] = rng.random((5, 6))
        y = np.stack([-(0.7 + 0.2 * x[:, 0] + 0.05 * x[:, 3] + 0.01 * rng.standard_normal(n)),
                      0.1 + 3.0 * x[:, 0] * (0.5 + x[:, 3]) + 0.2 * x[:, 4],
                      0.1 - 0.05 * x[:, 0] + 0.005 * rng.standard_normal(n)], axis=1)
        cv = np.maximum(0, 0.9 + y[:, 0]) + np.maximum(0, y[:, 1] - 2.5) + np.maximum(0, y[:, 2] - 0.09)
        gp_objs = ns["train_gps"](x, y)
        gp_cv = ns["train_gps"](x, cv[:, None])[0]
        cand = rng.random((500, 6))
        lam = ns["get_lambda_it"](7)
        acq = ns["penalized_acquisition"](cand, gp_objs, gp_cv, lam)
        mu = ns["predict_gps"](gp_objs + [gp_cv], cand)
        cases.append({"x": x.tolist(), "y": y.tolist(), "cv": cv.tolist(), "lam": lam,
                      "models": [_dump_gpr(g) for g in gp_objs + [gp_cv]],
                      "candidates": cand.tolist(), "mu": mu.tolist(), "acq": acq.tolist(),
                      "argmax": int(np.argmax(acq)),
                      "decoded": ns["vector_to_hparams"](cand[int(np.argmax(acq))])})
    enc = [{"hp": hp, "vec": ns["hparams_to_vector"](hp).tolist()} for hp in
           [{"filters": f, "kernel_size": k, "use_bn": b, "residual_blocks": r, "fc_layers": fc, "use_dropout": d}
            for f in GENES["filters"] for k in GENES["kernel_size"] for b in (True, False)
            for r in (1, 3) for fc in (1, 4) for d in (True, False)]]
    return {"source": "mobo_penalty.py:252-338", "sklearn": sklearn.__version__, "cases": cases, "encoding": enc}


def main() -> None:
    if not ex.reference_available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    for name, fn in [("nsga", golden_nsga), ("fpr", golden_fpr), ("infill", golden_infill),
                     ("quality", golden_quality), ("surrogate", golden_surrogate), ("mobo", golden_mobo)]:
        data = fn()
        path = os.path.join(OUT, f"{name}.json")
        with open(path, "w") as fh:
            json.dump(data, fh)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
