"""Pure-Python/NumPy restatement of the reference's selection-side arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines whose behaviour it restates.  Pinned by tests/golden/nsga_*.json,
which were produced by the reference's *own* functions (oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

EPSILON = 1e-6  # nsga_penalty.py:203 / sa_nsga_penalty.py:119


def get_lambda(gen: int, max_gen: int, lam0: float = 1.0, lam1: float = 50.0, guarded: bool = True) -> float:
    """Linear penalty schedule.  nsga_penalty.py:217-219 (unguarded: ZeroDivisionError
    at max_gen == 1) and sa_nsga_penalty.py:130-132 (guarded: frac = 1.0)."""
    if max_gen > 1 or not guarded:
        frac = gen / float(max_gen - 1)
    else:
        frac = 1.0
    return lam0 + frac * (lam1 - lam0)


def penalised(objs, cv: float, lam: float) -> list[float]:
    """P_i = f_i + lam*CV, multiply then add in fp64 (nsga_penalty.py:451-452)."""
    pen = lam * cv
    return [f + pen for f in objs]


def dominates(a: dict, b: dict, lam: float) -> bool:
    """Penalised Pareto dominance.  nsga_penalty.py:448-461 == sa_nsga_local.py:240-245."""
    pa = penalised(a["objs"], a["CV"], lam)
    pb = penalised(b["objs"], b["CV"], lam)
    no_worse = True
    better = False
    for x, y in zip(pa, pb):
        if x > y:
            no_worse = False
            break
        if x < y:
            better = True
    return no_worse and better


def fast_non_dominated_sort(results: list[dict], lam: float) -> list[list[int]]:
    """Deb's fast non-dominated sort, fronts in the reference's discovery order.

    sa_nsga_penalty.py:392-419 / sa_nsga_local.py:247-264: front 0 ascending index;
    front i+1 lists q in the order its dominator count reaches zero while walking
    (p in front i in order, q ascending inside S[p]).  Empty input -> [] (the
    nsga_penalty.py:467-501 copy raises IndexError instead).
    """
    n = len(results)
    dominated_by = [[] for _ in range(n)]   # S[p]
    count = [0] * n                         # n[p]
    first = []
    for p in range(n):
        for q in range(n):
            if p == q:
                continue
            if dominates(results[p], results[q], lam):
                dominated_by[p].append(q)
            elif dominates(results[q], results[p], lam):
                count[p] += 1
        if count[p] == 0:
            first.append(p)
    fronts = []
    cur = first
    while cur:
        fronts.append(cur)
        nxt = []
        for p in cur:
            for q in dominated_by[p]:
                count[q] -= 1
                if count[q] == 0:
                    nxt.append(q)
        cur = nxt
    return fronts


def crowding_distance(front: list[int], results: list[dict], eps: float = EPSILON,
                      skip_on_equal: bool = True) -> dict[int, float]:
    """Crowding distance on RAW objectives.

    sa_nsga_penalty.py:421-442 (range must be > eps: ``skip_on_equal=True``) and
    nsga_penalty.py:504-524 (skips when range < eps: ``skip_on_equal=False``).
    Stable sort per objective; end points are assigned inf; interior points get
    ``+= (next - prev) / (max - min)`` objective by objective in index order.
    """
    if not front:
        return {}
    dist = {i: 0.0 for i in front}
    m_objs = len(results[0]["objs"])
    for m in range(m_objs):
        order = sorted(front, key=lambda i: results[i]["objs"][m])
        dist[order[0]] = math.inf
        dist[order[-1]] = math.inf
        lo = results[order[0]]["objs"][m]
        hi = results[order[-1]]["objs"][m]
        span = hi - lo
        skip = (not span > eps) if skip_on_equal else (span < eps)
        if skip:
            continue
        for k in range(1, len(order) - 1):
            dist[order[k]] += (results[order[k + 1]]["objs"][m] - results[order[k - 1]]["objs"][m]) / span
    return dist


def environmental_selection(combined: list[dict], lam: float, pop_size: int) -> list[int]:
    """Indices kept by the (mu+lambda) truncation, nsga_penalty.py:676-690 /
    sa_nsga_local.py:504-515: whole fronts while they fit, then the overflowing
    front stably sorted by descending crowding distance."""
    keep: list[int] = []
    for front in fast_non_dominated_sort(combined, lam):
        if len(keep) + len(front) <= pop_size:
            keep.extend(front)
        else:
            room = pop_size - len(keep)
            d = crowding_distance(front, combined)
            ranked = sorted(front, key=lambda i: d[i], reverse=True)
            keep.extend(ranked[:room])
            break
    return keep


def constraint_violation(acc: float, size_mb: float, fpr: float, min_acc: float, max_size: float,
                         max_fpr: float) -> float:
    """CV = g1 + g2 + g3, nsga_penalty.py:433-436."""
    g1 = max(0.0, min_acc - acc)
    g2 = max(0.0, size_mb - max_size)
    g3 = max(0.0, fpr - max_fpr)
    return g1 + g2 + g3


# ---------------------------------------------------------------- FPR (a-5)
def confusion(y_true, y_pred, n_classes: int) -> np.ndarray:
    cm = np.zeros((n_classes, n_classes), dtype=np.int64)
    for t, p in zip(np.asarray(y_true).ravel(), np.asarray(y_pred).ravel()):
        if 0 <= t < n_classes and 0 <= p < n_classes:   # sklearn drops labels outside `labels`
            cm[int(t), int(p)] += 1
    return cm


def fpr_macro(cm: np.ndarray, mode: str = "all") -> float:
    """Macro false-positive rate from a confusion matrix.

    mode="all"      nsga_penalty.py:351-364: mean over all classes, 0.0 where FP+TN == 0.
    mode="filtered" sa_nsga_local.py:138-141: mean over classes with total-row_i > 0
                    (0.0 when none qualify).
    """
    total = int(cm.sum())
    vals = []
    for i in range(cm.shape[0]):
        fp = int(cm[:, i].sum()) - int(cm[i, i])
        denom = total - int(cm[i, :].sum())            # == FP + TN
        if denom > 0:
            vals.append(fp / denom)
        elif mode == "all":
            vals.append(0.0)
    if not vals:
        return 0.0
    return float(np.mean(vals))


# ----------------------------------------------------- model size (a-4)
FC_UNITS = {1: [64], 2: [128, 64], 3: [256, 128, 64], 4: [512, 256, 128, 64]}


def param_count(hp: dict, n_classes: int, variant: str = "A") -> int:
    """Keras ``model.count_params()`` as a closed form of the genotype.

    Variant A: nsga_penalty.py:250-332 (2-conv stem, 2-conv residual).  Variant B:
    sa_nsga_penalty.py:150-176 (1-conv stem, 1-conv residual).  BatchNormalization
    contributes 4 values per channel (gamma, beta, moving mean, moving var).
    """
    f, k = int(hp["filters"]), int(hp["kernel_size"])
    bn = 4 if hp["use_bn"] else 0
    total = 0

    def conv(cin, cout, ks, with_bn=True):
        return ks * ks * cin * cout + cout + (bn * cout if with_bn else 0)

    total += conv(1, f, k)
    if variant == "A":
        total += conv(f, f, k)
    for _ in range(int(hp["residual_blocks"])):
        total += conv(f, 2 * f, 1, with_bn=False)       # skip projection
        total += conv(f, 2 * f, k)
        if variant == "A":
            total += conv(2 * f, 2 * f, k)
        f *= 2
    width = f
    for units in FC_UNITS.get(int(hp["fc_layers"]), []):
        total += width * units + units
        width = units
    total += width * n_classes + n_classes
    return total


def model_size_mb(hp: dict, n_classes: int, variant: str = "A") -> float:
    """nsga_penalty.py:337-344: params * 4 / 1024**2."""
    return param_count(hp, n_classes, variant) * 4 / (1024 ** 2)


# -------------------------------------------------- infill selection (a-14)
def select_infill_points(pred: list[dict], k: int, eps: float = EPSILON):
    """sa_nsga_penalty.py:472-518 / sa_nsga_local.py:303-349."""
    feas = [(i, r) for i, r in enumerate(pred) if r["CV"] < eps]
    infeas = [(i, r) for i, r in enumerate(pred) if not r["CV"] < eps]
    order: list[int] = []
    if feas:
        objs = np.array([r["objs"] for _, r in feas])
        lo = objs.min(axis=0)
        span = objs.max(axis=0) - lo
        span[span < eps] = 1.0
        score = ((objs - lo) / span).sum(axis=1)
        order += [i for i, _ in sorted(zip([i for i, _ in feas], score), key=lambda t: t[1])]
    if infeas:
        order += [i for i, _ in sorted(infeas, key=lambda t: t[1]["CV"])]
    chosen = order[:k]
    return chosen, [pred[i]["hparams"] for i in chosen]


# -------------------------------------------------- front quality (a-16)
def dominates_min(a, b) -> bool:
    """compare.ipynb cell 0, 'dominates_min'."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return bool(np.all(a <= b) and np.any(a < b))


def nondominated_mask(points: np.ndarray) -> np.ndarray:
    """Union 'true front' filter, compare.ipynb cell 0 section 6."""
    pts = np.asarray(points, dtype=np.float64)
    keep = np.ones(len(pts), dtype=bool)
    for i in range(len(pts)):
        for j in range(len(pts)):
            if i != j and dominates_min(pts[j], pts[i]):
                keep[i] = False
                break
    return keep


def _pairwise(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    d = a[:, None, :] - b[None, :, :]
    return np.sqrt((d * d).sum(axis=2))


def generational_distance(front, true_front) -> float:
    """sqrt(mean(min_j d_ij^2)), compare.ipynb 'generational_distance'."""
    d = _pairwise(np.asarray(front, float), np.asarray(true_front, float)).min(axis=1)
    return float(np.sqrt(np.mean(d ** 2)))


def inverted_gd(front, true_front) -> float:
    return generational_distance(true_front, front)


def spread_metric(front, true_front) -> float:
    """compare.ipynb 'spread_metric' (distance-to-true-front variant of Deb's Delta)."""
    front = np.asarray(front, float)
    true_front = np.asarray(true_front, float)
    if len(front) < 2:
        return float("nan")
    d = _pairwise(front, true_front).min(axis=1)
    mean = d.mean()
    df = _pairwise(front, true_front.min(axis=0)[None, :]).min()
    dl = _pairwise(front, true_front.max(axis=0)[None, :]).min()
    num = df + dl + np.abs(d - mean).sum()
    den = df + dl + (len(front) - 1) * mean
    return float(num / den) if den != 0 else float("nan")


def coverage_metric(a, b) -> float:
    """C(A,B): fraction of B dominated by some point of A (compare.ipynb)."""
    b = np.asarray(b, float)
    if len(b) == 0:
        return 0
    hit = sum(1 for y in b if any(dominates_min(x, y) for x in np.asarray(a, float)))
    return hit / len(b)


def tchebycheff_scores(points: np.ndarray) -> np.ndarray:
    """Equal-weight Tchebycheff distance max_j w_j*|f_j - z*_j| to the ideal point
    z* = column minima ('Tchebycheff s_rank.ipynb', tchebycheff_score); post-hoc
    decision aid, out of the hot path, kept for completeness."""
    pts = np.asarray(points, float)
    w = np.full(pts.shape[1], 1.0 / pts.shape[1])
    return (w * np.abs(pts - pts.min(axis=0))).max(axis=1)
