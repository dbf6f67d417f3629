"""Drop-in selection-side functions, same names/records as the reference, CUDA underneath.

* ``fast_non_dominated_sort(results, lam)``  nsga_penalty.py:467-501, sa_nsga_penalty.py:392-419
* ``crowding_distance(front, results)``      nsga_penalty.py:504-524, sa_nsga_penalty.py:421-442
* ``environmental_selection``                 the (mu+lambda) truncation of nsga_penalty.py:676-690
  done from ONE kernel launch (ranks, reference-ordered fronts and crowding together).

``dominates`` / ``tournament_selection`` / ``crossover`` / ``mutate`` stay host-side
Python on purpose: they define the ``random`` stream the drop-in must not perturb
(SURVEY.md section 8a-17).  Records are the reference's ``{'hparams','objs','CV'}`` dicts.
"""
from __future__ import annotations

import random
from copy import deepcopy

import numpy as np

from . import _lib

EPSILON = 1e-6            # nsga_penalty.py:203
CROWD_RANGE_GT = 0        # skip objective unless (max-min) >  EPSILON   (sa_nsga_penalty.py:437)
CROWD_RANGE_LT = 1        # skip objective when   (max-min) <  EPSILON   (nsga_penalty.py:518)


def _pack(results):
    n = len(results)
    m = len(results[0]["objs"]) if n else 0
    objs = np.ascontiguousarray([r["objs"] for r in results], dtype=np.float64).reshape(n, m)
    cv = np.ascontiguousarray([r["CV"] for r in results], dtype=np.float64)
    return objs, cv


def nds_crowding_arrays(objs: np.ndarray, cv: np.ndarray | None, lam: float, *, eps: float = EPSILON,
                        crowd_mode: int = CROWD_RANGE_GT, want_crowd: bool = True):
    """Array-level entry: objs [n,m] or [batch,n,m] fp64.  Returns dict of numpy arrays
    (rank, order, front_offsets, n_fronts, crowd)."""
    lib = _lib.load()
    objs = np.ascontiguousarray(objs, dtype=np.float64)
    batched = objs.ndim == 3
    if not batched:
        objs = objs[None]
    batch, n, m = objs.shape
    if cv is not None:
        cv = np.ascontiguousarray(cv, dtype=np.float64).reshape(batch, n)
    rank = np.empty((batch, n), np.int32)
    order = np.empty((batch, n), np.int32)
    foff = np.empty((batch, n + 1), np.int32)
    nf = np.empty((batch,), np.int32)
    crowd = np.empty((batch, n), np.float64) if want_crowd else None
    _lib.check(lib.cmoop_nds_crowding_host(_lib.ptr(objs), _lib.ptr(cv), n, m, batch, float(lam), float(eps),
                                           int(crowd_mode), _lib.ptr(rank), _lib.ptr(order), _lib.ptr(foff),
                                           _lib.ptr(nf), _lib.ptr(crowd)), "cmoop_nds_crowding_host")
    out = dict(rank=rank, order=order, front_offsets=foff, n_fronts=nf, crowd=crowd)
    if not batched:
        out = {k: (v[0] if v is not None else None) for k, v in out.items()}
    return out


def _fronts_from(order, foff, nf):
    return [order[foff[i]:foff[i + 1]].tolist() for i in range(int(nf))]


def fast_non_dominated_sort(results, lam):
    """list[list[int]] of non-empty fronts in the reference's order; [] for empty input."""
    if len(results) == 0:
        return []
    objs, cv = _pack(results)
    r = nds_crowding_arrays(objs, cv, lam, want_crowd=False)
    return _fronts_from(r["order"], r["front_offsets"], r["n_fronts"])


def crowding_distance(front, results, *, eps: float = EPSILON, crowd_mode: int = CROWD_RANGE_GT):
    """dict index -> crowding distance for the given index list ({} for an empty front)."""
    if not front:
        return {}
    lib = _lib.load()
    objs, _ = _pack(results)
    idx = np.ascontiguousarray(front, dtype=np.int32)
    out = np.empty(len(idx), np.float64)
    _lib.check(lib.cmoop_crowding_distance_host(_lib.ptr(objs), objs.shape[0], objs.shape[1], _lib.ptr(idx), len(idx),
                                                float(eps), int(crowd_mode), _lib.ptr(out)),
               "cmoop_crowding_distance_host")
    return {int(i): float(d) for i, d in zip(idx, out)}


def environmental_selection(combined, lam, pop_size, *, eps: float = EPSILON, crowd_mode: int = CROWD_RANGE_GT):
    """Indices into ``combined`` kept for the next generation (nsga_penalty.py:676-690):
    whole fronts while they fit, the overflowing front by descending crowding distance
    (stable, i.e. ties keep the front's own order)."""
    if len(combined) == 0:
        return []
    objs, cv = _pack(combined)
    r = nds_crowding_arrays(objs, cv, lam, eps=eps, crowd_mode=crowd_mode)
    keep: list[int] = []
    for front in _fronts_from(r["order"], r["front_offsets"], r["n_fronts"]):
        if len(keep) + len(front) <= pop_size:
            keep.extend(front)
        else:
            room = pop_size - len(keep)
            ranked = sorted(front, key=lambda i: r["crowd"][i], reverse=True)
            keep.extend(ranked[:room])
            break
    return keep


# ---------------------------------------------------------------- host-side operators (unchanged semantics)
def get_lambda(gen, max_gen, lam_initial=1.0, lam_final=50.0):
    """sa_nsga_penalty.py:130-132 (guarded form)."""
    frac = gen / float(max_gen - 1) if max_gen > 1 else 1.0
    return lam_initial + frac * (lam_final - lam_initial)


def dominates(a, b, lam):
    """Penalised dominance of two records (nsga_penalty.py:448-461); scalar host code used
    by tournament_selection only."""
    pa = [f + lam * a["CV"] for f in a["objs"]]
    pb = [f + lam * b["CV"] for f in b["objs"]]
    return all(x <= y for x, y in zip(pa, pb)) and any(x < y for x, y in zip(pa, pb))


def tournament_selection(results, lam, k=2):
    """nsga_penalty.py:528-538; consumes random.sample exactly like the reference."""
    idxs = random.sample(range(len(results)), k)
    best = idxs[0]
    for idx in idxs[1:]:
        if dominates(results[idx], results[best], lam):
            best = idx
    return best


HPARAM_SPACE = {
    "filters": [16, 32, 64], "kernel_size": [3, 5], "use_bn": [True, False],
    "residual_blocks": [1, 2, 3], "fc_layers": [1, 2, 3, 4], "use_dropout": [True, False],
}


def initialize_population(pop_size):
    """nsga_penalty.py:402-415 (one random.choice per gene, gene order of the dict)."""
    return [{k: random.choice(v) for k, v in HPARAM_SPACE.items()} for _ in range(pop_size)]


def crossover(p1, p2):
    """Uniform gene swap with probability 0.5 per gene (nsga_penalty.py:541-577)."""
    c1, c2 = deepcopy(p1), deepcopy(p2)
    for key in HPARAM_SPACE:
        if random.random() < 0.5:
            c1[key], c2[key] = p2[key], p1[key]
    return c1, c2


def mutate(individual, mutation_prob=0.2):
    """Per-gene mutation (nsga_penalty.py:579-603): booleans flip, others re-drawn."""
    ind = deepcopy(individual)
    for key, options in HPARAM_SPACE.items():
        if random.random() < mutation_prob:
            if isinstance(options[0], bool):
                ind[key] = not ind[key]
            else:
                ind[key] = random.choice(options)
    return ind
