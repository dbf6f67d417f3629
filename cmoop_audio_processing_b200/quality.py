"""Front-quality metrics of compare.ipynb (cell 0, sections 4-9) on the GPU.

``hypervolume`` replaces ``pg.hypervolume(points).compute(ref)``; the other functions keep
the notebook's names: ``dominates_min``-based true-front filter, ``generational_distance``,
``inverted_gd``, ``spread_metric``, ``coverage_metric``.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _pts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float64)
    if a.ndim != 2:
        raise ValueError("points must be 2-D (n, m)")
    return a


def reference_point(all_points, eps: float = 1e-3) -> np.ndarray:
    """compare.ipynb section 4: per-objective max over the union of fronts + 1e-3."""
    return _pts(all_points).max(axis=0) + eps


def hypervolume(points, ref) -> float:
    pts = _pts(points)
    ref = np.ascontiguousarray(ref, np.float64)
    if ref.shape != (pts.shape[1],):
        raise ValueError("ref must have one entry per objective")
    out = np.zeros(1, np.float64)
    lib = _lib.load()
    _lib.check(lib.cmoop_hypervolume_host(_lib.ptr(pts), pts.shape[0], pts.shape[1], _lib.ptr(ref), _lib.ptr(out)),
               "cmoop_hypervolume_host")
    return float(out[0])


def nondominated_mask(points) -> np.ndarray:
    pts = _pts(points)
    mask = np.zeros(pts.shape[0], np.uint8)
    lib = _lib.load()
    _lib.check(lib.cmoop_nondominated_mask_host(_lib.ptr(pts), pts.shape[0], pts.shape[1], _lib.ptr(mask)),
               "cmoop_nondominated_mask_host")
    return mask.astype(bool)


def _metrics(front, true_front) -> np.ndarray:
    f, t = _pts(front), _pts(true_front)
    out = np.zeros(3, np.float64)
    lib = _lib.load()
    _lib.check(lib.cmoop_front_metrics_host(_lib.ptr(f), f.shape[0], _lib.ptr(t), t.shape[0], f.shape[1],
                                            _lib.ptr(out)), "cmoop_front_metrics_host")
    return out


def generational_distance(obtained_front, true_front) -> float:
    return float(_metrics(obtained_front, true_front)[0])


def inverted_gd(obtained_front, true_front) -> float:
    return float(_metrics(obtained_front, true_front)[1])


def spread_metric(front, true_front) -> float:
    if len(front) < 2:
        return float("nan")
    return float(_metrics(front, true_front)[2])


def front_metrics(front, true_front) -> dict:
    gd, igd, spread = _metrics(front, true_front)
    return {"gd": float(gd), "igd": float(igd), "spread": float(spread)}


def coverage_metric(a, b) -> float:
    a, b = _pts(a), _pts(b)
    if len(b) == 0:
        return 0
    out = np.zeros(1, np.float64)
    lib = _lib.load()
    _lib.check(lib.cmoop_coverage_host(_lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], a.shape[1], _lib.ptr(out)),
               "cmoop_coverage_host")
    return float(out[0])
