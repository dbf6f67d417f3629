"""Exact hypervolume (minimisation) in 2-D / 3-D, fp64.  PARITY UNPINNED.

TEST INFRASTRUCTURE ONLY.  The reference computes HV with the un-vendored
pygmo==2.19.4 (requirements.txt:129): ``pg.hypervolume(points).compute(r)`` with
``r = column max + 1e-3`` (compare.ipynb cell 0, sections 4-5).  pygmo is not
installable here and the .xlsx fronts the notebook read are not in the repo, so
this restates the published definition (Lebesgue measure of the union of boxes
[p, r]) with a slab sweep, and is validated against inclusion-exclusion and
hand-computed cases in tests/test_oracle_hv.py.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np


def _area2d(xy: np.ndarray, rx: float, ry: float) -> float:
    """Area of the union of [x_i, rx] x [y_i, ry]."""
    pts = [(x, y) for x, y in xy if x < rx and y < ry]
    if not pts:
        return 0.0
    pts.sort()
    area = 0.0
    best_y = ry
    for x, y in pts:                     # ascending x: each point adds the strip below best_y
        if y < best_y:
            area += (rx - x) * (best_y - y)
            best_y = y
    return area


def hypervolume(points, ref) -> float:
    pts = np.atleast_2d(np.asarray(points, np.float64))
    ref = np.asarray(ref, np.float64)
    if pts.size == 0:
        return 0.0
    m = pts.shape[1]
    if m == 2:
        return _area2d(pts, ref[0], ref[1])
    if m != 3:
        raise ValueError("only 2 or 3 objectives")
    pts = pts[np.all(pts < ref, axis=1)]
    if len(pts) == 0:
        return 0.0
    order = np.argsort(pts[:, 2], kind="stable")
    pts = pts[order]
    vol = 0.0
    for k in range(len(pts)):
        z_lo = pts[k, 2]
        z_hi = pts[k + 1, 2] if k + 1 < len(pts) else ref[2]
        if z_hi > z_lo:
            vol += _area2d(pts[: k + 1, :2], ref[0], ref[1]) * (z_hi - z_lo)
    return float(vol)


def hypervolume_inclusion_exclusion(points, ref) -> float:
    """Independent O(2^n) check for tiny fronts."""
    pts = np.atleast_2d(np.asarray(points, np.float64))
    ref = np.asarray(ref, np.float64)
    pts = pts[np.all(pts < ref, axis=1)]
    total = 0.0
    for r in range(1, len(pts) + 1):
        for combo in combinations(range(len(pts)), r):
            corner = pts[list(combo)].max(axis=0)
            total += (-1) ** (r + 1) * float(np.prod(ref - corner))
    return total


def reference_point(all_points, eps: float = 1e-3) -> np.ndarray:
    """compare.ipynb section 4: per-objective max over the union of fronts + 1e-3."""
    return np.asarray(all_points, np.float64).max(axis=0) + eps
