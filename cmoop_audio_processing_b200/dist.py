"""Candidate sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

The population shards naturally (nsga_penalty.py:426-441 has no cross-individual state), so the only
exchange is the all-gather of the P x 6 fp64 result rows after each evaluation batch; the deterministic
NDS / crowding then runs redundantly on every rank.

Load balance: longest-processing-time-first on a MEASURED per-genotype cost.  The analytic forward MACs
spread 100x over the genotype space while the measured device time of a candidate spreads 8x (narrow
networks are bound by activation traffic and fixed per-launch work, not by the tensor pipe), so a
MAC-proportional split starves the ranks that receive the "cheap" candidates' real cost (round-1 SCALE:
0.78 strong-scaling efficiency at 8 GPUs).  ``cost_table.json`` holds device milliseconds per candidate
and epoch for every (variant, filters, kernel_size, residual_blocks, use_bn), written by
``tools/calibrate_cost.py`` on a B200; other feature-map shapes / gene values fall back to a two-term model
(activation elements + MACs) fitted to the same table.
"""
from __future__ import annotations

import json
import os
from functools import lru_cache

_TABLE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cost_table.json")


def assign_lpt(costs, n_ranks: int) -> list[int]:
    """Longest-processing-time-first assignment: returns owner rank per item; ties broken by index so every
    rank computes the identical map."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * n_ranks
    owner = [0] * len(costs)
    for i in order:
        r = min(range(n_ranks), key=lambda j: (load[j], j))
        owner[i] = r
        load[r] += costs[i]
    return owner


def _features(hp, height, width, classes, variant):
    """(activation elements per sample, forward MACs per sample) of a genotype: the two cost drivers."""
    from .problem import forward_macs
    f = int(hp["filters"])
    h, w = height, width
    act = h * w * f * (2 if variant == "A" else 1)
    h, w = (h + 1) // 2, (w + 1) // 2
    for _ in range(int(hp["residual_blocks"])):
        act += h * w * 2 * f * (2 if variant == "A" else 1) + ((h + 1) // 2) * ((w + 1) // 2) * 2 * f
        f *= 2
        h, w = (h + 1) // 2, (w + 1) // 2
    return float(act), float(forward_macs(hp, height, width, classes, variant))


@lru_cache(maxsize=1)
def _table():
    try:
        with open(_TABLE_PATH) as fh:
            rec = json.load(fh)
    except OSError:
        return None
    table = dict(rec["table"])
    for key, val in list(table.items()):                 # a BN-free network is never dearer than its BN twin (timer noise)
        if key.endswith(":0"):
            twin = table.get(key[:-1] + "1")
            if twin is not None and val > twin:
                table[key] = 0.87 * twin
    rec["table"] = table
    # two-term fallback model per variant: ms = a * activation elements + b * MACs (least squares on the table)
    import numpy as np
    fits = {}
    h, w = rec["shape"]
    for variant in ("A", "B"):
        rows, ys = [], []
        for key, val in table.items():
            v, f, k, r, bn = key.split(":")
            if v != variant:
                continue
            hp = dict(filters=int(f), kernel_size=int(k), residual_blocks=int(r), fc_layers=1, use_bn=bool(int(bn)))
            act, macs = _features(hp, h, w, rec["classes"], variant)
            rows.append([act * (1.3 if int(bn) else 1.0), macs])
            ys.append(val)
        if rows:
            coef, *_ = np.linalg.lstsq(np.asarray(rows), np.asarray(ys), rcond=None)
            fits[variant] = [max(float(c), 0.0) for c in coef]
    rec["fits"] = fits
    return rec


def candidate_cost(hp, height: int, width: int, classes: int, variant: str = "B") -> float:
    """Relative device time of training one candidate for one epoch (arbitrary but consistent units)."""
    rec = _table()
    act, macs = _features(hp, height, width, classes, variant)
    if rec is None:
        return act + macs / 40.0                                     # no table shipped: activation-dominated guess
    if [height, width] == list(rec["shape"]):
        key = f"{variant}:{int(hp['filters'])}:{int(hp['kernel_size'])}:{int(hp['residual_blocks'])}:{int(bool(hp['use_bn']))}"
        if key in rec["table"]:
            return float(rec["table"][key])
    a, b = rec["fits"].get(variant, (1e-6, 2.5e-8))
    return a * act * (1.3 if hp.get("use_bn") else 1.0) + b * macs
