"""Candidate sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

The population shards naturally (nsga_penalty.py:426-441 has no cross-individual state), so the only
exchange is the all-gather of the P x 6 fp64 result rows after each evaluation batch; the deterministic
NDS / crowding then runs redundantly on every rank.
"""
from __future__ import annotations


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
