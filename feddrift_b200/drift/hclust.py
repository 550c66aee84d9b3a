"""Tiny agglomerative clustering used by FedDrift (K5) and CFL (K6).

Replaces ``scipy.cluster.hierarchy.linkage`` + ``fcluster(criterion='distance')``
(``FedAvgEnsDataLoader.py:947-951``) and sklearn ``AgglomerativeClustering(linkage='complete')`` on
``-S`` (``:1245-1249``) for the L ≤ #models / n ≤ #clients sizes that occur (so O(L³) is free and the
same code runs on the host next to the device-computed distance matrix).  Complete and average
linkage are monotone, hence the flat clustering at threshold t equals "apply every merge whose
height ≤ t".  Ties are broken towards the lexicographically smallest pair.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def _agglomerate(D: np.ndarray, method: str, stop_height=None, stop_clusters=None):
    """Generic Lance–Williams loop.  Returns (members per live cluster id, merge list)."""
    n = D.shape[0]
    dist = np.array(D, dtype=np.float64, copy=True)
    np.fill_diagonal(dist, np.inf)
    members = {i: [i] for i in range(n)}
    node_id = {i: i for i in range(n)}  # sklearn/scipy style node numbering
    alive = list(range(n))
    merges: List[Tuple[int, int, float]] = []
    while len(alive) > 1:
        if stop_clusters is not None and len(alive) <= stop_clusters:
            break
        best, bi, bj = np.inf, -1, -1
        for a_idx, a in enumerate(alive):
            for b in alive[a_idx + 1:]:
                if dist[a, b] < best:
                    best, bi, bj = dist[a, b], a, b
        if stop_height is not None and best > stop_height:
            break
        na, nb = len(members[bi]), len(members[bj])
        for k in alive:
            if k in (bi, bj):
                continue
            if method == "complete":
                d = max(dist[bi, k], dist[bj, k])
            elif method == "average":
                d = (na * dist[bi, k] + nb * dist[bj, k]) / (na + nb)
            elif method == "single":
                d = min(dist[bi, k], dist[bj, k])
            else:
                raise ValueError(method)
            dist[bi, k] = dist[k, bi] = d
        members[bi] = members[bi] + members[bj]
        node_id[bi] = n + len(merges)
        merges.append((bi, bj, float(best)))
        del members[bj]
        alive.remove(bj)
    return members, node_id, merges


def linkage_fcluster(D: np.ndarray, method: str = "complete", t: float = 0.0) -> np.ndarray:
    """Flat cluster labels (1-based, ordered by smallest member) with cophenetic distance ≤ t."""
    n = D.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    members, _, _ = _agglomerate(np.asarray(D), method, stop_height=t)
    labels = np.zeros(n, dtype=np.int64)
    for lab, (_, mem) in enumerate(sorted(members.items(), key=lambda kv: min(kv[1])), start=1):
        labels[mem] = lab
    return labels


def complete_linkage_bipartition(S: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Split indices into two groups by complete-linkage clustering of the distance ``-S`` down to two
    clusters.  Group 0 is the cluster with the larger dendrogram node id (sklearn ``_hc_cut`` order)."""
    n = S.shape[0]
    if n < 2:
        return np.arange(n), np.zeros(0, dtype=np.int64)
    members, node_id, _ = _agglomerate(-np.asarray(S, dtype=np.float64), "complete", stop_clusters=2)
    groups = sorted(members.keys(), key=lambda k: -node_id[k])
    return np.array(sorted(members[groups[0]])), np.array(sorted(members[groups[1]]))
