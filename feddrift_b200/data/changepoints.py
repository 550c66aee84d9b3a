"""Change-point matrices: which concept each client sees at each time step.

Parity: ``data/changepoints/*.cp`` (int matrix ``[T+1, clients]``) and the
``rand`` generator in ``fedml_api/data_preprocessing/sea/data_loader.py:49-64``
(one random change point per client, 0 → 1, optionally ``drift_together``).
Additions for scale-out: named matrices are tiled along the client axis when
more than 10 clients are requested, and extended along time by repeating the
last row (a stationary tail).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from ._cp_tables import NAMED


def named(name: str) -> np.ndarray:
    rows = NAMED[name]
    return np.array([[int(ch) for ch in r] for r in rows], dtype=np.int64)


def random_single(train_iteration: int, num_client: int, drift_together: bool = False, stretch: int = 1,
                  rng: Optional[np.random.RandomState] = None) -> np.ndarray:
    rng = rng if rng is not None else np.random
    hi = max(2, train_iteration // stretch)
    if drift_together:
        cps = [rng.randint(1, hi)] * num_client
    else:
        cps = [rng.randint(1, hi) for _ in range(num_client)]
    mat = np.zeros((train_iteration // stretch + 1, num_client), dtype=np.int64)
    for c, t in enumerate(cps):
        mat[t:, c] = 1
    return mat


def load(spec, train_iteration: int, num_client: int, drift_together: bool = False, stretch: int = 1,
         rng=None) -> np.ndarray:
    """``spec``: matrix | 'rand' | name in NAMED | path to a .cp text file.  Returns ``[rows, num_client]``
    with at least ``train_iteration // stretch + 1`` rows."""
    if isinstance(spec, np.ndarray):
        mat = spec.astype(np.int64)
    elif spec in (None, "", "rand"):
        mat = random_single(train_iteration, num_client, drift_together, stretch, rng)
    elif spec in NAMED:
        mat = named(spec)
    elif os.path.exists(str(spec)):
        mat = np.atleast_2d(np.loadtxt(spec, dtype=np.int64))
    else:
        raise KeyError(f"unknown change-point spec {spec!r}")
    need_rows = train_iteration // stretch + 1
    if mat.shape[0] < need_rows:
        mat = np.concatenate([mat, np.repeat(mat[-1:], need_rows - mat.shape[0], axis=0)], axis=0)
    if mat.shape[1] < num_client:  # tile clients (client c behaves like c mod 10)
        reps = -(-num_client // mat.shape[1])
        mat = np.tile(mat, (1, reps))
    return mat[:, :num_client]


def concept_at(mat: np.ndarray, it: int, client: int, stretch: int = 1) -> int:
    return int(mat[it // stretch][client])


def save(path: str, mat: np.ndarray) -> None:
    np.savetxt(path, mat, fmt="%u")
