"""fMoW (WILDS) drift partitions.

Parity: ``fedml_api/data_preprocessing/fmow/data_loader.py:8-101`` — the reference keeps, per (client, time step), a CSV of
WILDS-fMoW sample indices (``data/fmow/partitions/<A..F>/client_<c>_iter_<t>.csv``, one index per line) and wraps them in
``torch.utils.data.Dataset``s over ``wilds.get_dataset("fmow")`` with 8 persistent DataLoader workers per (client, step).

Here the same index files are read into a dense, device-resident :class:`~feddrift_b200.data.drift.DriftData`
(``X[T+1, C, S, 3, R, R]`` with per-(step, client) valid lengths) that the engine consumes without worker processes:

* :func:`read_partition_indices` / :func:`load_partition_tables` — the CSV reader (single-value files, blank lines, empty
  files are handled like ``np.loadtxt`` + the reference's size-1 special case);
* :class:`WildsFmowSource` — a WILDS-free reader of the on-disk fMoW layout (``fmow_v1.1/rgb_metadata.csv`` +
  ``images/rgb_img_<i>.png``, 62 categories in sorted order — what ``wilds`` does internally);
* :class:`SyntheticFmowSource` — deterministic fMoW-shaped images when the 50 GB dataset is not on the box (no network
  here): image ``i`` of class ``y`` is a fixed random texture plus a class template, so learning curves are meaningful;
* :func:`fmow_drift_data` — partitions + source → ``DriftData``;
* :func:`load_partition_data_fmow` / :func:`load_all_data_fmow` — the reference's FedML 9-tuple / ``[client][iter]`` views
  (``class_num = 1000`` because the pretrained torchvision heads are kept, ``data_loader.py:50``).
"""
from __future__ import annotations

import csv
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .drift import DriftData, select_iterations

FMOW_CLASSES = 62


# ------------------------------------------------------------------------------------------------ partition CSVs
def read_partition_indices(path: str) -> np.ndarray:
    """One dataset index per line (``np.loadtxt(..., dtype=int, delimiter=',')`` semantics incl. the size-1 case)."""
    out: List[int] = []
    with open(path) as fh:
        for line in fh:
            for tok in line.replace(",", " ").split():
                out.append(int(float(tok)))
    return np.asarray(out, dtype=np.int64)


def partition_path(root: str, partition: str, client: int, iteration: int) -> str:
    return os.path.join(root, partition, f"client_{client}_iter_{iteration}.csv")


def load_partition_tables(root: str, partition: str, num_client: int, steps: int) -> List[List[np.ndarray]]:
    """``tables[t][c]`` = index array of client c at time step t (t = 0 … steps-1; step t+1 is step t's test set)."""
    return [[read_partition_indices(partition_path(root, partition, c, t)) for c in range(num_client)] for t in range(steps)]


# ------------------------------------------------------------------------------------------------ image sources
class SyntheticFmowSource:
    """Deterministic stand-in for the WILDS images: label = hash(index) mod 62, image = class template + index noise."""

    def __init__(self, resolution: int = 224, classes: int = FMOW_CLASSES, seed: int = 0):
        self.resolution, self.classes, self.seed = int(resolution), int(classes), int(seed)
        g = torch.Generator().manual_seed(seed)
        self._templates = torch.randn(classes, 3, 8, 8, generator=g)

    def label(self, idx: int) -> int:
        x = (int(idx) * 2654435761 + self.seed * 97) & 0xFFFFFFFF
        x ^= x >> 15
        return int(x % self.classes)

    def image(self, idx: int) -> torch.Tensor:
        y = self.label(idx)
        g = torch.Generator().manual_seed((int(idx) * 1000003 + self.seed) & 0x7FFFFFFF)
        r = self.resolution
        base = torch.nn.functional.interpolate(self._templates[y][None], size=(r, r), mode="nearest")[0]
        return (0.5 + 0.25 * base + 0.1 * torch.randn(3, r, r, generator=g)).clamp(0, 1)

    def __call__(self, idx: int) -> Tuple[torch.Tensor, int]:
        return self.image(idx), self.label(idx)


class WildsFmowSource:
    """Reads the WILDS fMoW v1.1 directory without the ``wilds`` package: labels from ``rgb_metadata.csv`` (category →
    index in sorted category order), pixels from ``images/rgb_img_<i>.png`` scaled to [0, 1] (``ToTensor``)."""

    def __init__(self, data_dir: str, resolution: int = 224):
        root = data_dir if os.path.exists(os.path.join(data_dir, "rgb_metadata.csv")) else os.path.join(data_dir, "fmow_v1.1")
        self.root, self.resolution = root, int(resolution)
        with open(os.path.join(root, "rgb_metadata.csv")) as fh:
            rows = list(csv.DictReader(fh))
        cats = sorted({r["category"] for r in rows})
        self.category_to_idx = {c: i for i, c in enumerate(cats)}
        self.labels = np.asarray([self.category_to_idx[r["category"]] for r in rows], dtype=np.int64)

    def label(self, idx: int) -> int:
        return int(self.labels[idx])

    def image(self, idx: int) -> torch.Tensor:
        from PIL import Image
        img = Image.open(os.path.join(self.root, "images", f"rgb_img_{int(idx)}.png")).convert("RGB")
        if img.size != (self.resolution, self.resolution):
            img = img.resize((self.resolution, self.resolution))
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255.0)

    def __call__(self, idx: int) -> Tuple[torch.Tensor, int]:
        return self.image(idx), self.label(idx)


def make_source(data_dir: Optional[str], resolution: int = 224):
    """The real dataset when it is on disk, the synthetic source otherwise."""
    if data_dir:
        for cand in (data_dir, os.path.join(data_dir, "fmow_v1.1")):
            if os.path.exists(os.path.join(cand, "rgb_metadata.csv")):
                return WildsFmowSource(cand, resolution)
    return SyntheticFmowSource(resolution)


# ------------------------------------------------------------------------------------------------ dense tensors
def fmow_drift_data(partitions_root: str, partition: str, train_iteration: int, num_client: int, source=None,
                    resolution: int = 224, max_samples: Optional[int] = None) -> DriftData:
    """Partitions + image source → dense ``DriftData`` (``T+1`` steps; ``nsamp[t, c]`` = file length, optionally capped)."""
    source = source if source is not None else SyntheticFmowSource(resolution)
    T1 = train_iteration + 1
    tables = load_partition_tables(partitions_root, partition, num_client, T1)
    S = max(1, max(len(a) for row in tables for a in row))
    if max_samples is not None:
        S = min(S, int(max_samples))
    X = torch.zeros(T1, num_client, S, 3, resolution, resolution, dtype=torch.float32)
    Y = torch.zeros(T1, num_client, S, dtype=torch.int64)
    ns = torch.zeros(T1, num_client, dtype=torch.int32)
    for t in range(T1):
        for c in range(num_client):
            idx = tables[t][c][:S]
            ns[t, c] = len(idx)
            for j, i in enumerate(idx):
                x, y = source(int(i))
                X[t, c, j], Y[t, c, j] = x, y
    cp = np.zeros((T1, num_client), dtype=np.int64)   # the concept ids are implicit in the index files
    return DriftData("fmow", X, Y, ns, cp, FMOW_CLASSES, 1, {"partition": partition, "indices": tables})


# ------------------------------------------------------------------------------------------------ FedML views
def _batches(x: torch.Tensor, y: torch.Tensor, batch_size: int, shuffle_rng=None):
    n = x.shape[0]
    if n == 0:
        return []
    perm = torch.from_numpy(shuffle_rng.permutation(n)) if shuffle_rng is not None else torch.arange(n)
    return [(x[perm[i:i + batch_size]], y[perm[i:i + batch_size]]) for i in range(0, n, batch_size)]


def load_all_data_fmow(data: DriftData, batch_size: int, current_train_iteration: int) -> List[List[list]]:
    """``[client][iter]`` → list of batches (``data_loader.py:8-12``)."""
    return [[data.batches(c, it, batch_size) for it in range(current_train_iteration + 1)] for c in range(data.client_num)]


def load_partition_data_fmow(data: DriftData, batch_size: int, current_train_iteration: int, retrain_data: str, rng=None):
    """The reference's 9-tuple (``data_loader.py:15-54``): training set = the iterations chosen by ``retrain_data``
    (``poisson`` = Poisson(1) bootstrap of the current step), test set = the NEXT time step, ``class_num = 1000``."""
    rng = rng if rng is not None else np.random.RandomState(0)
    t, C = current_train_iteration, data.client_num
    train_num = test_num = 0
    local_num: Dict[int, int] = {}
    train_local: Dict[int, list] = {}
    test_local: Dict[int, list] = {}
    gx, gy = [], []
    for c in range(C):
        iters = select_iterations(retrain_data, t, c)
        xs = [data.X[i, c, : int(data.nsamp[i, c])] for i in iters]
        ys = [data.Y[i, c, : int(data.nsamp[i, c])] for i in iters]
        x = torch.cat(xs) if xs else data.X[0, c, :0]
        y = torch.cat(ys) if ys else data.Y[0, c, :0]
        if retrain_data.startswith("poisson") and x.shape[0] > 1:
            w = rng.poisson(1.0, size=x.shape[0]).astype(np.float64)
            if w.sum() > 0:
                pick = torch.from_numpy(rng.choice(x.shape[0], size=x.shape[0], replace=True, p=w / w.sum()))
                x, y = x[pick], y[pick]
        nt = int(data.nsamp[t + 1, c]) if t + 1 < data.steps else 0
        xt, yt = (data.X[t + 1, c, :nt], data.Y[t + 1, c, :nt]) if nt else (data.X[0, c, :0], data.Y[0, c, :0])
        train_num += x.shape[0]
        test_num += nt
        local_num[c] = int(x.shape[0])
        train_local[c] = _batches(x, y, batch_size, rng)
        test_local[c] = _batches(xt, yt, batch_size, rng)
        n_now = int(data.nsamp[t, c])
        gx.append(data.X[t, c, :n_now]); gy.append(data.Y[t, c, :n_now])
    train_global = _batches(torch.cat(gx), torch.cat(gy), batch_size, rng)   # win-1 data of every client
    return C, train_num, test_num, train_global, None, local_num, train_local, test_local, 1000
