"""Concept-drift datasets (SEA / sine / circle / MNIST-label-swap / fMoW-shaped) held as dense
device-friendly tensors ``X[T+1, C, S, ...]``, ``Y[T+1, C, S]``.

Parity (behaviour, not layout): ``fedml_api/data_preprocessing/{sea,sine,circle}/data_loader.py``,
``MNIST/data_loader_cont.py:50-214``, ``fmow/data_loader.py`` and the retrain-window selector
``common/retrain.py:7-91``.  The reference round-trips every (client, time-step) sample set through
CSV files and re-reads ALL clients on EVERY rank each time step (``main_fedavg.py:313-315``); here the
whole experiment is generated once into one tensor that is uploaded to HBM once (SEA-4 / 10 clients /
11 steps is 53 KB; MNIST 64 clients × 11 × 500 × 784 fp32 is 1.1 GB — trivial against 180 GB).
Time step t+1 is time step t's test set, so T+1 steps are generated (``retrain.py:78-83``).

CSV import/export (``client_{c}_iter_{t}.csv``) is kept for interoperability with reference data dirs.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import changepoints as cp

SEA_THRESHOLDS = (8.0, 9.0, 7.0, 9.0)  # fitted to the shipped concept{1..4}.csv
DEFAULT_DELTAS = {"sea": 0.04, "sine": 0.20, "circle": 0.10, "MNIST": 0.10}
FEATURE_SHAPE = {"sea": (3,), "sine": (2,), "circle": (2,), "MNIST": (784,), "fmow": (3, 224, 224), "cifar10": (3, 32, 32),
                 "shakespeare": (80,)}
CLASS_NUM = {"sea": 2, "sine": 2, "circle": 2, "MNIST": 10, "fmow": 62, "cifar10": 10, "shakespeare": 90}


# ----------------------------------------------------------------------------- concept samplers
def _sea(rng, n, concept, label_noise=0.10, thresholds=SEA_THRESHOLDS):
    x = rng.uniform(0.0, 10.0, size=(n, 3))
    y = (x[:, 1] + x[:, 2] > thresholds[concept % len(thresholds)]).astype(np.int64)
    if label_noise > 0:  # the shipped SEA pools carry ~10 % label noise by construction
        flip = rng.rand(n) < label_noise
        y = np.where(flip, 1 - y, y)
    return x, y


def _sine(rng, n, concept):
    x = rng.rand(n, 2)
    below = x[:, 1] <= np.sin(x[:, 0])
    y = below.astype(np.int64) if concept == 0 else (~below).astype(np.int64)
    return x, y


def _circle(rng, n, concept):
    x = rng.rand(n, 2)
    cx, cy, r = (0.2, 0.5, 0.15) if concept == 0 else (0.6, 0.5, 0.25)
    y = (((x[:, 0] - cx) ** 2 + (x[:, 1] - cy) ** 2 - r * r) > 0).astype(np.int64)
    return x, y


_SWAPS = {1: (1, 2), 2: (3, 4), 3: (5, 6)}


class DigitPool:
    """MNIST-shaped pool.  Real LEAF MNIST json is used when ``root`` has it; otherwise a deterministic
    synthetic pool (class prototypes + noise, 784-d in [0,1]) of the same shape — there is no network here."""

    def __init__(self, root: Optional[str] = None, size: int = 20000, seed: int = 100):
        self.used = 0
        X = Y = None
        if root and os.path.isdir(os.path.join(root, "train")):
            try:
                X, Y = self._read_leaf(os.path.join(root, "train"))
            except Exception:
                X = Y = None
        if X is None:
            rng = np.random.RandomState(seed)
            protos = rng.rand(10, 784) < 0.18  # sparse binary "strokes" per class
            Y = rng.randint(0, 10, size=size)
            X = np.clip(protos[Y].astype(np.float32) * 0.9 + rng.randn(size, 784).astype(np.float32) * 0.25, 0, 1)
        perm = np.random.RandomState(seed).permutation(len(X))
        self.X, self.Y = np.asarray(X, dtype=np.float32)[perm], np.asarray(Y, dtype=np.int64)[perm]

    @staticmethod
    def _read_leaf(train_dir):
        xs, ys = [], []
        for f in sorted(os.listdir(train_dir)):
            if f.endswith(".json"):
                with open(os.path.join(train_dir, f)) as fh:
                    blob = json.load(fh)
                for u in blob["users"]:
                    xs.extend(blob["user_data"][u]["x"])
                    ys.extend(blob["user_data"][u]["y"])
        return np.asarray(xs, dtype=np.float32), np.asarray(ys, dtype=np.int64)

    def take(self, n: int, concept: int, mode: str = "swap"):
        if self.used + n >= len(self.X):  # wrap around like the reference (data_loader_cont.py:182-186)
            self.used = 0
        x = self.X[self.used:self.used + n].copy()
        y = self.Y[self.used:self.used + n].copy()
        self.used += n
        if concept != 0:
            if mode == "rotate":  # the variant BASELINE.json names; reference has it commented out
                x = np.rot90(x.reshape(-1, 28, 28), k=concept, axes=(1, 2)).reshape(-1, 784).copy()
            elif concept in _SWAPS:
                a, b = _SWAPS[concept]
                ya, yb = y == a, y == b
                y[ya], y[yb] = b, a
        return x, y


_CHAR_T = {}


def _char_concept(rng, n, concept, seq_len=80, vocab=90):
    """Character-LM concept: a concept is a (sparse) bigram transition table; x = 80 char ids, y = the next char."""
    if concept not in _CHAR_T:
        r = np.random.RandomState(1000 + concept)
        _CHAR_T[concept] = r.dirichlet(np.full(vocab - 4, 0.05), size=vocab)
    T = _CHAR_T[concept]
    cdf = np.cumsum(T, axis=1)
    out = np.zeros((n, seq_len + 1), dtype=np.int64)
    s = rng.randint(4, vocab, size=n)
    for i in range(seq_len + 1):
        u = rng.rand(n)
        s = 4 + (cdf[s] < u[:, None]).sum(1).clip(max=vocab - 5)
        out[:, i] = s
    return out[:, :seq_len], out[:, seq_len]


def _image_concept(rng, n, concept, shape, classes):
    """fMoW/CIFAR-shaped synthetic images: class-dependent low-frequency pattern; a concept permutes labels."""
    y = rng.randint(0, classes, size=n)
    c, h, w = shape
    base = rng.randn(n, c, 4, 4).astype(np.float32)
    img = np.repeat(np.repeat(base, h // 4, axis=2), w // 4, axis=3)
    img += (y[:, None, None, None] / classes - 0.5).astype(np.float32)
    y = (y + concept * 7) % classes
    return img, y


# ----------------------------------------------------------------------------- dataset container
@dataclass
class DriftData:
    name: str
    X: torch.Tensor          # [T1, C, S, *feat] float32
    Y: torch.Tensor          # [T1, C, S] int64
    nsamp: torch.Tensor      # [T1, C] int32 (valid prefix length of S)
    change_points: np.ndarray
    class_num: int
    stretch: int = 1
    meta: Dict = field(default_factory=dict)

    @property
    def steps(self) -> int:  # number of stored time steps (T+1)
        return self.X.shape[0]

    @property
    def client_num(self) -> int:
        return self.X.shape[1]

    @property
    def feature_num(self) -> int:
        return int(np.prod(self.X.shape[3:]))

    def to(self, device) -> "DriftData":
        return DriftData(self.name, self.X.to(device), self.Y.to(device), self.nsamp.to(device),
                         self.change_points, self.class_num, self.stretch, dict(self.meta))

    def pin(self) -> "DriftData":
        return DriftData(self.name, self.X.pin_memory(), self.Y.pin_memory(), self.nsamp.pin_memory(),
                         self.change_points, self.class_num, self.stretch, dict(self.meta))

    # FedML-style list-of-batches view ------------------------------------------------
    def batches(self, c: int, t: int, batch_size: int, index: Optional[torch.Tensor] = None):
        n = int(self.nsamp[t, c])
        x, y = self.X[t, c, :n], self.Y[t, c, :n]
        if index is not None:
            x, y = x[index], y[index]
            n = x.shape[0]
        return [(x[i:i + batch_size], y[i:i + batch_size]) for i in range(0, n, batch_size)]

    def concept(self, t: int, c: int) -> int:
        return cp.concept_at(self.change_points, t, c, self.stretch)

    # CSV interop ---------------------------------------------------------------------
    def to_csv_dir(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        cols = [f"f{i + 1}" for i in range(self.feature_num)] + ["label"]
        for t in range(self.steps):
            for c in range(self.client_num):
                n = int(self.nsamp[t, c])
                arr = np.concatenate([self.X[t, c, :n].reshape(n, -1).cpu().numpy(),
                                      self.Y[t, c, :n, None].cpu().numpy()], axis=1)
                np.savetxt(os.path.join(path, f"client_{c}_iter_{t}.csv"), arr, delimiter=",",
                           header=",".join(cols), comments="", fmt="%.9g")

    @staticmethod
    def from_csv_dir(path: str, name: str, num_client: int, steps: int, class_num: int,
                     change_points: Optional[np.ndarray] = None) -> "DriftData":
        frames = [[np.loadtxt(os.path.join(path, f"client_{c}_iter_{t}.csv"), delimiter=",", skiprows=1, ndmin=2)
                   for c in range(num_client)] for t in range(steps)]
        S = max(f.shape[0] for row in frames for f in row)
        F = frames[0][0].shape[1] - 1
        X = torch.zeros(steps, num_client, S, F)
        Y = torch.zeros(steps, num_client, S, dtype=torch.int64)
        ns = torch.zeros(steps, num_client, dtype=torch.int32)
        for t in range(steps):
            for c in range(num_client):
                f = frames[t][c]
                n = f.shape[0]
                X[t, c, :n] = torch.from_numpy(f[:, :-1]).float()
                Y[t, c, :n] = torch.from_numpy(f[:, -1]).long()
                ns[t, c] = n
        cpm = change_points if change_points is not None else np.zeros((steps, num_client), dtype=np.int64)
        return DriftData(name, X, Y, ns, cpm, class_num)


def generate_drift_data(dataset: str, train_iteration: int, num_client: int, sample_num: int,
                        noise_prob: float = 0.0, stretch: int = 1, change_points="rand",
                        drift_together: bool = False, seed: int = 0, data_dir: Optional[str] = None,
                        mnist_mode: str = "swap", image_shape: Optional[Tuple[int, int, int]] = None,
                        sea_label_noise: float = 0.10) -> DriftData:
    """Equivalent of ``prepare_data.py`` + ``generate_data_<dataset>``: T+1 steps × clients × ``sample_num``
    samples drawn from the concept given by the change-point matrix; labels flipped (or re-drawn for
    multi-class) with probability ``noise_prob``.  Seeds are fixed like the reference (``prepare_data.py:104``)."""
    rng = np.random.RandomState(seed)
    mat = cp.load(change_points, train_iteration, num_client, drift_together, stretch, rng)
    T1 = train_iteration + 1
    key = "MNIST" if dataset.lower() == "mnist" else dataset.lower()
    if key in ("fed_shakespeare",):
        key = "shakespeare"
    if key == "fmow":
        feat = tuple(image_shape) if image_shape else FEATURE_SHAPE["fmow"]
    else:
        feat = FEATURE_SHAPE[key]
    classes = CLASS_NUM[key]
    X = np.zeros((T1, num_client, sample_num) + feat, dtype=np.int64 if key == "shakespeare" else np.float32)
    Y = np.zeros((T1, num_client, sample_num), dtype=np.int64)
    pool = DigitPool(data_dir) if key == "MNIST" else None
    for it in range(T1):
        for c in range(num_client):
            k = cp.concept_at(mat, it, c, stretch)
            if key == "sea":
                x, y = _sea(rng, sample_num, k, sea_label_noise)
            elif key == "sine":
                x, y = _sine(rng, sample_num, k)
            elif key == "circle":
                x, y = _circle(rng, sample_num, k)
            elif key == "MNIST":
                x, y = pool.take(sample_num, k, mnist_mode)
            elif key == "shakespeare":
                x, y = _char_concept(rng, sample_num, k)
            else:
                x, y = _image_concept(rng, sample_num, k, feat, classes)
            if noise_prob > 0:
                flip = rng.rand(sample_num) < noise_prob
                if classes == 2:
                    y = np.where(flip, 1 - y, y)
                else:  # a different random class (data_loader_cont.py:40-48)
                    y = np.where(flip, (y + rng.randint(1, classes, size=sample_num)) % classes, y)
            perm = rng.permutation(sample_num)  # the reference shuffles when batching
            X[it, c], Y[it, c] = x[perm], y[perm]
    ns = torch.full((T1, num_client), sample_num, dtype=torch.int32)
    return DriftData(key, torch.from_numpy(X), torch.from_numpy(Y), ns, mat, classes, stretch,
                     {"seed": seed, "noise_prob": noise_prob, "sample_num": sample_num})


# ----------------------------------------------------------------------------- retrain-window selectors
def select_iterations(method: str, t_cur: int, client: Optional[int] = None) -> List[int]:
    """Iterations (with multiplicity) whose data forms the training set (``retrain.py:12-63``)."""
    if method == "all":
        return list(range(t_cur + 1))
    if method.startswith("win-"):
        w = int(method[4:])
        return list(range(max(0, t_cur - w + 1), t_cur + 1))
    if method.startswith("weight-"):
        lin = method[7:] == "linear"
        out: List[int] = []
        for it in range(t_cur + 1):
            out += [it] * ((it + 1) if lin else 2 ** it)
        return out
    if method.startswith("sel-"):
        body = method[4:]
        return [int(s) for s in body.split(",") if s != ""]
    if method.startswith("clientsel-"):
        table = json.loads(method[len("clientsel-"):])
        if client is None:
            raise ValueError("clientsel- needs a client index")
        return [int(i) for i in table[client]]
    if method.startswith("poisson"):
        return [t_cur]
    raise NameError(method)


def poisson_bootstrap_index(n: int, rng) -> Optional[torch.Tensor]:
    """Poisson(1) bootstrap weights -> a with-replacement resample of size n (``retrain.py:65-74``)."""
    w = rng.poisson(1.0, size=n).astype(np.float64)
    if w.sum() == 0:
        return None
    idx = rng.choice(n, size=n, replace=True, p=w / w.sum())
    return torch.from_numpy(idx)


def load_partition_data(data: DriftData, batch_size: int, t_cur: int, retrain: str, rng=None):
    """FedML tuple ``(client_num, train_num, test_num, train_global, test_global, local_num_dict,
    train_local_dict, test_local_dict, class_num)`` built from the dense tensors
    (parity: ``load_partition_data_sea`` ``sea/data_loader.py:102-141``)."""
    rng = rng if rng is not None else np.random
    C = data.client_num
    train_local, test_local, local_num = {}, {}, {}
    train_global, test_global = [], []
    train_num = test_num = 0
    for c in range(C):
        its = select_iterations(retrain, t_cur, c)
        batches, n_c = [], 0
        if its:
            xs, ys = [], []
            for it in its:
                n = int(data.nsamp[it, c])
                x, y = data.X[it, c, :n], data.Y[it, c, :n]
                if retrain.startswith("poisson"):
                    idx = poisson_bootstrap_index(n, rng)
                    if idx is not None:
                        x, y = x[idx], y[idx]
                xs.append(x)
                ys.append(y)
            x, y = torch.cat(xs), torch.cat(ys)
            perm = torch.from_numpy(rng.permutation(x.shape[0]))
            x, y = x[perm], y[perm]
            n_c = x.shape[0]
            batches = [(x[i:i + batch_size], y[i:i + batch_size]) for i in range(0, n_c, batch_size)]
        local_num[c] = n_c
        train_num += n_c
        if n_c > 0:
            train_local[c] = batches
            train_global += batches
        if t_cur + 1 < data.steps:
            tb = data.batches(c, t_cur + 1, batch_size)
            test_num += int(data.nsamp[t_cur + 1, c])
            if tb:
                test_local[c] = tb
                test_global += tb
    return C, train_num, test_num, train_global, test_global, local_num, train_local, test_local, data.class_num


def load_all_data(data: DriftData, batch_size: int, t_cur: int):
    """``[client][iter] -> list of (x, y) batches`` (parity: ``load_all_data_sea`` ``sea/data_loader.py:84-99``)."""
    return [[data.batches(c, it, batch_size) for it in range(t_cur + 1)] for c in range(data.client_num)]
