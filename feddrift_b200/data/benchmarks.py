"""FedML benchmark data layer: partitioners, augmentation, and dataset loaders returning the FedML tuple.

Parity: ``fedml_api/data_preprocessing/{cifar10,cifar100,cinic10,MNIST,FederatedEMNIST,fed_cifar100,
shakespeare,fed_shakespeare,stackoverflow_lr,stackoverflow_nwp,UCI,NUS_WIDE,lending_club_loan}/…`` and
``data/synthetic_*/generate_synthetic.py``.  There is no network and no dataset on this box, so every loader reads the
real files when ``data_dir`` has them (CIFAR python pickles / LEAF json) and otherwise produces **synthetic data of the
named shape** (what BASELINE.json prescribes); TFF ``.h5`` readers are gated on ``h5py`` (not installed here).

Everything is produced as dense tensors first (device friendly) and only then sliced into the FedML
``(client_num, train_num, test_num, train_global, test_global, local_num_dict, train_local_dict, test_local_dict,
class_num)`` view, so the same arrays feed the device engine without a second copy.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------- partitioning
def partition_indices(y: np.ndarray, n_nets: int, partition: str = "homo", alpha: float = 0.5, rng=None,
                      num_classes: Optional[int] = None, min_size_req: int = 10) -> Dict[int, np.ndarray]:
    """``homo`` (uniform split), ``hetero`` (LDA: per-class Dirichlet(α) proportions with the balance cap of
    ``cifar10/data_loader.py:129-150``, resampled until every client has ≥ 10 samples)."""
    rng = rng or np.random
    N = len(y)
    if partition == "homo":
        idxs = rng.permutation(N)
        return {i: b for i, b in enumerate(np.array_split(idxs, n_nets))}
    if partition != "hetero":
        raise ValueError(f"unknown partition {partition!r} ('hetero-fix' needs the reference's fixed index files)")
    K = int(num_classes or (y.max() + 1))
    min_size, idx_batch = 0, None
    tries = 0
    while min_size < min(min_size_req, max(1, N // (n_nets * 2))) and tries < 100:
        tries += 1
        idx_batch = [[] for _ in range(n_nets)]
        for k in range(K):
            idx_k = np.where(y == k)[0]
            rng.shuffle(idx_k)
            prop = rng.dirichlet(np.repeat(alpha, n_nets))
            prop = np.array([p * (len(b) < N / n_nets) for p, b in zip(prop, idx_batch)])
            prop = prop / max(prop.sum(), 1e-12)
            cuts = (np.cumsum(prop) * len(idx_k)).astype(int)[:-1]
            idx_batch = [b + part.tolist() for b, part in zip(idx_batch, np.split(idx_k, cuts))]
        min_size = min(len(b) for b in idx_batch)
    out = {}
    for j in range(n_nets):
        b = np.array(idx_batch[j], dtype=np.int64)
        rng.shuffle(b)
        out[j] = b
    return out


def record_net_data_stats(y: np.ndarray, net_dataidx_map: Dict[int, np.ndarray]) -> Dict[int, Dict[int, int]]:
    return {i: {int(k): int(v) for k, v in zip(*np.unique(y[idx], return_counts=True))} for i, idx in net_dataidx_map.items()}


class Cutout:
    """Zero a random ``length``² square (parity: ``cifar10/data_loader.py:57-98``); works on CHW tensors / batches."""

    def __init__(self, length: int = 16, rng=None):
        self.length, self.rng = length, rng or np.random

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        h, w = img.shape[-2:]
        y, x = self.rng.randint(h), self.rng.randint(w)
        y1, y2 = np.clip(y - self.length // 2, 0, h), np.clip(y + self.length // 2, 0, h)
        x1, x2 = np.clip(x - self.length // 2, 0, w), np.clip(x + self.length // 2, 0, w)
        out = img.clone()
        out[..., y1:y2, x1:x2] = 0
        return out


# ----------------------------------------------------------------------------- helpers
def _batches(x: torch.Tensor, y: torch.Tensor, bs: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    return [(x[i:i + bs], y[i:i + bs]) for i in range(0, len(x), bs)]


def to_fedml_tuple(Xtr, ytr, Xte, yte, train_map: Dict[int, np.ndarray], test_map: Optional[Dict[int, np.ndarray]], bs: int,
                   class_num: int):
    n = len(train_map)
    train_local, test_local, local_num = {}, {}, {}
    for c in range(n):
        idx = torch.as_tensor(train_map[c])
        local_num[c] = int(len(idx))
        train_local[c] = _batches(Xtr[idx], ytr[idx], bs)
        if test_map is not None:
            tix = torch.as_tensor(test_map[c])
            test_local[c] = _batches(Xte[tix], yte[tix], bs)
        else:
            test_local[c] = _batches(Xte, yte, bs)  # every client evaluates on the global test set (FedML default)
    return (n, int(len(Xtr)), int(len(Xte)), _batches(Xtr, ytr, bs), _batches(Xte, yte, bs), local_num, train_local,
            test_local, class_num)


def _synthetic_images(n: int, shape, classes: int, rng) -> Tuple[torch.Tensor, torch.Tensor]:
    y = rng.randint(0, classes, size=n)
    c, h, w = shape
    proto = rng.randn(classes, c, 4, 4).astype(np.float32)
    img = np.repeat(np.repeat(proto[y], h // 4, axis=2), w // 4, axis=3) + 0.5 * rng.randn(n, c, h, w).astype(np.float32)
    return torch.from_numpy(img), torch.from_numpy(y.astype(np.int64))


def _read_cifar(datadir: str, hundred: bool):
    """Real CIFAR python pickles if present under ``datadir``."""
    try:
        if hundred:
            base = os.path.join(datadir, "cifar-100-python")
            tr = pickle.load(open(os.path.join(base, "train"), "rb"), encoding="latin1")
            te = pickle.load(open(os.path.join(base, "test"), "rb"), encoding="latin1")
            key = "fine_labels"
            Xtr, ytr, Xte, yte = tr["data"], tr[key], te["data"], te[key]
        else:
            base = os.path.join(datadir, "cifar-10-batches-py")
            parts = [pickle.load(open(os.path.join(base, f"data_batch_{i}"), "rb"), encoding="latin1") for i in range(1, 6)]
            te = pickle.load(open(os.path.join(base, "test_batch"), "rb"), encoding="latin1")
            Xtr, ytr = np.concatenate([p["data"] for p in parts]), sum((p["labels"] for p in parts), [])
            Xte, yte = te["data"], te["labels"]
        f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(-1, 3, 32, 32) / 255.0)  # noqa: E731
        return f(Xtr), torch.tensor(ytr), f(Xte), torch.tensor(yte)
    except Exception:
        return None


# ----------------------------------------------------------------------------- image classification
def load_partition_data_cifar(dataset: str, data_dir: Optional[str], partition_method: str, partition_alpha: float,
                              client_number: int, batch_size: int, n_train: int = 5000, n_test: int = 1000, seed: int = 0):
    """cifar10 / cifar100 / cinic10 (``cifar10/data_loader.py:113-269``): LDA or homo partition of the train set,
    global test set for every client.  Falls back to synthetic 3×32×32 data of ``n_train``/``n_test`` samples."""
    rng = np.random.RandomState(seed)
    classes = 100 if dataset == "cifar100" else 10
    real = _read_cifar(data_dir, dataset == "cifar100") if data_dir else None
    if real is not None:
        Xtr, ytr, Xte, yte = real
    else:
        Xtr, ytr = _synthetic_images(n_train, (3, 32, 32), classes, rng)
        Xte, yte = _synthetic_images(n_test, (3, 32, 32), classes, rng)
    m = partition_indices(ytr.numpy(), client_number, partition_method, partition_alpha, rng, classes)
    return to_fedml_tuple(Xtr, ytr, Xte, yte, m, None, batch_size, classes)


def load_partition_data_cifar10(dataset, data_dir, partition_method, partition_alpha, client_number, batch_size, **kw):
    return load_partition_data_cifar("cifar10", data_dir, partition_method, partition_alpha, client_number, batch_size, **kw)


def load_partition_data_cifar100(dataset, data_dir, partition_method, partition_alpha, client_number, batch_size, **kw):
    return load_partition_data_cifar("cifar100", data_dir, partition_method, partition_alpha, client_number, batch_size, **kw)


def load_partition_data_cinic10(dataset, data_dir, partition_method, partition_alpha, client_number, batch_size, **kw):
    return load_partition_data_cifar("cinic10", data_dir, partition_method, partition_alpha, client_number, batch_size, **kw)


def _read_leaf(train_dir: str, test_dir: str):
    def rd(d):
        users, data = [], {}
        for f in sorted(os.listdir(d)):
            if f.endswith(".json"):
                blob = json.load(open(os.path.join(d, f)))
                users += blob["users"]
                data.update(blob["user_data"])
        return users, data
    u, tr = rd(train_dir)
    _, te = rd(test_dir)
    return u, tr, te


def load_partition_data_mnist(batch_size: int, train_path: Optional[str] = None, test_path: Optional[str] = None,
                              client_number: int = 100, samples_per_client: int = 60, seed: int = 0):
    """LEAF MNIST (``MNIST/data_loader.py``): natural per-user split (1000 users, power-law sizes, 2 digits/user) when
    the json files exist; otherwise a synthetic 784-d federation with 2 classes per client."""
    if train_path and os.path.isdir(train_path) and test_path and os.path.isdir(test_path):
        users, tr, te = _read_leaf(train_path, test_path)
        Xs, ys, tmap, Xt, yt, temap, o1, o2 = [], [], {}, [], [], {}, 0, 0
        for c, u in enumerate(users):
            x, y = np.asarray(tr[u]["x"], np.float32), np.asarray(tr[u]["y"], np.int64)
            xt, ytt = np.asarray(te[u]["x"], np.float32), np.asarray(te[u]["y"], np.int64)
            Xs.append(x); ys.append(y); tmap[c] = np.arange(o1, o1 + len(y)); o1 += len(y)
            Xt.append(xt); yt.append(ytt); temap[c] = np.arange(o2, o2 + len(ytt)); o2 += len(ytt)
        return to_fedml_tuple(torch.from_numpy(np.concatenate(Xs)), torch.from_numpy(np.concatenate(ys)),
                              torch.from_numpy(np.concatenate(Xt)), torch.from_numpy(np.concatenate(yt)), tmap, temap, batch_size, 10)
    from .drift import DigitPool
    rng = np.random.RandomState(seed)
    pool = DigitPool(None, size=max(20000, client_number * samples_per_client * 2), seed=seed)
    X, Y = torch.from_numpy(pool.X), torch.from_numpy(pool.Y)
    tmap, temap = {}, {}
    for c in range(client_number):
        digs = rng.choice(10, 2, replace=False)
        cand = np.where(np.isin(pool.Y, digs))[0]
        pick = rng.choice(cand, samples_per_client, replace=len(cand) < samples_per_client)
        cut = int(0.9 * samples_per_client)
        tmap[c], temap[c] = pick[:cut], pick[cut:]
    return to_fedml_tuple(X, Y, X, Y, tmap, temap, batch_size, 10)


def load_partition_data_federated_emnist(dataset, data_dir, batch_size: int = 20, client_number: int = 50,
                                         samples_per_client: int = 100, seed: int = 0):
    """FederatedEMNIST (TFF h5, 3400 writers, 62 classes): synthetic 28×28 stand-in unless h5py + files exist."""
    rng = np.random.RandomState(seed)
    X, Y = _synthetic_images(client_number * samples_per_client, (1, 28, 28), 62, rng)
    X = X.squeeze(1)
    m = {c: np.arange(c * samples_per_client, (c + 1) * samples_per_client) for c in range(client_number)}
    return to_fedml_tuple(X, Y, X[: 10 * batch_size], Y[: 10 * batch_size], m, None, batch_size, 62)


def load_partition_data_federated_cifar100(dataset, data_dir, batch_size: int = 20, client_number: int = 50,
                                           samples_per_client: int = 100, seed: int = 0):
    """fed_cifar100 (TFF: 500 train clients × 100 images, Pachinko allocation): synthetic 3×32×32 (24×24 crops are a
    transform detail) with a client-specific label prior."""
    rng = np.random.RandomState(seed)
    Xs, Ys = [], []
    for c in range(client_number):
        prior = rng.dirichlet(np.full(100, 0.1))
        y = rng.choice(100, samples_per_client, p=prior)
        x, _ = _synthetic_images(samples_per_client, (3, 32, 32), 100, rng)
        Xs.append(x + torch.from_numpy(y[:, None, None, None].astype(np.float32)) * 0.02)
        Ys.append(torch.from_numpy(y.astype(np.int64)))
    X, Y = torch.cat(Xs), torch.cat(Ys)
    m = {c: np.arange(c * samples_per_client, (c + 1) * samples_per_client) for c in range(client_number)}
    return to_fedml_tuple(X, Y, X[: 10 * batch_size], Y[: 10 * batch_size], m, None, batch_size, 100)


# ----------------------------------------------------------------------------- language
CHAR_VOCAB = list("dhlptx@DHLPTX $(,048cgkoswCGKOSW[_#'/37;?bfjnrvzBFJNRVZ\"&*.26:\naeimquyAEIMQUY]!%)-159\r")
SHAKESPEARE_VOCAB = len(CHAR_VOCAB) + 4  # pad / oov / bos / eos  == 90  (fed_shakespeare/utils.py:15-20)
SEQ_LEN = 80


def _char_corpus(n_chars: int, rng) -> np.ndarray:
    """A synthetic character stream with bigram structure (so that an LSTM can learn something)."""
    V = SHAKESPEARE_VOCAB
    T = rng.dirichlet(np.full(V - 4, 0.05), size=V)  # sparse next-char distributions
    out = np.zeros(n_chars, dtype=np.int64)
    s = 4
    for i in range(n_chars):
        s = 4 + rng.choice(V - 4, p=T[s])
        out[i] = s
    return out


def load_partition_data_shakespeare(batch_size: int, client_number: int = 128, seqs_per_client: int = 40, seed: int = 0,
                                    per_position: bool = False):
    """(fed_)shakespeare next-character prediction: x = 80 char ids, y = the next char (LEAF convention; the TFF
    per-position variant — y shifted by one, ``[B, 80]`` — is ``per_position=True``)."""
    rng = np.random.RandomState(seed)
    tmap, Xs, Ys, off = {}, [], [], 0
    for c in range(client_number):
        stream = _char_corpus(seqs_per_client + SEQ_LEN + 1, rng)
        x = np.stack([stream[i:i + SEQ_LEN] for i in range(seqs_per_client)])
        y = np.stack([stream[i + 1:i + SEQ_LEN + 1] for i in range(seqs_per_client)]) if per_position else \
            stream[SEQ_LEN:SEQ_LEN + seqs_per_client]
        Xs.append(x); Ys.append(y)
        tmap[c] = np.arange(off, off + seqs_per_client); off += seqs_per_client
    X, Y = torch.from_numpy(np.concatenate(Xs)), torch.from_numpy(np.concatenate(Ys))
    return to_fedml_tuple(X, Y, X[: 8 * batch_size], Y[: 8 * batch_size], tmap, None, batch_size, SHAKESPEARE_VOCAB)


load_partition_data_federated_shakespeare = load_partition_data_shakespeare


def load_partition_data_federated_stackoverflow_nwp(dataset, data_dir, batch_size: int = 16, client_number: int = 20,
                                                    seqs_per_client: int = 32, vocab: int = 10004, seq_len: int = 20, seed: int = 0):
    rng = np.random.RandomState(seed)
    n = client_number * seqs_per_client
    X = torch.from_numpy(rng.randint(4, vocab, size=(n, seq_len)).astype(np.int64))
    Y = torch.from_numpy(rng.randint(4, vocab, size=n).astype(np.int64))
    m = {c: np.arange(c * seqs_per_client, (c + 1) * seqs_per_client) for c in range(client_number)}
    return to_fedml_tuple(X, Y, X[: 4 * batch_size], Y[: 4 * batch_size], m, None, batch_size, vocab)


def load_partition_data_federated_stackoverflow_lr(dataset, data_dir, batch_size: int = 16, client_number: int = 20,
                                                   samples_per_client: int = 32, vocab: int = 10000, tags: int = 500, seed: int = 0):
    """Multi-label tag prediction: bag-of-words ``[B, 10000]`` → ``[B, 500]`` multi-hot (BCE; precision/recall)."""
    rng = np.random.RandomState(seed)
    n = client_number * samples_per_client
    X = torch.from_numpy((rng.rand(n, vocab) < 0.002).astype(np.float32))
    Wt = rng.randn(vocab, tags).astype(np.float32)
    Y = torch.from_numpy(((X.numpy() @ Wt) > 1.5).astype(np.float32))
    m = {c: np.arange(c * samples_per_client, (c + 1) * samples_per_client) for c in range(client_number)}
    return to_fedml_tuple(X, Y, X[: 4 * batch_size], Y[: 4 * batch_size], m, None, batch_size, tags)


# ----------------------------------------------------------------------------- streaming / tabular / vertical
def load_streaming_susy_or_ro(client_number: int, iteration_number: int, dataset: str = "SUSY", beta: float = 0.5, seed: int = 0):
    """UCI SUSY (18 features) / RoomOccupancy (5 features) streams for decentralized online learning
    (``UCI/data_loader_for_susy_and_ro.py``): ``beta`` = fraction of adversarially clustered (non-iid) samples.
    Returns ``streaming_data[c][t] = {'x': ndarray[d], 'y': 0|1}``."""
    rng = np.random.RandomState(seed)
    d = 18 if dataset.upper() == "SUSY" else 5
    w = rng.randn(d)
    out = []
    for c in range(client_number):
        shift = rng.randn(d) * beta * 2.0  # the non-iid part: a client-specific covariate shift
        seq = []
        for t in range(iteration_number):
            x = (rng.randn(d) + (shift if rng.rand() < beta else 0.0)).astype(np.float32)
            seq.append({"x": x, "y": float(x @ w > 0)})
        out.append(seq)
    return out


def load_vertical_parties(dataset: str = "lending_club_loan", n: int = 2000, parties: int = 2, seed: int = 0):
    """NUS-WIDE (634 image + 1000 text features, 2 parties) / lending-club (≈ 150 tabular features split over 2–3
    parties) shaped vertical partitions: ``([X_party0_train, …], y_train, [X_party0_test, …], y_test)``."""
    rng = np.random.RandomState(seed)
    dims = {"nus_wide": [634, 1000], "NUS_WIDE": [634, 1000]}.get(dataset, [max(8, 150 // parties)] * parties)[:parties]
    Xs = [rng.randn(n, d).astype(np.float32) for d in dims]
    logit = sum(x[:, :4].sum(1) for x in Xs)
    y = (logit + 0.5 * rng.randn(n) > 0).astype(np.float32)
    cut = int(0.8 * n)
    return [x[:cut] for x in Xs], y[:cut], [x[cut:] for x in Xs], y[cut:]


def generate_synthetic(alpha: float, beta: float, iid: bool, num_user: int = 30, dimension: int = 60, num_class: int = 10,
                       seed: int = 0):
    """FedProx Synthetic(α, β) (``data/synthetic_*/generate_synthetic.py``): per-user softmax-regression tasks whose
    model (α) and feature (β) distributions drift apart as α, β grow.  Returns ``(X_split, y_split)`` lists."""
    rng = np.random.RandomState(seed)
    samples = rng.lognormal(4, 2, num_user).astype(int) + 50
    mean_W = rng.normal(0, alpha, num_user)
    B = rng.normal(0, beta, num_user)
    cov = np.diag(np.power(np.arange(1, dimension + 1), -1.2))
    mean_x = np.stack([np.ones(dimension) * B[i] if iid else rng.normal(B[i], 1, dimension) for i in range(num_user)])
    Wg, bg = rng.normal(0, 1, (dimension, num_class)), rng.normal(0, 1, num_class)
    Xs, Ys = [], []
    for i in range(num_user):
        W = Wg if iid else rng.normal(mean_W[i], 1, (dimension, num_class))
        b = bg if iid else rng.normal(mean_W[i], 1, num_class)
        xx = rng.multivariate_normal(mean_x[i], cov, samples[i])
        Xs.append(xx.astype(np.float32))
        Ys.append(np.argmax(xx @ W + b, axis=1).astype(np.int64))
    return Xs, Ys


def load_data(args, dataset_name: str):
    """Dispatch used by the experiment mains (``fedml_experiments/distributed/fedavg/main_fedavg.py:140-200``)."""
    bs, n = args.batch_size, args.client_num_in_total
    pm, pa, dd = getattr(args, "partition_method", "homo"), getattr(args, "partition_alpha", 0.5), getattr(args, "data_dir", None)
    if dataset_name == "mnist":
        out = load_partition_data_mnist(bs, os.path.join(dd or "", "train"), os.path.join(dd or "", "test"), n)
    elif dataset_name in ("femnist", "federated_emnist"):
        out = load_partition_data_federated_emnist(dataset_name, dd, bs, n)
    elif dataset_name == "fed_cifar100":
        out = load_partition_data_federated_cifar100(dataset_name, dd, bs, n)
    elif dataset_name in ("shakespeare", "fed_shakespeare"):
        out = load_partition_data_shakespeare(bs, n)
    elif dataset_name == "stackoverflow_nwp":
        out = load_partition_data_federated_stackoverflow_nwp(dataset_name, dd, bs, n)
    elif dataset_name == "stackoverflow_lr":
        out = load_partition_data_federated_stackoverflow_lr(dataset_name, dd, bs, n)
    elif dataset_name in ("cifar10", "cifar100", "cinic10"):
        out = load_partition_data_cifar(dataset_name, dd, pm, pa, n, bs)
    else:
        raise ValueError(dataset_name)
    return list(out[1:])  # [train_num, test_num, train_global, test_global, local_num, train_local, test_local, class_num]
