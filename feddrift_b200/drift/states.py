"""Per-algorithm drift states other than the soft-cluster family.

Parity: ``AdaState`` (``FedAvgEnsDataLoader.py:75-125``), ``KueState`` (``:32-55``), ``DriftSurfState``
(``:146-266``), ``MultiModelAccState`` (``:317-449``), AUE ensemble sizing (``:20-29``).  Differences by
design: states never pickle ``nn.Module``s (the reference's ``ds_state.pkl`` / ``mm_state.pkl`` do) — models
live in the :class:`ModelBank` and states only hold row indices; scoring goes through :class:`Evaluator`.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ..parallel.arena import ModelBank
from .evaluator import Evaluator


# ----------------------------------------------------------------------------- Adaptive-FedAvg
class AdaState:
    """Server learning-rate schedule from EMA mean / variance / ratio of the flat global parameters."""

    def __init__(self, init_lr: float = 1e-2, beta1: float = 0.5, beta2: float = 0.5, beta3: float = 0.5):
        self.init_lr, self.beta1, self.beta2, self.beta3 = init_lr, beta1, beta2, beta3
        self.eta = init_lr
        self.mu: Optional[torch.Tensor] = None
        self.s = 0.0
        self.gam = 0.0

    def update(self, theta: torch.Tensor, t: int) -> None:
        """``theta`` = flat fp32 parameter row (stays on its device; only one scalar crosses to the host)."""
        t = t + 1
        theta = theta.detach().reshape(-1)
        prev_mu = self.mu if self.mu is not None else torch.zeros_like(theta)
        if t != 1:
            prev_muh = prev_mu / (1 - self.beta1 ** (t - 1))
            prev_sh = self.s / (1 - self.beta2 ** (t - 1))
        else:
            prev_muh = torch.zeros_like(theta)
            prev_sh = 0.0
        new_mu = self.beta1 * prev_mu + (1 - self.beta1) * theta
        msd = ops.ada_stats(theta, prev_muh)  # mean((θ - μ̂)²): the only O(P) reduction (K8)
        new_s = self.beta2 * self.s + (1 - self.beta2) * msd
        new_sh = new_s / (1 - self.beta2 ** t)
        ratio = new_sh / prev_sh if prev_sh != 0 else 1.0
        new_gam = self.beta3 * self.gam + (1 - self.beta3) * ratio
        new_gamh = new_gam / (1 - self.beta3 ** t)
        self.eta = min(self.init_lr, (self.init_lr * new_gamh) / t)
        self.mu, self.s, self.gam = new_mu.clone(), float(new_s), float(new_gam)

    def current_lr(self) -> float:
        return float(self.eta)


# ----------------------------------------------------------------------------- KUE
class KueState:
    def __init__(self, model_num: int, feature_num: int, rng: Optional[np.random.RandomState] = None):
        self.model_num, self.feature_num = model_num, feature_num
        self.worst_idx = 0
        self.rng = rng if rng is not None else np.random.RandomState(0)
        self.masks = np.zeros((model_num, feature_num), dtype=bool)
        for m in range(model_num):
            self.initialize_mask(m)

    def set_worst_idx(self, m: int) -> None:
        self.worst_idx = int(m)

    def get_worst_idx(self) -> int:
        return self.worst_idx

    def get_masks(self) -> np.ndarray:
        return self.masks

    def initialize_mask(self, m: int) -> None:
        """Random feature subspace of random size r ∈ [1, F] (bits are OR-ed in, like the reference —
        a re-initialised model's mask can only grow)."""
        r = self.rng.randint(low=1, high=self.feature_num + 1)
        self.masks[m][self.rng.choice(self.feature_num, size=r, replace=False)] = True

    def masks_tensor(self, device="cpu") -> torch.Tensor:
        return torch.from_numpy(self.masks.astype(np.float32)).to(device)


# ----------------------------------------------------------------------------- DriftSurf
class DriftSurfState:
    """Two-model stable/reactive state machine.  ``models[key]`` is a row index of the bank (or None)."""

    def __init__(self, delta: float = 0.1, r: int = 3, wl: int = 10):
        self.reac_len, self.delta, self.win_len = r, delta, wl
        self.models: Dict[str, Optional[int]] = {"pred": None, "stab": None, "reac": None}
        self.snapshots: Dict[str, Optional[torch.Tensor]] = {"pred": None, "stab": None, "reac": None}
        self.train_data_dict: Dict[str, Optional[List[int]]] = {"pred": [0], "stab": [0], "reac": None}
        self.train_keys = ["pred", "stab"]
        self.acc_best = 0.0
        self.acc_dict = None
        self.reac_ctr = None
        self.state = "stab"
        self.model_key = "pred"

    # scoring: accuracy of the snapshot of ``key`` on the newest global data ------------------
    def _score(self, key: str, bank: ModelBank, ev: Evaluator, t: int, scratch_row: int) -> float:
        snap = self.snapshots[key]
        if snap is None:
            return 0.0
        saved = bank.theta[scratch_row].clone()
        bank.theta[scratch_row].copy_(snap.to(bank.device))
        acc = ev.acc_matrix([scratch_row], t)[0]
        bank.theta[scratch_row].copy_(saved)
        ns = ev.data.nsamp[t].double().cpu().numpy()
        return float((acc * ns).sum() / max(ns.sum(), 1.0))

    def _append(self, key: str, it: int) -> None:
        self.train_data_dict[key].append(it)
        if len(self.train_data_dict[key]) > self.win_len:
            self.train_data_dict[key].pop(0)

    def _reset(self, key: str) -> None:
        self.snapshots[key] = None
        self.train_data_dict[key] = []

    def get_train_keys(self):
        return self.train_keys

    def get_train_data(self, key):
        return self.train_data_dict[key]

    def get_model_key(self):
        return self.model_key

    def set_snapshot(self, key: str, row: torch.Tensor) -> None:
        self.snapshots[key] = row.detach().clone().cpu()

    def run_ds_algo(self, bank: ModelBank, ev: Evaluator, curr_iter: int, scratch_row: int = 0) -> None:
        acc_pred = self._score("pred", bank, ev, curr_iter, scratch_row)
        self.acc_best = max(self.acc_best, acc_pred)
        if self.state == "stab":
            acc_stab = 0.0 if len(self.train_data_dict["stab"]) == 0 else \
                self._score("stab", bank, ev, curr_iter, scratch_row)
            if acc_pred < self.acc_best - self.delta or acc_pred < acc_stab - self.delta / 2:
                self.state = "reac"
                self._reset("reac")
                self.reac_ctr = 0
                self.acc_dict = {"pred": np.zeros(self.reac_len), "reac": np.zeros(self.reac_len)}
            else:
                self._append("pred", curr_iter)
                self._append("stab", curr_iter)
                self.train_keys = ["pred", "stab"]
        if self.state == "reac":
            if self.reac_ctr > 0:
                acc_reac = self._score("reac", bank, ev, curr_iter, scratch_row)
                self.acc_dict["pred"][self.reac_ctr - 1] = acc_pred
                self.acc_dict["reac"][self.reac_ctr - 1] = acc_reac
                self.model_key = "reac" if acc_reac > acc_pred else "pred"
            self._append("pred", curr_iter)
            self._append("reac", curr_iter)
            self.train_keys = ["pred", "reac"]
            self.reac_ctr += 1
            if self.reac_ctr == self.reac_len:
                self.state = "stab"
                self._reset("stab")
                if np.mean(self.acc_dict["pred"]) < np.mean(self.acc_dict["reac"]):
                    self.snapshots["pred"] = self.snapshots["reac"]
                    self.train_data_dict["pred"] = self.train_data_dict["reac"]
                    self.acc_best = float(np.amax(self.acc_dict["reac"]))
                    self.model_key = "pred"
                self.acc_dict = None
                self.reac_ctr = None


# ----------------------------------------------------------------------------- legacy multi-model (mmacc / oracle)
class MultiModelAccState:
    def __init__(self, client_num: int, model_num: int = 2, delta: float = 0.1):
        self.client_num, self.model_num, self.delta = client_num, model_num, delta
        self.train_data_dict = {m: [[] for _ in range(client_num)] for m in range(model_num)}
        self.models: Dict[int, bool] = {}  # model slots that have been trained at least once
        self.train_model_idx: Dict[int, int] = {}
        self.test_model_idx: Dict[int, int] = {}
        self.acc_dict: Dict[int, float] = {}

    def run_model_select(self, ev: Optional[Evaluator], curr_iter: int) -> None:
        if curr_iter == 0:
            for c in range(self.client_num):
                self.train_data_dict[0][c].append(0)
                self.train_model_idx[c] = self.test_model_idx[c] = 0
            return
        next_free = next((m for m in range(self.model_num) if m not in self.models), -1)
        known = sorted(self.models.keys())
        acc = ev.acc_matrix(known, curr_iter)
        for c in range(self.client_num):
            best_model, best_acc = -1, 0.0
            for r, m in enumerate(known):
                if acc[r][c] > best_acc:
                    best_acc, best_model = float(acc[r][c]), m
            if self.acc_dict[c] - best_acc > self.delta and next_free != -1:
                best_model = next_free
            self.train_data_dict[best_model][c].append(curr_iter)
            self.train_model_idx[c] = self.test_model_idx[c] = best_model

    def model_select_geni(self, curr_iter: int, change_points, stretch: int) -> None:
        for c in range(self.client_num):
            m = int(change_points[curr_iter // stretch][c])
            self.train_data_dict[m][c].append(curr_iter)
            self.train_model_idx[c] = self.test_model_idx[c] = m

    def model_select_geniex(self, curr_iter: int, change_points, stretch: int) -> None:
        rows = [t for t in range(change_points.shape[0]) if change_points[t].any()]
        min_cp = rows[0] * stretch if rows else 10 ** 6
        for c in range(self.client_num):
            train_m = int(change_points[curr_iter // stretch][c])
            test_m = int(change_points[(curr_iter + 1) // stretch][c]) if curr_iter >= min_cp else train_m
            self.train_data_dict[train_m][c].append(curr_iter)
            self.train_model_idx[c], self.test_model_idx[c] = train_m, test_m

    def set_model(self, key: int) -> None:
        self.models[key] = True

    def set_acc(self, client: int, acc: float) -> None:
        self.acc_dict[client] = float(acc)

    def get_train_data_by_model(self, key: int) -> str:
        td = self.train_data_dict[key]
        return json.dumps(td) if any(len(x) > 0 for x in td) else ""

    def get_test_model_idx(self, c: int) -> int:
        return self.test_model_idx[c]

    def get_train_model_idx(self, c: int) -> int:
        return self.train_model_idx[c]


def aue_model_num(curr_train_iteration: int, ensemble_window: int) -> int:
    return min(curr_train_iteration + 1, ensemble_window)
