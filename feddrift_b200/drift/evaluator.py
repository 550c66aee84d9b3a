"""Model × client evaluation (K4).  One object answers every "accuracy of model m on data d" question
the drift algorithms ask, so the algorithms never touch devices, batches or ``.item()``.

Parity of semantics: ``_infer`` (``FedAvgEnsAggregatorSoftCluster.py:305-326``: correct count,
sample count, Σ batch_mean_loss·batch_size == Σ per-sample loss), ``train_acc_matrix``
(``FedAvgEnsDataLoader.py:1074-1085``), ``_infer_subset`` (``:1111-1138``, evaluates
``subset_size + 1`` batches because of the ``>`` comparison).

Small MLPs go through ``ops.mlp_eval_matrix`` (one launch for the whole [M, C] matrix, no host sync
until the single result copy); other models go through the bound ``nn.Module`` + ``ops.eval_logits``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops
from ..data.drift import DriftData
from ..parallel.arena import ModelBank


class Evaluator:
    def __init__(self, bank: ModelBank, data: DriftData, batch_size: int, eval_batch: int = 1024):
        self.bank, self.data, self.batch_size = bank, data, batch_size
        self.eval_batch = eval_batch

    # -- primitive ---------------------------------------------------------------------
    def infer_samples(self, m: int, x: torch.Tensor, y: torch.Tensor, mask: Optional[torch.Tensor] = None
                      ) -> Tuple[float, float, float]:
        """-> (correct, total, loss_sum) of model m on samples (x, y)."""
        n = int(y.shape[0])
        if n == 0:
            return 0.0, 0.0, 0.0
        dev = self.bank.device
        s = self.bank.mlp
        if s is not None and mask is None:  # small MLP: the K4 kernel on a single pooled "client"
            xs = x.reshape(1, n, -1).to(dev).float()
            corr, loss = ops.mlp_eval_matrix(self.bank.theta[m:m + 1], xs, y.reshape(1, n).to(dev),
                                             torch.tensor([n], dtype=torch.int32, device=dev), s["kind"], s["in"], s["hidden"],
                                             s["out"])
            return float(corr[0, 0]), float(n), float(loss[0, 0])
        acc = torch.zeros(3, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for i in range(0, n, self.eval_batch):
                xb = x[i:i + self.eval_batch].to(dev)
                if mask is not None:
                    xb = xb * mask.to(dev).reshape((1,) + tuple(xb.shape[1:]))
                logits = self.bank.forward(m, xb)
                ops.eval_logits(logits, y[i:i + self.eval_batch].to(dev), acc)
        a = acc.tolist()
        return a[0], a[2], a[1]

    def infer_client(self, m: int, c: int, t: int, mask=None):
        n = int(self.data.nsamp[t, c])
        return self.infer_samples(m, self.data.X[t, c, :n], self.data.Y[t, c, :n], mask)

    # -- matrices ----------------------------------------------------------------------
    def acc_matrix(self, model_ids: Sequence[int], t: int) -> np.ndarray:
        """acc[row, c] of models ``model_ids`` on every client's time-t data."""
        C = self.data.client_num
        if len(model_ids) == 0:
            return np.zeros((0, C))
        s = self.bank.mlp
        if s is not None:
            theta = self.bank.theta[list(model_ids)]
            X = self.data.X[t].reshape(C, self.data.X.shape[2], -1).to(self.bank.device)
            corr, _ = ops.mlp_eval_matrix(theta, X, self.data.Y[t].to(self.bank.device),
                                          self.data.nsamp[t].to(self.bank.device),
                                          s["kind"], s["in"], s["hidden"], s["out"])
            ns = self.data.nsamp[t].clamp(min=1).to(corr.device).float()
            acc = (corr / ns).double().cpu().numpy()
            acc[:, (self.data.nsamp[t] == 0).cpu().numpy()] = 0.0
            return acc
        # module path: ONE batched forward per model over (chunks of) all clients' samples, per-client correct counts by
        # a masked reduction on device, a single host copy for the whole matrix (instead of M·C forwards + .tolist()s)
        dev = self.bank.device
        S = self.data.X.shape[2]
        X, Y = self.data.X[t].to(dev), self.data.Y[t].to(dev).long()
        ns = self.data.nsamp[t].to(dev)
        mask = (torch.arange(S, device=dev)[None, :] < ns[:, None]).float()
        per_chunk = max(1, self.eval_batch * 8 // max(S, 1))
        corr = torch.zeros(len(model_ids), C, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for r, m in enumerate(model_ids):
                for c0 in range(0, C, per_chunk):
                    c1 = min(C, c0 + per_chunk)
                    logits = self.bank.forward(m, X[c0:c1].reshape((c1 - c0) * S, *X.shape[2:]))
                    hit = (logits.argmax(-1) == Y[c0:c1].reshape(-1)).float().reshape(c1 - c0, S)
                    corr[r, c0:c1] = (hit * mask[c0:c1]).sum(1)
        acc = (corr / ns.clamp(min=1).float()[None, :]).double().cpu().numpy()
        acc[:, (self.data.nsamp[t] == 0).cpu().numpy()] = 0.0
        return acc

    def pooled_acc(self, m: int, pairs: List[Tuple[int, int]], max_batches: int, rng) -> float:
        """Accuracy of model m on the pooled, shuffled batches of the (client, time) pairs — at most
        ``max_batches + 1`` batches (reference quirk)."""
        batches = [(c, t, b) for (c, t) in pairs
                   for b in range(-(-int(self.data.nsamp[t, c]) // self.batch_size))]
        if not batches:
            return 0.0
        order = rng.permutation(len(batches))[: max_batches + 1]
        xs, ys = [], []
        for i in order:
            c, t, b = batches[i]
            lo, hi = b * self.batch_size, min((b + 1) * self.batch_size, int(self.data.nsamp[t, c]))
            xs.append(self.data.X[t, c, lo:hi])
            ys.append(self.data.Y[t, c, lo:hi])
        k, n, _ = self.infer_samples(m, torch.cat(xs), torch.cat(ys))
        return k / n if n else 0.0
