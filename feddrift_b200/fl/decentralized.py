"""Decentralized online learning: DSGD and PushSum over (optionally time-varying) gossip topologies.

Parity: ``fedml_api/standalone/decentralized/{decentralized_fl_api,client_dsgd,client_pushsum,topology_manager}.py``
(SURVEY §2.4): streaming logistic regression (SUSY / RoomOccupancy shaped), one sample per client per iteration,
``x ← x − η∇f(z)``, gossip ``x_j ← W_jj x_j + Σ_i W_ij x_i`` (K12), PushSum de-biasing ``z = x / ω``, metric =
average regret.  The reference keeps one ``nn.Module`` pair per client and mixes parameter-by-parameter in Python;
here all clients' parameters are ONE ``[n, P]`` matrix, the per-client gradients of an iteration are computed in one
batched pass and the mixing is ONE ``ops.gossip_mix`` launch (``Wᵀ·X``).
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from .. import ops
from ..core.topology import AsymmetricTopologyManager, SymmetricTopologyManager
from ..utils.metrics import get_sink


class TopologyManager:
    """The standalone package's combined manager (``topology_manager.py``): symmetric or asymmetric."""

    def __init__(self, n, b_symmetric, undirected_neighbor_num=5, out_directed_neighbor=5, rng=None):
        self.n, self.b_symmetric = n, b_symmetric
        self.undirected_neighbor_num, self.out_directed_neighbor = undirected_neighbor_num, out_directed_neighbor
        self.rng = rng
        self.topology_symmetric = self.topology_asymmetric = None

    def generate_topology(self):
        if self.b_symmetric:
            m = SymmetricTopologyManager(self.n, self.undirected_neighbor_num)
            m.generate_topology()
            self.topology_symmetric = m.topology
        else:
            m = AsymmetricTopologyManager(self.n, self.undirected_neighbor_num, self.out_directed_neighbor, rng=self.rng)
            m.generate_topology()
            self.topology_asymmetric = m.topology
        return self.matrix()

    def matrix(self) -> np.ndarray:
        return self.topology_symmetric if self.b_symmetric else self.topology_asymmetric

    def get_symmetric_neighbor_list(self, client_idx):
        return [] if client_idx >= self.n else self.topology_symmetric[client_idx]

    def get_asymmetric_neighbor_list(self, client_idx):
        return [] if client_idx >= self.n else self.topology_asymmetric[client_idx]


def cal_regret(loss_history: torch.Tensor, client_number: int, t: int) -> float:
    """Σ_clients Σ_{τ≤t} loss / (n·(t+1))  (parity: ``decentralized_fl_api.py:11-17``)."""
    return float(loss_history[: t + 1].sum()) / (client_number * (t + 1))


class DecentralizedSimulator:
    """Linear-model online learner for all clients at once.  ``streaming_data[c][t] = {'x': ndarray[d], 'y': 0|1}``."""

    def __init__(self, client_number: int, streaming_data, input_dim: int, args, device="cpu"):
        self.n, self.args, self.device = client_number, args, torch.device(device)
        T = args.iteration_number
        self.X = torch.tensor(np.stack([[np.asarray(streaming_data[c][t]["x"], dtype=np.float32) for t in range(T)]
                                        for c in range(client_number)]), device=self.device)       # [n, T, d]
        self.Y = torch.tensor(np.asarray([[float(streaming_data[c][t]["y"]) for t in range(T)]
                                          for c in range(client_number)], dtype=np.float32), device=self.device)
        g = torch.Generator().manual_seed(int(getattr(args, "seed", 0)))
        bound = 1.0 / np.sqrt(input_dim)
        init = (torch.rand(input_dim + 1, generator=g) * 2 - 1) * bound                               # Linear(d, 1) init
        self.x = init.repeat(client_number, 1).to(self.device)    # push-sum numerators (model_x)
        self.z = self.x.clone()                                   # de-biased models the loss is evaluated at
        self.omega = torch.ones(client_number, device=self.device)
        self.topo = TopologyManager(client_number, bool(args.b_symmetric), args.topology_neighbors_num_undirected,
                                    getattr(args, "topology_neighbors_num_directed", 0),
                                    rng=np.random.RandomState(int(getattr(args, "seed", 0))))
        self.W = torch.tensor(self.topo.generate_topology(), dtype=torch.float32, device=self.device)
        self.loss_history: List[torch.Tensor] = []
        self.sink = get_sink()

    def _grad(self, t: int):
        """BCE(sigmoid(w·x + b), y) and its gradient at z, for every client's sample of iteration t."""
        xt, yt = self.X[:, t, :], self.Y[:, t]
        logit = (self.z[:, :-1] * xt).sum(1) + self.z[:, -1]
        p = torch.sigmoid(logit)
        loss = -(yt * torch.log(p.clamp_min(1e-12)) + (1 - yt) * torch.log((1 - p).clamp_min(1e-12)))
        d = (p - yt)
        return loss, torch.cat([d[:, None] * xt, d[:, None]], dim=1)

    def step(self, t_global: int) -> float:
        a = self.args
        t = t_global % a.iteration_number
        mode = getattr(a, "mode", "DOL")
        if mode == "PUSHSUM" and getattr(a, "time_varying", False):
            np.random.seed(t)
            self.topo.rng = np.random.RandomState(t)
            self.W = torch.tensor(self.topo.generate_topology(), dtype=torch.float32, device=self.device)
        loss, g = self._grad(t)
        wd = float(getattr(a, "weight_decay", 0.0))
        self.x.add_(g + wd * self.z, alpha=-a.learning_rate)
        if mode in ("DOL", "PUSHSUM"):
            Wt = self.W.t().contiguous()              # receiver j mixes column j of the (row-stochastic) matrix
            self.x = ops.gossip_mix(self.x, Wt)
            if mode == "PUSHSUM":
                self.omega = Wt @ self.omega
                self.z = self.x / self.omega[:, None]
            else:
                self.z = self.x.clone()
        else:                                          # 'LOCAL': no communication
            self.z = self.x.clone()
        self.loss_history.append(loss.sum())
        return float(loss.mean())

    def run(self) -> List[float]:
        a = self.args
        regrets = []
        total = a.iteration_number * getattr(a, "epoch", 1)
        for t in range(total):
            self.step(t)
            if (t + 1) % max(1, getattr(a, "log_every", 1)) == 0 or t == total - 1:
                r = cal_regret(torch.stack(self.loss_history), self.n, t)
                regrets.append(r)
                self.sink.log({"Average Loss": r, "iteration": t})
        return regrets


def FedML_decentralized_fl(client_number, client_id_list, streaming_data, model, model_cache, args, device="cpu"):
    """Drop-in for ``decentralized_fl_api.FedML_decentralized_fl`` (``model``/``model_cache`` give the input dim)."""
    d = model.linear.in_features if hasattr(model, "linear") else next(model.parameters()).shape[1]
    sim = DecentralizedSimulator(client_number, [streaming_data[c] for c in client_id_list], d, args, device)
    return sim.run()
