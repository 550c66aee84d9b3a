"""Single-process FL simulators: FedAvg, FedOpt (+ OptRepo), hierarchical FL.

Parity: ``fedml_api/standalone/fedavg/{fedavg_trainer,client}.py``, ``standalone/fedopt/{fedopt_trainer,optrepo}.py``,
``standalone/hierarchical_fl/trainer.py``.  The reference deep-copies the global ``nn.Module`` for
every client and averages state_dicts key by key on the CPU; here the global model is one arena row, every client
trains into its own row of a ``[clients, P]`` upload arena (bank-bound module views, no deep copies) and
aggregation is ONE launch of the K1 kernel (``ops.weighted_average``); FedOpt's server step on the pseudo-gradient
``θ_old − θ_avg`` runs fused (``ops.server_opt_step_``) for sgd/adam/adagrad/yogi and through ``OptRepo`` for any
other ``torch.optim`` class.
"""
from __future__ import annotations

import logging
from typing import Dict, List

import numpy as np
import torch
from torch import nn

from .. import ops
from ..models import utils as mutils
from ..parallel.arena import ModelBank
from ..utils.metrics import get_sink


class OptRepo:
    """name → ``torch.optim.Optimizer`` subclass registry (parity: ``optrepo.py:7-65``)."""

    repo = {x.__name__.lower(): x for x in torch.optim.Optimizer.__subclasses__()}

    @classmethod
    def _update_repo(cls) -> None:
        cls.repo = {x.__name__.lower(): x for x in torch.optim.Optimizer.__subclasses__()}

    @classmethod
    def get_opt_names(cls) -> List[str]:
        cls._update_repo()
        return list(cls.repo.keys())

    @classmethod
    def name2cls(cls, name: str):
        try:
            return cls.repo[name.lower()]
        except KeyError:
            cls._update_repo()
            if name.lower() in cls.repo:
                return cls.repo[name.lower()]
            logging.error("Invalid optimizer: %s! registered: %s", name, cls.get_opt_names())
            raise

    @classmethod
    def supported_parameters(cls, opt) -> List[str]:
        opt_ = cls.name2cls(opt) if isinstance(opt, str) else opt
        res = list(opt_.__init__.__code__.co_varnames)
        for k in ("defaults", "self", "params"):
            if k in res:
                res.remove(k)
        return res


class Client:
    """One simulated client: full-epoch local training + local evaluation (parity: ``standalone/fedavg/client.py``).
    ``stackoverflow_lr`` is multi-label: BCE loss, exact-match accuracy, precision / recall sums."""

    def __init__(self, client_idx, local_training_data, local_test_data, local_sample_number, args, device):
        self.client_idx = client_idx
        self.local_training_data, self.local_test_data = local_training_data, local_test_data
        self.local_sample_number = local_sample_number
        self.args, self.device = args, device
        self.multilabel = getattr(args, "dataset", "") == "stackoverflow_lr"
        self.criterion = (nn.BCELoss(reduction="sum") if self.multilabel else nn.CrossEntropyLoss()).to(device)

    def update_local_dataset(self, client_idx, local_training_data, local_test_data, local_sample_number):
        self.client_idx = client_idx
        self.local_training_data, self.local_test_data = local_training_data, local_test_data
        self.local_sample_number = local_sample_number

    def get_sample_number(self):
        return self.local_sample_number

    def train(self, net: nn.Module):
        """Trains ``net`` IN PLACE (its parameters may be arena-row views) and returns (state_dict, mean loss)."""
        net.train()
        if self.args.client_optimizer == "sgd":
            opt = torch.optim.SGD(net.parameters(), lr=self.args.lr)
        else:
            opt = torch.optim.Adam(filter(lambda p: p.requires_grad, net.parameters()), lr=self.args.lr,
                                   weight_decay=self.args.wd, amsgrad=True)
        epoch_loss = []
        for _ in range(self.args.epochs):
            acc = torch.zeros((), device=self.device)
            nb = 0
            for x, labels in self.local_training_data:
                x, labels = x.to(self.device), labels.to(self.device)
                net.zero_grad()
                loss = self.criterion(net(x), labels)
                loss.backward()
                opt.step()
                acc += loss.detach()
                nb += 1
            epoch_loss.append(float(acc) / max(nb, 1))  # one host sync per epoch, not per batch
        return net.state_dict(), sum(epoch_loss) / max(len(epoch_loss), 1)

    def local_test(self, model_global: nn.Module, b_use_test_dataset: bool = False) -> Dict[str, float]:
        model_global.eval()
        m = {"test_correct": 0.0, "test_loss": 0.0, "test_precision": 0.0, "test_recall": 0.0, "test_total": 0.0}
        data = self.local_test_data if b_use_test_dataset else self.local_training_data
        acc = torch.zeros(3, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for x, target in (data or []):
                x, target = x.to(self.device), target.to(self.device)
                pred = model_global(x)
                if self.multilabel:
                    predicted = (pred > .5).int()
                    m["test_correct"] += float(predicted.eq(target).sum(-1).eq(target.size(1)).sum())
                    tp = ((target * predicted) > .1).int().sum(-1)
                    m["test_precision"] += float((tp / (predicted.sum(-1) + 1e-13)).sum())
                    m["test_recall"] += float((tp / (target.sum(-1) + 1e-13)).sum())
                    m["test_loss"] += float(self.criterion(pred, target.float())) * 1.0
                    m["test_total"] += target.size(0)
                else:
                    ops.eval_logits(pred, target, acc)
        if not self.multilabel:
            a = acc.tolist()
            m["test_correct"], m["test_loss"], m["test_total"] = a[0], a[1], a[2]
        return m


class FedAvgTrainer:
    """``FedAvgTrainer(dataset, model, device, args).train()`` (parity: ``fedavg_trainer.py:10-198``).
    ``dataset`` is the FedML 8-tuple ``[train_num, test_num, train_global, test_global, local_num_dict,
    train_local_dict, test_local_dict, class_num]``."""

    def __init__(self, dataset, model, device, args):
        self.device, self.args = torch.device(device), args
        [self.train_data_num_in_total, self.test_data_num_in_total, self.train_global, self.test_global,
         self.train_data_local_num_dict, self.train_data_local_dict, self.test_data_local_dict, self.class_num] = dataset[:8]
        self.sink = get_sink()
        self.bank = ModelBank(model, 1 + args.client_num_per_round, self.device)   # row 0 = global, rows 1.. = clients
        self.bank.load_state_dict(0, model.state_dict())
        self.model_global = self.bank.module(0)
        self.client_list = [Client(i, self.train_data_local_dict.get(i), self.test_data_local_dict.get(i),
                                   self.train_data_local_num_dict.get(i, 0), args, self.device)
                            for i in range(args.client_num_per_round)]
        self.weight_mask = mutils.weight_param_mask(self.bank.spec).to(self.device)

    def client_sampling(self, round_idx, client_num_in_total, client_num_per_round):
        if client_num_in_total == client_num_per_round:
            return list(range(client_num_in_total))
        np.random.seed(round_idx)
        return np.random.choice(range(client_num_in_total), min(client_num_per_round, client_num_in_total), replace=False)

    def _local_round(self, client_indexes, base_row: int = 0):
        """Every sampled client trains a copy of row ``base_row`` inside its own arena row."""
        ns, losses = [], []
        for slot, client in enumerate(self.client_list[: len(client_indexes)]):
            ci = int(client_indexes[slot])
            client.update_local_dataset(ci, self.train_data_local_dict.get(ci), self.test_data_local_dict.get(ci),
                                        self.train_data_local_num_dict.get(ci, 0))
            self.bank.copy(1 + slot, base_row)
            _, loss = client.train(self.bank.module(1 + slot))
            ns.append(float(client.get_sample_number()))
            losses.append(loss)
        return ns, losses

    def aggregate_rows(self, rows: torch.Tensor, ns: List[float]) -> torch.Tensor:
        return ops.weighted_average(rows, torch.tensor(ns, dtype=torch.float32, device=self.device))

    def server_update(self, avg: torch.Tensor, round_idx: int) -> None:
        self.bank.theta[0].copy_(avg)

    def train(self):
        a = self.args
        for round_idx in range(a.comm_round):
            idx = self.client_sampling(round_idx, a.client_num_in_total, a.client_num_per_round)
            ns, losses = self._local_round(idx)
            n = len(ns)
            avg = self.aggregate_rows(self.bank.theta[1:1 + n], ns)
            self.server_update(avg, round_idx)
            self.sink.log({"Train/LocalLoss": sum(losses) / max(len(losses), 1), "round": round_idx})
            if round_idx % a.frequency_of_the_test == 0 or round_idx == a.comm_round - 1:
                self.local_test_on_all_clients(self.model_global, round_idx)
        return self.model_global

    def local_test_on_all_clients(self, model_global, round_idx):
        tr = {"num_samples": 0.0, "num_correct": 0.0, "precisions": 0.0, "recalls": 0.0, "losses": 0.0}
        te = dict(tr)
        client = self.client_list[0]
        for ci in range(self.args.client_num_in_total):
            if self.test_data_local_dict.get(ci) is None and self.train_data_local_dict.get(ci) is None:
                continue
            client.update_local_dataset(0, self.train_data_local_dict.get(ci), self.test_data_local_dict.get(ci),
                                        self.train_data_local_num_dict.get(ci, 0))
            for agg, use_test in ((tr, False), (te, True)):
                m = client.local_test(model_global, use_test)
                agg["num_samples"] += m["test_total"]; agg["num_correct"] += m["test_correct"]
                agg["losses"] += m["test_loss"]; agg["precisions"] += m["test_precision"]; agg["recalls"] += m["test_recall"]
            if getattr(self.args, "ci", 0) == 1:
                break
        out = {"Train/Acc": tr["num_correct"] / max(tr["num_samples"], 1), "Train/Loss": tr["losses"] / max(tr["num_samples"], 1),
               "Test/Acc": te["num_correct"] / max(te["num_samples"], 1), "Test/Loss": te["losses"] / max(te["num_samples"], 1)}
        if getattr(self.args, "dataset", "") == "stackoverflow_lr":
            out.update({"Train/Pre": tr["precisions"] / max(tr["num_samples"], 1), "Train/Rec": tr["recalls"] / max(tr["num_samples"], 1),
                        "Test/Pre": te["precisions"] / max(te["num_samples"], 1), "Test/Rec": te["recalls"] / max(te["num_samples"], 1)})
        for k, v in out.items():
            self.sink.log({k: v, "round": round_idx})
        return out


class FedOptTrainer(FedAvgTrainer):
    """FedAvg + server optimizer on the pseudo-gradient (parity: ``fedopt_trainer.py``; Reddi et al. 2020)."""

    FUSED = ("sgd", "adam", "adagrad", "yogi")

    def __init__(self, dataset, model, device, args):
        super().__init__(dataset, model, device, args)
        self.opt_name = getattr(args, "server_optimizer", "sgd").lower()
        self.server_lr = float(getattr(args, "server_lr", 1.0))
        self.server_momentum = float(getattr(args, "server_momentum", 0.0))
        self.opt_state: Dict = {}
        self.opt = None
        if self.opt_name not in self.FUSED:
            self.opt = OptRepo.name2cls(self.opt_name)(self.model_global.parameters(), lr=self.server_lr)

    def server_update(self, avg: torch.Tensor, round_idx: int) -> None:
        theta = self.bank.theta[0]
        if self.opt is None:
            kw = {"momentum": self.server_momentum} if self.opt_name == "sgd" else {}
            # BN statistics are not optimiser state: they follow the average directly
            stats = ~self.weight_mask
            keep = avg[stats].clone() if bool(stats.any()) else None
            ops.server_opt_step_(theta, avg, self.opt_state, self.opt_name, self.server_lr, **kw)
            if keep is not None:
                theta[stats] = keep
            return
        self.opt.zero_grad()
        views = mutils.unflatten_to_state_dict(avg, self.bank.spec)
        with torch.no_grad():
            for name, p in self.model_global.named_parameters():
                p.grad = p.data - views[name]
        self.opt.step()
        with torch.no_grad():
            for name, b in self.model_global.named_buffers():
                if name in views:
                    b.copy_(views[name])


class HierarchicalTrainer(FedAvgTrainer):
    """Two-level FL: clients → random groups → global (parity: ``hierarchical_fl/trainer.py:33-116``)."""

    def __init__(self, dataset, model, device, args):
        super().__init__(dataset, model, device, args)
        if getattr(args, "group_method", "random") != "random":
            raise Exception(args.group_method)
        self.group_indexes = np.random.randint(0, args.group_num, args.client_num_in_total)
        self.group_bank = ModelBank(model, args.group_num, self.device)

    def group_sampling(self, global_round_idx):
        idx = self.client_sampling(global_round_idx, self.args.client_num_in_total, self.args.client_num_per_round)
        groups: Dict[int, List[int]] = {}
        for ci in idx:
            groups.setdefault(int(self.group_indexes[ci]), []).append(int(ci))
        return groups

    def train(self):
        a = self.args
        for g_round in range(a.global_comm_round):
            groups = self.group_sampling(g_round)
            for g in range(a.group_num):
                self.group_bank.theta[g].copy_(self.bank.theta[0])
            for grp_round in range(a.group_comm_round):
                all_rows, all_ns = [], []
                for g, members in groups.items():
                    self.bank.theta[0].copy_(self.group_bank.theta[g])
                    # reuse the client slots (members may exceed the slot count → chunk)
                    rows, ns = [], []
                    for off in range(0, len(members), len(self.client_list)):
                        chunk = members[off:off + len(self.client_list)]
                        n_, _ = self._local_round(chunk)
                        rows.append(self.bank.theta[1:1 + len(chunk)].clone())
                        ns += n_
                    rows = torch.cat(rows)
                    self.group_bank.theta[g].copy_(self.aggregate_rows(rows, ns))
                    all_rows.append(rows)
                    all_ns += ns
                w_glob = self.aggregate_rows(torch.cat(all_rows), all_ns)
                gi = a.group_comm_round * g_round + grp_round
                if gi % a.frequency_of_the_test == 0 or g_round == a.global_comm_round - 1:
                    self.bank.theta[0].copy_(w_glob)
                    self.local_test_on_all_clients(self.model_global, gi)
            self.bank.theta[0].copy_(w_glob)
        return self.model_global
