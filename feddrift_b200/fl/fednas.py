"""FedNAS: federated DARTS search — clients alternate MiLeNAS α steps and weight SGD, the server averages weights AND α
(``fedml_api/distributed/fednas/{FedNASTrainer,FedNASAggregator,FedNAS*Manager}.py``).

Because α are parameters of the search network (``models/darts.py``) they sit in the same arena row as the weights:
the federated average of both is ONE ``ops.weighted_average`` launch (the reference averages α in a second python loop
that rebinds its loop variable, effectively returning client 0's α — intent reproduced, bug not).
"""
from __future__ import annotations

import copy
from typing import Dict

import torch
from torch import nn

from .. import ops
from ..models.darts import Architect, Network
from ..parallel.arena import ModelBank
from ..utils.metrics import get_sink


class FedNASTrainer:
    def __init__(self, client_index, train_local, test_local, local_sample_number, device, model: Network, args):
        self.client_index, self.device, self.args = client_index, device, args
        self.train_local, self.test_local, self.local_sample_number = train_local, test_local, local_sample_number
        self.model = model.to(device)
        self.criterion = nn.CrossEntropyLoss().to(device)

    def update_model(self, state_dict):
        self.model.load_state_dict(state_dict)

    def search(self):
        """One local search epoch per ``args.epochs``: for every train batch take a validation batch, α step (step_v2),
        then a weight step (``FedNASTrainer.py:82-127``).  Returns (state_dict incl. α, sample count, acc, loss)."""
        a, m = self.args, self.model
        m.train()
        w_opt = torch.optim.SGD(m.weight_parameters(), getattr(a, "learning_rate", 0.025), momentum=getattr(a, "momentum", 0.9),
                                weight_decay=getattr(a, "weight_decay", 3e-4))
        architect = Architect(m, self.criterion, a, self.device)
        stats = torch.zeros(3, device=self.device)
        for _ in range(a.epochs):
            valid_iter = iter(self.test_local)
            for x, y in self.train_local:
                x, y = x.to(self.device), y.to(self.device)
                try:
                    xv, yv = next(valid_iter)
                except StopIteration:
                    valid_iter = iter(self.test_local)
                    xv, yv = next(valid_iter)
                architect.step_v2(x, y, xv.to(self.device), yv.to(self.device))
                w_opt.zero_grad()
                logits = m(x)
                loss = self.criterion(logits, y)
                loss.backward()
                nn.utils.clip_grad_norm_(m.weight_parameters(), getattr(a, "grad_clip", 5.0))
                w_opt.step()
                ops.eval_logits(logits.detach(), y, stats)
        c, l, n = stats.tolist()
        return {k: v.detach().cpu() for k, v in m.state_dict().items()}, self.local_sample_number, c / max(n, 1), l / max(n, 1)

    def infer(self):
        self.model.eval()
        stats = torch.zeros(3, device=self.device)
        with torch.no_grad():
            for x, y in self.test_local:
                ops.eval_logits(self.model(x.to(self.device)), y.to(self.device), stats)
        c, l, n = stats.tolist()
        return c / max(n, 1), l / max(n, 1)


class FedNASAggregator:
    def __init__(self, train_global, test_global, all_train_data_num, client_num, model: Network, device, args):
        self.train_global, self.test_global, self.client_num, self.device, self.args = train_global, test_global, client_num, device, args
        self.bank = ModelBank(model, 1 + client_num, device)
        self.bank.load_state_dict(0, model.state_dict())
        self.model = self.bank.module(0)
        self.sample_num: Dict[int, float] = {}
        self.flags = {i: False for i in range(client_num)}
        self.best_accuracy, self.genotype_history = 0.0, []

    def get_model(self):
        return self.model

    def add_local_trained_result(self, index, model_params, sample_num, train_acc=None, train_loss=None):
        self.bank.load_state_dict(1 + index, model_params)
        self.sample_num[index] = float(sample_num)
        self.flags[index] = True

    def check_whether_all_receive(self):
        if not all(self.flags.values()):
            return False
        self.flags = {i: False for i in range(self.client_num)}
        return True

    def aggregate(self):
        """Weights, BN statistics and α averaged with sample-count weights in one pass over the arena rows."""
        ids = sorted(self.sample_num)
        w = torch.tensor([self.sample_num[i] for i in ids], dtype=torch.float32, device=self.bank.device)
        rows = self.bank.theta[[1 + i for i in ids]]
        avg = ops.weighted_average(rows, w)
        # integer buffers (num_batches_tracked) are copied, not averaged
        fm = self.bank.float_mask.to(avg.device)
        self.bank.theta[0] = torch.where(fm, avg, rows[0])
        return {k: v.detach().cpu().clone() for k, v in self.bank.state_dict(0).items()}

    def statistics(self, round_idx):
        self.model.eval()
        stats = torch.zeros(3, device=self.bank.device)
        with torch.no_grad():
            for x, y in self.test_global:
                ops.eval_logits(self.model(x.to(self.bank.device)), y.to(self.bank.device), stats)
        c, l, n = stats.tolist()
        acc = c / max(n, 1)
        self.best_accuracy = max(self.best_accuracy, acc)
        g = self.model.genotype()
        self.genotype_history.append(g)
        sink = get_sink()
        sink.log({"Test/Acc": acc, "Test/Loss": l / max(n, 1), "round": round_idx})
        sink.log({"genotype": str(g), "round": round_idx})
        sink.set_summary("best_valid_accuracy", self.best_accuracy)
        return acc, g


def FedML_FedNAS_distributed(model: Network, client_loaders, test_global, device, args):
    """Synchronous FedNAS rounds in one process (server ↔ clients protocol of ``FedNASServerManager.py``)."""
    n = len(client_loaders)
    agg = FedNASAggregator(None, test_global, sum(len(tr) for tr, _ in client_loaders), n, model, device, args)
    clients = [FedNASTrainer(i, tr, te, sum(int(y.shape[0]) for _, y in tr), device, copy.deepcopy(model), args)
               for i, (tr, te) in enumerate(client_loaders)]
    hist = []
    for r in range(args.comm_round):
        params = {k: v.detach().cpu().clone() for k, v in agg.bank.state_dict(0).items()}
        for i, cl in enumerate(clients):
            cl.update_model(params)
            sd, ns, acc, loss = cl.search()
            agg.add_local_trained_result(i, sd, ns, acc, loss)
        assert agg.check_whether_all_receive()
        agg.aggregate()
        hist.append(agg.statistics(r))
    return agg, hist
