"""Model-split federated methods: SplitNN, FedGKT, classical vertical FL (SURVEY §2.4 / §2.10 "PP-like").

Parity: ``fedml_api/distributed/split_nn/{client,server,client_manager,server_manager}.py``,
``fedml_api/distributed/fedgkt/{GKTClientTrainer,GKTServerTrainer,GKT*Manager,utils}.py``,
``fedml_api/distributed/classical_vertical_fl/{guest_trainer,host_trainer,*_manager}.py`` and
``fedml_api/standalone/classical_vertical_fl``.

The reference moves activations / feature maps / logits between processes as pickled **numpy** dictionaries (FedGKT
ships every batch's ``[B,16,32,32]`` feature map for the whole train and test set each round — its README warns of
256 GB host RAM).  Here the two stages live on (possibly different) CUDA devices and exchange **device tensors**
(`.to(peer_device, non_blocking=True)` = one NVLink peer copy on a multi-GPU box, a no-op on one GPU); the FedGKT
distillation loss is the fused ``ops.kd_kl_loss`` kernel (K14) and the vertical-FL logit sum + BCE gradient is
``ops.vfl_bce_grad`` (K15).  The message-level protocol (who sends what, in which order) is preserved.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from .. import ops
from ..utils.metrics import get_sink


# ====================================================================================== SplitNN
class SplitNNClient:
    """Holds the bottom ``split_layer`` children of the model (parity ``split_nn/client.py:9-40``)."""

    def __init__(self, model: nn.Module, trainloader, testloader, rank: int, device, epochs: int = 1, lr: float = 0.1):
        self.model, self.rank, self.device = model.to(device), rank, device
        self.trainloader, self.testloader, self.MAX_EPOCH_PER_NODE = trainloader, testloader, epochs
        self.optimizer = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9, weight_decay=5e-4)
        self.dataloader = iter(trainloader)
        self.acts = None

    def forward_pass(self):
        inputs, labels = next(self.dataloader)
        inputs, labels = inputs.to(self.device), labels.to(self.device)
        self.optimizer.zero_grad()
        self.acts = self.model(inputs)
        return self.acts, labels

    def backward_pass(self, grads):
        self.acts.backward(grads.to(self.acts.device))
        self.optimizer.step()

    def train_mode(self):
        self.dataloader = iter(self.trainloader)
        self.model.train()

    def eval_mode(self):
        self.dataloader = iter(self.testloader)
        self.model.eval()


class SplitNNServer:
    """Holds the top part; returns ``acts.grad`` to the active client (parity ``split_nn/server.py:10-75``)."""

    def __init__(self, model: nn.Module, device, max_rank: int, lr: float = 0.1):
        self.model, self.device, self.MAX_RANK = model.to(device), device, max_rank
        self.optimizer = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9, weight_decay=5e-4)
        self.criterion = nn.CrossEntropyLoss()
        self.epoch, self.active_node = 0, 1
        self.stats = torch.zeros(3, device=device)  # correct, loss_sum, count — device accumulated
        self.phase = "train"

    def reset_local_params(self):
        self.stats.zero_()

    def train_mode(self):
        self.model.train()
        self.phase = "train"
        self.reset_local_params()

    def eval_mode(self):
        self.model.eval()
        self.phase = "validation"
        self.reset_local_params()

    def forward_pass(self, acts: torch.Tensor, labels: torch.Tensor):
        self.acts = acts.detach().to(self.device, non_blocking=True).requires_grad_(True)   # stage boundary
        labels = labels.to(self.device, non_blocking=True)
        self.optimizer.zero_grad()
        logits = self.model(self.acts)
        self.loss = self.criterion(logits, labels)
        ops.eval_logits(logits.detach(), labels, self.stats)
        return logits

    def backward_pass(self):
        self.loss.backward()
        self.optimizer.step()
        return self.acts.grad

    def validation_over(self) -> Dict[str, float]:
        c, l, n = self.stats.tolist()
        out = {"acc": c / max(n, 1), "loss": l / max(n, 1)}
        self.active_node = (self.active_node % self.MAX_RANK) + 1
        self.epoch += 1
        self.train_mode()
        return out


def split_model(model: nn.Module, split_layer: int = 1):
    """Cut after the first ``split_layer`` children (``main_split_nn.py:128-139``)."""
    kids = list(model.children())
    return nn.Sequential(*kids[:split_layer]), nn.Sequential(*kids[split_layer:])


def SplitNN_distributed(client_models: List[nn.Module], server_model: nn.Module, loaders, device, epochs: int = 1,
                        server_device=None, lr: float = 0.1) -> List[Dict[str, float]]:
    """Round-robin relay: client k trains ``epochs`` passes against the shared server half, validates, then hands the
    semaphore to client k+1 (``client_manager.py:40-55``).  ``loaders[k] = (train_loader, test_loader)``."""
    server_device = server_device or device
    server = SplitNNServer(server_model, server_device, len(client_models), lr)
    clients = [SplitNNClient(m, tr, te, k + 1, device, epochs, lr) for k, (m, (tr, te)) in enumerate(zip(client_models, loaders))]
    sink, results = get_sink(), []
    for k, cl in enumerate(clients):
        for _ in range(cl.MAX_EPOCH_PER_NODE):
            cl.train_mode(); server.train_mode()
            for _ in range(len(cl.trainloader)):
                acts, labels = cl.forward_pass()
                server.forward_pass(acts, labels)
                cl.backward_pass(server.backward_pass())
            cl.eval_mode(); server.eval_mode()
            with torch.no_grad():
                for _ in range(len(cl.testloader)):
                    acts, labels = cl.forward_pass()
                    logits = server.model(acts.to(server_device))
                    ops.eval_logits(logits, labels.to(server_device), server.stats)
            r = server.validation_over()
            r["client"] = k + 1
            results.append(r)
            sink.log({"SplitNN/Val-Acc": r["acc"], "SplitNN/Val-Loss": r["loss"], "client": k + 1})
        if k + 1 < len(clients):  # the next client continues from this client's bottom weights
            clients[k + 1].model.load_state_dict(cl.model.state_dict())
    return results


# ====================================================================================== FedGKT
class KL_Loss(nn.Module):
    """T²·KL(softmax(teacher/T)+1e-7 ‖ softmax(student/T)), batch-mean (parity ``fedgkt/utils.py:75-94``)."""

    def __init__(self, temperature: float = 1.0):
        super().__init__()
        self.T = temperature

    def forward(self, output_batch, teacher_outputs):
        return ops.kd_kl_loss(output_batch, teacher_outputs, self.T)


class GKTClientTrainer:
    """Edge side: trains the small model with CE (+ α·KD against the server's logits), then extracts feature maps,
    logits and labels for every local batch (``GKTClientTrainer.py:40-129``)."""

    def __init__(self, client_index, local_training_data, local_test_data, device, client_model, args):
        self.client_index, self.device, self.args = client_index, device, args
        self.local_training_data, self.local_test_data = local_training_data, local_test_data
        self.client_model = client_model.to(device)
        if getattr(args, "optimizer", "SGD") == "SGD":
            self.optimizer = torch.optim.SGD(self.client_model.parameters(), lr=args.lr, momentum=0.9, nesterov=True,
                                             weight_decay=args.wd)
        else:
            self.optimizer = torch.optim.Adam(self.client_model.parameters(), lr=args.lr, weight_decay=args.wd, amsgrad=True)
        self.criterion_CE, self.criterion_KL = nn.CrossEntropyLoss(), KL_Loss(args.temperature)
        self.server_logits_dict: Dict[int, torch.Tensor] = {}

    def get_sample_number(self):
        return sum(int(y.shape[0]) for _, y in self.local_training_data)

    def update_large_model_logits(self, logits: Dict[int, torch.Tensor]):
        self.server_logits_dict = logits

    def train(self):
        a = self.args
        if getattr(a, "whether_training_on_client", 1) == 1:
            self.client_model.train()
            for _ in range(a.epochs_client):
                for b, (x, y) in enumerate(self.local_training_data):
                    x, y = x.to(self.device), y.to(self.device)
                    log_probs, _ = self.client_model(x)
                    loss = self.criterion_CE(log_probs, y)
                    if len(self.server_logits_dict) != 0:
                        loss = loss + a.alpha * self.criterion_KL(log_probs, self.server_logits_dict[b].to(self.device))
                    self.optimizer.zero_grad()
                    loss.backward()
                    self.optimizer.step()
        self.client_model.eval()
        feats, logits, labels, feats_te, labels_te = {}, {}, {}, {}, {}
        with torch.no_grad():
            for b, (x, y) in enumerate(self.local_training_data):
                lp, f = self.client_model(x.to(self.device))
                feats[b], logits[b], labels[b] = f, lp, y.to(self.device)      # device tensors, no numpy
            for b, (x, y) in enumerate(self.local_test_data):
                _, f = self.client_model(x.to(self.device))
                feats_te[b], labels_te[b] = f, y.to(self.device)
        return feats, logits, labels, feats_te, labels_te


class GKTServerTrainer:
    """Server side: trains the large model on the clients' feature maps with KD + α·CE, returns per-batch logits
    (``GKTServerTrainer.py:14-325``); ReduceLROnPlateau on the test accuracy; best/last checkpoints are kept in memory and,
    like the reference (``GKTServerTrainer.py:213-231``), written to ``<checkpoint_dir>/{last,best}.pth`` (model + optimizer
    + epoch + accuracies) and ``test_best_metrics.json`` — ``args.checkpoint_dir`` defaults to ``./checkpoint``; ``None``/``""``
    disables the files.  ``resume(path)`` restores model, optimizer and best accuracy."""

    def __init__(self, client_num, device, server_model, args):
        self.client_num, self.device, self.args = client_num, device, args
        self.model_global = server_model.to(device)
        if getattr(args, "optimizer", "SGD") == "SGD":
            self.optimizer = torch.optim.SGD(self.model_global.parameters(), lr=args.lr, momentum=0.9, nesterov=True,
                                             weight_decay=args.wd)
        else:
            self.optimizer = torch.optim.Adam(self.model_global.parameters(), lr=args.lr, weight_decay=args.wd, amsgrad=True)
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, "max")
        self.criterion_CE, self.criterion_KL = nn.CrossEntropyLoss(), KL_Loss(args.temperature)
        self.best_acc, self.checkpoints = 0.0, {}
        self.client_feats, self.client_logits, self.client_labels, self.client_feats_te, self.client_labels_te = {}, {}, {}, {}, {}
        self.server_logits_dict: Dict[int, Dict[int, torch.Tensor]] = {}
        self.flags = {i: False for i in range(client_num)}

    def add_local_trained_result(self, index, feats, logits, labels, feats_te, labels_te):
        self.client_feats[index], self.client_logits[index], self.client_labels[index] = feats, logits, labels
        self.client_feats_te[index], self.client_labels_te[index] = feats_te, labels_te
        self.flags[index] = True

    def check_whether_all_receive(self):
        if not all(self.flags.values()):
            return False
        self.flags = {i: False for i in range(self.client_num)}
        return True

    def get_global_logits(self, client_index):
        return self.server_logits_dict[client_index]

    def train(self, round_idx):
        a = self.args
        for _ in range(self.get_server_epoch_strategy(round_idx)):
            m = self.train_large_model_on_the_server()
        ev = self.eval_large_model_on_the_server()
        self.scheduler.step(ev["test_accTop1"])
        self.checkpoints["last"] = copy.deepcopy(self.model_global.state_dict())
        is_best = ev["test_accTop1"] >= self.best_acc
        if is_best:
            self.best_acc = ev["test_accTop1"]
            self.checkpoints["best"] = self.checkpoints["last"]
        self._save_files(round_idx, ev, is_best)
        sink = get_sink()
        sink.log({"Train/Loss": m["train_loss"], "Train/AccTop1": m["train_accTop1"], "Test/AccTop1": ev["test_accTop1"],
                  "Test/Loss": ev["test_loss"], "round": round_idx})
        return m, ev

    def _save_files(self, round_idx: int, ev: Dict, is_best: bool) -> None:
        import json
        import os
        import shutil
        cdir = getattr(self.args, "checkpoint_dir", "./checkpoint")
        if not cdir:
            return
        os.makedirs(cdir, exist_ok=True)
        last = os.path.join(cdir, "last.pth")
        tmp = last + ".tmp"
        torch.save({"state_dict": {k: v.detach().cpu() for k, v in self.model_global.state_dict().items()},
                    "optim_dict": self.optimizer.state_dict(), "epoch": round_idx + 1,
                    "test_accTop1": float(ev["test_accTop1"]), "test_accTop5": float(ev.get("test_accTop5", 0.0))}, tmp)
        os.replace(tmp, last)   # atomic: a crash never leaves a torn checkpoint
        if is_best:
            best = {k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in ev.items()}
            best["epoch"] = round_idx + 1
            with open(os.path.join(cdir, "test_best_metrics.json"), "w") as fh:
                json.dump(best, fh, indent=4)
            shutil.copyfile(last, os.path.join(cdir, "best.pth"))

    def resume(self, path: str) -> int:
        """Restore model / optimizer / best accuracy from ``last.pth`` or ``best.pth``; returns the next round index."""
        blob = torch.load(path, map_location=self.device, weights_only=True)
        self.model_global.load_state_dict(blob["state_dict"])
        self.optimizer.load_state_dict(blob["optim_dict"])
        self.best_acc = max(self.best_acc, float(blob.get("test_accTop1", 0.0)))
        return int(blob.get("epoch", 0))

    def get_server_epoch_strategy(self, round_idx):
        return int(getattr(self.args, "epochs_server", 1))

    def train_large_model_on_the_server(self):
        self.server_logits_dict = {}
        self.model_global.train()
        stats = torch.zeros(3, device=self.device)
        loss_sum, nb = torch.zeros((), device=self.device), 0
        for ci, feats in self.client_feats.items():
            out_logits = self.server_logits_dict.setdefault(ci, {})
            for b, f in feats.items():
                f = f.to(self.device, non_blocking=True)
                y = self.client_labels[ci][b].to(self.device).long()
                out = self.model_global(f)
                if getattr(self.args, "whether_distill_on_the_server", 1) == 1:
                    loss = self.criterion_KL(out, self.client_logits[ci][b].to(self.device).float()) \
                        + self.args.alpha * self.criterion_CE(out, y)
                else:
                    loss = self.criterion_CE(out, y)
                self.optimizer.zero_grad()
                loss.backward()
                self.optimizer.step()
                ops.eval_logits(out.detach(), y, stats)
                loss_sum += loss.detach()
                nb += 1
                out_logits[b] = out.detach()
        c, _, n = stats.tolist()
        return {"train_loss": float(loss_sum) / max(nb, 1), "train_accTop1": 100.0 * c / max(n, 1)}

    def eval_large_model_on_the_server(self):
        self.model_global.eval()
        stats = torch.zeros(3, device=self.device)
        with torch.no_grad():
            for ci, feats in self.client_feats_te.items():
                for b, f in feats.items():
                    ops.eval_logits(self.model_global(f.to(self.device)), self.client_labels_te[ci][b].to(self.device), stats)
        c, l, n = stats.tolist()
        return {"test_loss": l / max(n, 1), "test_accTop1": 100.0 * c / max(n, 1)}


def FedML_FedGKT_distributed(client_models, server_model, client_loaders, device, args, server_device=None):
    """Synchronous GKT rounds: all clients train+extract → server trains on all features → logits go back
    (``GKTServerManager.py:36-60`` / ``GKTClientManager.py:30-60``)."""
    server_device = server_device or device
    server = GKTServerTrainer(len(client_models), server_device, server_model, args)
    clients = [GKTClientTrainer(i, tr, te, device, m, args) for i, (m, (tr, te)) in enumerate(zip(client_models, client_loaders))]
    hist = []
    for r in range(args.comm_round):
        for i, cl in enumerate(clients):
            server.add_local_trained_result(i, *cl.train())
        assert server.check_whether_all_receive()
        hist.append(server.train(r))
        for i, cl in enumerate(clients):
            cl.update_large_model_logits(server.get_global_logits(i))
    return server, hist


# ====================================================================================== classical vertical FL
class VFLHostTrainer:
    """A feature-holding party without labels: sends ``[B,1]`` logits, receives ``∂L/∂logit`` (``host_trainer.py``)."""

    def __init__(self, client_index, device, X_train, X_test, model_feature_extractor, model_classifier, args):
        self.client_index, self.device, self.args = client_index, device, args
        self.X_train = torch.as_tensor(np.asarray(X_train), dtype=torch.float32, device=device)
        self.X_test = torch.as_tensor(np.asarray(X_test), dtype=torch.float32, device=device)
        self.fe, self.clf = model_feature_extractor.to(device), model_classifier.to(device)
        params = list(self.fe.parameters()) + list(self.clf.parameters())
        self.optimizer = torch.optim.SGD(params, momentum=0.9, weight_decay=0.01, lr=args.lr)
        self.batch_size = args.batch_size
        self.n_batches = -(-self.X_train.shape[0] // self.batch_size)
        self.batch_idx, self.cached = 0, None

    def get_batch_num(self):
        return self.n_batches

    def computer_logits(self, round_idx):
        b = self.batch_idx
        x = self.X_train[b * self.batch_size:(b + 1) * self.batch_size]
        self.cached = self.clf(self.fe(x))
        self.batch_idx = (self.batch_idx + 1) % self.n_batches
        test_logits = None
        if (round_idx + 1) % getattr(self.args, "frequency_of_the_test", 1) == 0:
            with torch.no_grad():
                test_logits = self.clf(self.fe(self.X_test))
        return self.cached.detach(), test_logits

    def update_model(self, gradient):
        self.optimizer.zero_grad()
        self.cached.backward(gradient.to(self.device))
        self.optimizer.step()


class VFLGuestTrainer(VFLHostTrainer):
    """The label-holding party: sums all parties' logits, BCE-with-logits, returns the common gradient
    (``guest_trainer.py:73-111``) — the sum + loss + gradient is ONE fused kernel (K15)."""

    def __init__(self, client_num, device, X_train, y_train, X_test, y_test, model_feature_extractor, model_classifier, args):
        super().__init__(0, device, X_train, X_test, model_feature_extractor, model_classifier, args)
        self.client_num = client_num
        self.y_train = torch.as_tensor(np.asarray(y_train), dtype=torch.float32, device=device).reshape(-1, 1)
        self.y_test = torch.as_tensor(np.asarray(y_test), dtype=torch.float32, device=device).reshape(-1, 1)
        self.host_logits: Dict[int, torch.Tensor] = {}
        self.host_test_logits: Dict[int, Optional[torch.Tensor]] = {}
        self.loss_list: List[float] = []

    def add_client_local_result(self, index, train_logits, test_logits):
        self.host_logits[index], self.host_test_logits[index] = train_logits, test_logits

    def check_whether_all_receive(self):
        return len(self.host_logits) == self.client_num - 1

    def train(self, round_idx):
        b = self.batch_idx
        y = self.y_train[b * self.batch_size:(b + 1) * self.batch_size]
        own, own_test = self.computer_logits(round_idx)
        parts = torch.stack([own] + [self.host_logits[i].to(self.device) for i in sorted(self.host_logits)])  # [K,B,1]
        loss, grad = ops.vfl_bce_grad(parts, y)
        self.update_model(grad)
        self.loss_list.append(float(loss))
        metrics = None
        if own_test is not None and all(v is not None for v in self.host_test_logits.values()):
            z = own_test + sum(self.host_test_logits[i].to(self.device) for i in self.host_test_logits)
            p = torch.sigmoid(z)
            pred = (p > 0.5).float()
            acc = float((pred == self.y_test).float().mean())
            metrics = {"test_acc": acc, "test_auc": _auc(p.flatten().cpu().numpy(), self.y_test.flatten().cpu().numpy()),
                       "loss": float(np.mean(self.loss_list))}
            get_sink().log({"VFL/Test-Acc": acc, "VFL/Test-AUC": metrics["test_auc"], "VFL/Loss": metrics["loss"], "round": round_idx})
            self.loss_list = []
        self.host_logits, self.host_test_logits = {}, {}
        return grad, metrics


def _auc(score: np.ndarray, y: np.ndarray) -> float:
    order = np.argsort(score)
    ranks = np.empty_like(order, dtype=np.float64)
    ranks[order] = np.arange(1, len(score) + 1)
    pos = y > 0.5
    n1, n0 = pos.sum(), (~pos).sum()
    if n1 == 0 or n0 == 0:
        return 0.5
    return float((ranks[pos].sum() - n1 * (n1 + 1) / 2) / (n1 * n0))


def FedML_VFL_distributed(guest: VFLGuestTrainer, hosts: List[VFLHostTrainer], comm_round: int):
    """Guest ↔ hosts protocol, one mini-batch per step; total steps = ``comm_round × n_batches``
    (``guest_manager.py:41``).  Returns the metric history."""
    hist = []
    for r in range(comm_round * guest.get_batch_num()):
        for h in hosts:
            guest.add_client_local_result(h.client_index, *h.computer_logits(r))
        assert guest.check_whether_all_receive()
        grad, m = guest.train(r)
        for h in hosts:
            h.update_model(grad)
        if m is not None:
            hist.append(m)
    return hist


class VerticalMultiplePartyLogisticRegressionFederatedLearning:
    """Standalone VFL (``standalone/classical_vertical_fl/vfl.py``): party A holds labels, others features only."""

    def __init__(self, party_A, main_party_id="_main"):
        self.main_party_id, self.party_a, self.party_dict, self.is_debug = main_party_id, party_A, {}, False

    def add_party(self, *, id, party_model):
        self.party_dict[id] = party_model

    def get_main_party_id(self):
        return self.main_party_id

    def fit(self, X_A, y, party_X_dict, global_step):
        self.party_a.set_batch(X_A, y, global_step)
        for pid, X in party_X_dict.items():
            self.party_dict[pid].set_batch(X, global_step)
        comp = {pid: p.send_components() for pid, p in self.party_dict.items()}
        self.party_a.receive_components(list(comp.values()))
        self.party_a.fit()
        grads = self.party_a.send_gradients()
        for p in self.party_dict.values():
            p.receive_gradients(grads)
        return self.party_a.get_loss()

    def predict(self, X_A, party_X_dict):
        comps = [self.party_dict[pid].predict(X) for pid, X in party_X_dict.items()]
        return self.party_a.predict(X_A, component_list=comps)


class VFLGuestModel:
    """Party A of the standalone simulator (numpy in/out ``DenseModel``/``LocalModel``; ``party_models.py:12-76``)."""

    def __init__(self, local_model, dense_model_dim=None, learning_rate=0.01, optimizer="sgd"):
        from ..models.vfl import DenseModel
        self.localModel = local_model
        self.feature_dim = local_model.get_output_dim()
        self.dense_model = DenseModel(self.feature_dim, 1, learning_rate, bias=True)
        self.parties_grad_component_list, self.current_global_step = [], None
        self.X = self.y = None

    def set_batch(self, X, y, global_step):
        self.X, self.y, self.current_global_step = X, np.asarray(y, dtype=np.float32).reshape(-1, 1), global_step

    def _fit(self, X, y):
        self.temp_K_Z = self.localModel.forward(X)
        self.K_U = self.dense_model.forward(self.temp_K_Z)
        parts = torch.as_tensor(np.stack([self.K_U] + self.parties_grad_component_list), dtype=torch.float32)
        loss, grad = ops.vfl_bce_grad(parts, torch.as_tensor(y))
        self.loss, self.top_grads = float(loss), grad.numpy()
        back = self.dense_model.backward(self.temp_K_Z, self.top_grads)
        self.localModel.backward(X, back)

    def receive_components(self, component_list):
        self.parties_grad_component_list = [np.asarray(c) for c in component_list]

    def fit(self):
        self._fit(self.X, self.y)
        self.parties_grad_component_list = []

    def send_gradients(self):
        return self.top_grads

    def get_loss(self):
        return self.loss

    def predict(self, X, component_list):
        U = self.dense_model.forward(self.localModel.forward(X)) + sum(np.asarray(c) for c in component_list)
        return 1.0 / (1.0 + np.exp(-U))


class VFLHostModel:
    def __init__(self, local_model, learning_rate=0.01):
        from ..models.vfl import DenseModel
        self.localModel = local_model
        self.dense_model = DenseModel(local_model.get_output_dim(), 1, learning_rate, bias=False)
        self.X = None

    def set_batch(self, X, global_step):
        self.X = X

    def _forward(self, X):
        self.A_Z = self.localModel.forward(X)
        return self.dense_model.forward(self.A_Z)

    def send_components(self):
        return self._forward(self.X)

    def receive_gradients(self, gradients):
        back = self.dense_model.backward(self.A_Z, gradients)
        self.localModel.backward(self.X, back)

    def predict(self, X):
        return self._forward(X)
