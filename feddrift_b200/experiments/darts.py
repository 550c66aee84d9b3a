"""Standalone (non-federated) DARTS architecture search and genotype evaluation.

Parity: ``fedml_api/model/cv/darts/train_search.py`` (bi-level search, DARTS / MiLeNAS ``DARTS_V2`` optimisation,
DARTS or GDAS search space, cosine LR, gradient clipping, genotype logging) and ``darts/train.py`` (training of the
discrete network built from a genotype, auxiliary head, drop-path schedule, cutout).  Same flag names.

B200-first differences: the whole CIFAR-sized dataset lives on the device as one uint8/fp32 tensor and batches are
index-gathers (no DataLoader workers, no per-batch H2D); random crop + flip + cutout are batched device ops; the
reference's ``nn.DataParallel`` (``train.py:88``, ``train_search.py:114-119``) is replaced by one process per GPU:
``train`` wraps the network in ``DistributedDataParallel`` (NCCL; gloo on CPU) when launched under torchrun and shards
the sample indices by rank; ``search`` is single-GPU (α steps use ``autograd.grad`` outside DDP's reducer).

    python -m feddrift_b200.experiments.darts search --epochs 2 --layers 5 --init_channels 8
    python -m feddrift_b200.experiments.darts train  --arch FedNAS_V1 --epochs 2 --auxiliary
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..data import benchmarks
from ..models import darts
from ..utils.metrics import MetricsSink


class AvgrageMeter:
    """Running average (name kept from ``darts/utils.py:9-22``)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.avg, self.sum, self.cnt = 0.0, 0.0, 0

    def update(self, val, n=1):
        self.sum += float(val) * n
        self.cnt += n
        self.avg = self.sum / max(self.cnt, 1)


def accuracy(output, target, topk=(1,)):
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * (100.0 / target.size(0)) for k in topk]


def add_args(p: argparse.ArgumentParser) -> argparse.ArgumentParser:
    p.add_argument("mode", choices=["search", "train"])
    p.add_argument("--run_id", type=int, default=0)
    p.add_argument("--data", type=str, default=None, help="CIFAR-10 python-pickle directory; synthetic when absent")
    p.add_argument("--dataset", type=str, default="cifar10")
    p.add_argument("--batch_size", type=int, default=64)
    p.add_argument("--learning_rate", type=float, default=0.025)
    p.add_argument("--learning_rate_min", type=float, default=0.001)
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--weight_decay", type=float, default=3e-4)
    p.add_argument("--report_freq", type=int, default=50)
    p.add_argument("--gpu", type=str, default="0")
    p.add_argument("--epochs", type=int, default=50)
    p.add_argument("--init_channels", type=int, default=16)
    p.add_argument("--layers", type=int, default=8)
    p.add_argument("--model_path", type=str, default="saved_models")
    p.add_argument("--cutout", action="store_true")
    p.add_argument("--cutout_length", type=int, default=16)
    p.add_argument("--drop_path_prob", type=float, default=0.3)
    p.add_argument("--save", type=str, default="EXP")
    p.add_argument("--seed", type=int, default=2)
    p.add_argument("--grad_clip", type=float, default=5.0)
    p.add_argument("--train_portion", type=float, default=0.5)
    p.add_argument("--unrolled", action="store_true")
    p.add_argument("--arch_learning_rate", type=float, default=3e-4)
    p.add_argument("--arch_weight_decay", type=float, default=1e-3)
    p.add_argument("--optimization", type=str, default="DARTS", help="DARTS | DARTS_V2 (MiLeNAS mixed-level)")
    p.add_argument("--arch_search_method", type=str, default="DARTS", help="DARTS | GDAS")
    p.add_argument("--lambda_train_regularizer", type=float, default=1.0)
    p.add_argument("--lambda_valid_regularizer", type=float, default=1.0)
    p.add_argument("--early_stopping", type=int, default=0)
    p.add_argument("--group_id", type=int, default=0)
    p.add_argument("--w_update_times", type=int, default=1)
    p.add_argument("--tau_max", type=float, default=10.0)
    p.add_argument("--tau_min", type=float, default=1.0)
    # evaluation-network flags (train.py)
    p.add_argument("--arch", type=str, default="FedNAS_V1", help="genotype name in models.darts or a genotype JSON file")
    p.add_argument("--auxiliary", action="store_true")
    p.add_argument("--auxiliary_weight", type=float, default=0.4)
    p.add_argument("--n_train", type=int, default=2048, help="synthetic-data size when --data is not given")
    return p


class DeviceDataset:
    """Whole image set resident on the device; ``batches`` yields augmented index-gathered minibatches."""

    def __init__(self, X: torch.Tensor, y: torch.Tensor, device, augment: bool, cutout: int = 0):
        self.X, self.y = X.to(device), y.to(device)
        self.augment, self.cutout = augment, cutout

    def __len__(self):
        return self.X.shape[0]

    def _aug(self, x: torch.Tensor, g: torch.Generator) -> torch.Tensor:
        B, C, H, W = x.shape
        pad = F.pad(x, (4, 4, 4, 4))   # RandomCrop(32, padding=4) as one batched gather
        ox = torch.randint(0, 9, (B,), device=x.device, generator=g)
        oy = torch.randint(0, 9, (B,), device=x.device, generator=g)
        ar = torch.arange(H, device=x.device)
        rows, cols = oy[:, None] + ar[None, :], ox[:, None] + torch.arange(W, device=x.device)[None, :]
        bi = torch.arange(B, device=x.device)[:, None, None, None]
        ci = torch.arange(C, device=x.device)[None, :, None, None]
        x = pad[bi, ci, rows[:, None, :, None], cols[:, None, None, :]]
        flip = torch.rand(B, device=x.device, generator=g) < 0.5
        x = torch.where(flip[:, None, None, None], x.flip(3), x)
        if self.cutout > 0:  # ``darts/utils.py:40-59``
            cy = torch.randint(0, H, (B,), device=x.device, generator=g)
            cx = torch.randint(0, W, (B,), device=x.device, generator=g)
            half = self.cutout // 2
            m = ((ar[None, :, None] >= (cy - half)[:, None, None]) & (ar[None, :, None] < (cy + half)[:, None, None]) &
                 (ar[None, None, :] >= (cx - half)[:, None, None]) & (ar[None, None, :] < (cx + half)[:, None, None]))
            x = x * (~m)[:, None].to(x.dtype)
        return x

    def batches(self, idx: torch.Tensor, batch_size: int, g: torch.Generator, shuffle: bool = True):
        if shuffle:
            idx = idx[torch.randperm(idx.numel(), device=idx.device, generator=g)]
        for i in range(0, idx.numel(), batch_size):
            b = idx[i:i + batch_size]
            x = self.X[b]
            yield (self._aug(x, g) if self.augment else x), self.y[b]


def _load(args, device):
    if args.data and os.path.isdir(args.data):
        real = benchmarks._read_cifar(args.data, args.dataset == "cifar100")
    else:
        real = None
    classes = 100 if args.dataset == "cifar100" else 10
    if real is None:
        rng = np.random.RandomState(args.seed)
        Xtr, ytr = benchmarks._synthetic_images(args.n_train, (3, 32, 32), classes, rng)
        Xte, yte = benchmarks._synthetic_images(max(args.n_train // 4, 64), (3, 32, 32), classes, rng)
    else:
        Xtr, ytr, Xte, yte = real
    return Xtr.float(), ytr.long(), Xte.float(), yte.long(), classes


def _device(args):
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", str(args.gpu).split(",")[0])))
    return torch.device("cpu")


def _resolve_genotype(name: str) -> darts.Genotype:
    if os.path.isfile(name):
        d = json.load(open(name))
        return darts.Genotype(normal=[tuple(e) for e in d["normal"]], normal_concat=d["normal_concat"],
                              reduce=[tuple(e) for e in d["reduce"]], reduce_concat=d["reduce_concat"])
    return getattr(darts, name)


def search(args, sink: MetricsSink | None = None):
    """``train_search.py:59-213``: alternate α steps (validation half) and weight steps (training half)."""
    dev = _device(args)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    Xtr, ytr, _, _, classes = _load(args, dev)
    ds = DeviceDataset(Xtr, ytr, dev, augment=True)
    n = len(ds)
    split = int(np.floor(args.train_portion * n))
    perm = torch.arange(n, device=dev)
    train_idx, valid_idx = perm[:split], perm[split:]
    criterion = nn.CrossEntropyLoss().to(dev)
    cls = darts.Network_GumbelSoftmax if args.arch_search_method.upper() == "GDAS" else darts.Network
    model = cls(args.init_channels, classes, args.layers, criterion).to(dev)
    logging.info("param size = %.3f MB", darts.count_parameters_in_MB(model))
    w_opt = torch.optim.SGD(model.weight_parameters(), args.learning_rate, momentum=args.momentum, weight_decay=args.weight_decay)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(w_opt, float(args.epochs), eta_min=args.learning_rate_min)
    architect = darts.Architect(model, criterion, args, dev)
    g = torch.Generator(device=dev).manual_seed(args.seed)
    best, history, patience = 0.0, [], 0
    for epoch in range(args.epochs):
        if isinstance(model, darts.Network_GumbelSoftmax):
            model.set_tau(args.tau_max - (args.tau_max - args.tau_min) * epoch / max(args.epochs - 1, 1))
        model.train()
        top1, objs = AvgrageMeter(), AvgrageMeter()
        valid_iter = ds.batches(valid_idx, args.batch_size, g)
        t0 = time.time()
        for step, (x, y) in enumerate(ds.batches(train_idx, args.batch_size, g)):
            try:
                xv, yv = next(valid_iter)
            except StopIteration:
                valid_iter = ds.batches(valid_idx, args.batch_size, g)
                xv, yv = next(valid_iter)
            if args.optimization == "DARTS_V2":
                architect.step_v2(x, y, xv, yv, args.lambda_train_regularizer, args.lambda_valid_regularizer)
            else:
                architect.step(xv, yv)
            for _ in range(args.w_update_times):
                w_opt.zero_grad()
                logits = model(x)
                loss = criterion(logits, y)
                loss.backward()
                nn.utils.clip_grad_norm_(model.weight_parameters(), args.grad_clip)
                w_opt.step()
            objs.update(loss.item(), x.size(0))
            top1.update(accuracy(logits, y)[0].item(), x.size(0))
            if step % args.report_freq == 0:
                logging.info("search %03d %03d loss %.4f top1 %.2f", epoch, step, objs.avg, top1.avg)
        sched.step()
        vacc, vloss = infer(model, ds, valid_idx, criterion, args.batch_size, g)
        geno = model.genotype()
        rec = {"epoch": epoch, "train_acc": top1.avg, "train_loss": objs.avg, "valid_acc": vacc, "valid_loss": vloss,
               "searching_cnn_count(%s)" % "skip_connect": sum(1 for op, _ in geno.normal if op == "skip_connect"),
               "epoch_s": time.time() - t0, "genotype": _geno_dict(geno)}
        history.append(rec)
        if sink is not None:
            sink.log({k: v for k, v in rec.items() if k != "genotype"})
        logging.info("genotype = %s", geno)
        if vacc > best:
            best, patience = vacc, 0
        else:
            patience += 1
        if args.early_stopping and patience >= args.early_stopping:
            break
    os.makedirs(args.model_path, exist_ok=True)
    torch.save(model.state_dict(), os.path.join(args.model_path, "search_weights.pt"))
    with open(os.path.join(args.model_path, "genotype.json"), "w") as fh:
        json.dump(_geno_dict(model.genotype()), fh)
    with open(os.path.join(args.model_path, "normal.dot"), "w") as fh:
        fh.write(darts.genotype_to_dot(model.genotype(), "normal"))
    return {"history": history, "genotype": model.genotype(), "best_valid_acc": best}


def _geno_dict(g: darts.Genotype) -> dict:
    return {"normal": [list(e) for e in g.normal], "normal_concat": list(g.normal_concat),
            "reduce": [list(e) for e in g.reduce], "reduce_concat": list(g.reduce_concat)}


@torch.no_grad()
def infer(model, ds: DeviceDataset, idx, criterion, batch_size, g):
    model.eval()
    top1, objs = AvgrageMeter(), AvgrageMeter()
    aug = ds.augment
    ds.augment = False
    try:
        for x, y in ds.batches(idx, batch_size, g, shuffle=False):
            out = model(x)
            logits = out[0] if isinstance(out, tuple) else out
            objs.update(criterion(logits, y).item(), x.size(0))
            top1.update(accuracy(logits, y)[0].item(), x.size(0))
    finally:
        ds.augment = aug
    return top1.avg, objs.avg


def train(args, sink: MetricsSink | None = None):
    """``train.py:57-161``: train the discrete network of ``--arch`` (aux head, linear drop-path schedule, cutout)."""
    dev = _device(args)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    Xtr, ytr, Xte, yte, classes = _load(args, dev)
    tr = DeviceDataset(Xtr, ytr, dev, augment=True, cutout=args.cutout_length if args.cutout else 0)
    te = DeviceDataset(Xte, yte, dev, augment=False)
    genotype = _resolve_genotype(args.arch)
    net = darts.NetworkCIFAR(args.init_channels, classes, args.layers, args.auxiliary, genotype).to(dev)
    logging.info("param size = %.3f MB", darts.count_parameters_in_MB(net))
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    model = net
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl" if dev.type == "cuda" else "gloo")
        model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index] if dev.type == "cuda" else None,
                                                    find_unused_parameters=True)
    criterion = nn.CrossEntropyLoss().to(dev)
    opt = torch.optim.SGD(model.parameters(), args.learning_rate, momentum=args.momentum, weight_decay=args.weight_decay)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, float(args.epochs))
    g = torch.Generator(device=dev).manual_seed(args.seed)
    all_tr, all_te = torch.arange(rank, len(tr), world, device=dev), torch.arange(len(te), device=dev)
    history, best = [], 0.0
    for epoch in range(args.epochs):
        net.drop_path_prob = args.drop_path_prob * epoch / max(args.epochs, 1)
        model.train()
        top1, objs = AvgrageMeter(), AvgrageMeter()
        for step, (x, y) in enumerate(tr.batches(all_tr, args.batch_size, g)):
            opt.zero_grad()
            logits, logits_aux = model(x)
            loss = criterion(logits, y)
            if args.auxiliary and logits_aux is not None:
                loss = loss + args.auxiliary_weight * criterion(logits_aux, y)
            loss.backward()
            nn.utils.clip_grad_norm_(model.parameters(), args.grad_clip)
            opt.step()
            objs.update(loss.item(), x.size(0))
            top1.update(accuracy(logits, y)[0].item(), x.size(0))
            if step % args.report_freq == 0:
                logging.info("train %03d %03d loss %.4f top1 %.2f", epoch, step, objs.avg, top1.avg)
        sched.step()
        vacc, vloss = infer(model, te, all_te, criterion, args.batch_size, g)
        rec = {"epoch": epoch, "train_acc": top1.avg, "train_loss": objs.avg, "valid_acc": vacc, "valid_loss": vloss}
        history.append(rec)
        if sink is not None:
            sink.log(rec)
        if vacc >= best:
            best = vacc
            if rank == 0:
                os.makedirs(args.model_path, exist_ok=True)
                torch.save(net.state_dict(), os.path.join(args.model_path, "weights.pt"))
    return {"history": history, "best_valid_acc": best}


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    args = add_args(argparse.ArgumentParser("feddrift_b200 DARTS")).parse_args(argv)
    sink = MetricsSink()
    out = search(args, sink) if args.mode == "search" else train(args, sink)
    print(json.dumps({"mode": args.mode, "best_valid_acc": out["best_valid_acc"], "epochs": len(out["history"])}))
    return out


if __name__ == "__main__":
    main()
