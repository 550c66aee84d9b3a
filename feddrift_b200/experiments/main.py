"""One CLI for the non-drift experiment mains of ``fedml_experiments/{distributed,standalone}/*``:

  python -m feddrift_b200.experiments.main fedavg        --dataset cifar10 --model resnet56 ...
  python -m feddrift_b200.experiments.main fedavg_robust --defense_type weak_dp --norm_bound 5 --stddev 0.025 ...
  python -m feddrift_b200.experiments.main fedopt        --server_optimizer adam --server_lr 0.01 ...
  python -m feddrift_b200.experiments.main hierarchical  --group_num 2 --global_comm_round 5 --group_comm_round 2 ...
  python -m feddrift_b200.experiments.main decentralized --mode PUSHSUM --iteration_number 2000 ...
  python -m feddrift_b200.experiments.main fedgkt | fednas | split_nn | vfl | base | decentralized_demo

Every sub-command accepts the flags of the corresponding reference ``main_*.py`` (same names and defaults where they
exist); data come from ``data/benchmarks.py`` (real files under ``--data_dir`` if present, synthetic of the same shape
otherwise).  The continual-drift experiments (``fedavg_cont_ens`` / ``fedavg_cont_one``) live in
``experiments/fedavg_cont_ens.py``.
"""
from __future__ import annotations

import argparse
import copy
import json
import sys

import numpy as np
import torch


def common_args(p: argparse.ArgumentParser) -> argparse.ArgumentParser:
    a = p.add_argument
    a("--model", type=str, default="cnn"); a("--dataset", type=str, default="mnist"); a("--data_dir", type=str, default=None)
    a("--partition_method", type=str, default="hetero"); a("--partition_alpha", type=float, default=0.5)
    a("--client_num_in_total", type=int, default=8); a("--client_num_per_round", type=int, default=8)
    a("--batch_size", type=int, default=32); a("--client_optimizer", type=str, default="sgd")
    a("--lr", type=float, default=0.03); a("--wd", type=float, default=0.001); a("--epochs", type=int, default=1)
    a("--comm_round", type=int, default=5); a("--frequency_of_the_test", type=int, default=1)
    a("--is_mobile", type=int, default=0); a("--ci", type=int, default=0); a("--report_client", type=int, default=0)
    a("--gpu", type=int, default=0); a("--device", type=str, default=None); a("--dummy_arg", type=int, default=0)
    a("--backend", type=str, default="INPROC")
    # fedavg_robust
    a("--defense_type", type=str, default="weak_dp"); a("--norm_bound", type=float, default=5.0); a("--stddev", type=float, default=0.025)
    # fedopt
    a("--server_optimizer", type=str, default="sgd"); a("--server_lr", type=float, default=1.0); a("--server_momentum", type=float, default=0.0)
    # hierarchical
    a("--group_method", type=str, default="random"); a("--group_num", type=int, default=2)
    a("--global_comm_round", type=int, default=2); a("--group_comm_round", type=int, default=2)
    # decentralized online learning
    a("--mode", type=str, default="DOL"); a("--iteration_number", type=int, default=200); a("--learning_rate", type=float, default=0.1)
    a("--weight_decay", type=float, default=0.0); a("--epoch", type=int, default=1); a("--b_symmetric", type=int, default=1)
    a("--topology_neighbors_num_undirected", type=int, default=4); a("--topology_neighbors_num_directed", type=int, default=2)
    a("--time_varying", type=int, default=0); a("--latency", type=float, default=0.0); a("--beta", type=float, default=0.5)
    # gkt / nas
    a("--epochs_client", type=int, default=1); a("--epochs_server", type=int, default=1); a("--temperature", type=float, default=3.0)
    a("--alpha", type=float, default=1.0); a("--optimizer", type=str, default="SGD")
    a("--whether_training_on_client", type=int, default=1); a("--whether_distill_on_the_server", type=int, default=1)
    a("--init_channels", type=int, default=8); a("--layers", type=int, default=3); a("--arch_learning_rate", type=float, default=3e-4)
    a("--arch_weight_decay", type=float, default=1e-3); a("--momentum", type=float, default=0.9); a("--weight_decay_nas", type=float, default=3e-4)
    a("--grad_clip", type=float, default=5.0); a("--lambda_train_regularizer", type=float, default=1.0)
    a("--lambda_valid_regularizer", type=float, default=1.0)
    return p


def _device(args):
    return torch.device(args.device or (f"cuda:{args.gpu}" if torch.cuda.is_available() else "cpu"))


def _model(args, class_num, ds):
    from ..models import create_model
    x0 = ds[5][next(iter(ds[5]))][0][0]
    feat = int(np.prod(x0.shape[1:]))
    return create_model(args.model, class_num, feat)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return None
    cmd = argv.pop(0)
    args = common_args(argparse.ArgumentParser(prog=f"feddrift_b200.experiments.main {cmd}")).parse_args(argv)
    np.random.seed(0)
    torch.manual_seed(10)
    from ..utils.metrics import MetricsSink, set_sink
    sink = set_sink(MetricsSink())
    dev = _device(args)
    from ..data import benchmarks as B
    out = {}
    if cmd in ("fedavg", "fedavg_robust"):
        from ..drift.fedavg_ens import FedML_init
        from ..fl.fedavg import FedML_FedAvg_distributed
        ds = B.load_data(args, args.dataset)
        comm, pid, size = FedML_init(args.backend, args.client_num_per_round + 1)
        FedML_FedAvg_distributed(pid, size, dev, comm, _model(args, ds[7], ds), ds[0], ds[2], ds[3], ds[4], ds[5], ds[6], args,
                                 robust=(cmd == "fedavg_robust"))
    elif cmd in ("fedavg_standalone", "fedopt", "hierarchical"):
        from ..fl.standalone import FedAvgTrainer, FedOptTrainer, HierarchicalTrainer
        ds = B.load_data(args, args.dataset)
        cls = {"fedavg_standalone": FedAvgTrainer, "fedopt": FedOptTrainer, "hierarchical": HierarchicalTrainer}[cmd]
        cls(ds, _model(args, ds[7], ds), dev, args).train()
    elif cmd == "decentralized":
        from ..fl.decentralized import DecentralizedSimulator
        data = B.load_streaming_susy_or_ro(args.client_num_in_total, args.iteration_number, args.dataset if args.dataset in ("SUSY", "RO") else "SUSY",
                                           args.beta)
        out["regret"] = DecentralizedSimulator(args.client_num_in_total, data, len(data[0][0]["x"]), args, dev).run()[-1]
    elif cmd in ("base", "decentralized_demo"):
        from ..drift.fedavg_ens import FedML_init
        from ..fl.frameworks import FedML_Base_distributed, FedML_Decentralized_Demo_distributed
        comm, _, size = FedML_init("INPROC", args.client_num_per_round + (1 if cmd == "base" else 0))
        r = (FedML_Base_distributed if cmd == "base" else FedML_Decentralized_Demo_distributed)(0, size, comm, args)
        out["result"] = r.history if cmd == "base" else [m.completed for m in r]
    elif cmd in ("fedgkt", "split_nn", "fednas"):
        a2 = copy.copy(args)
        a2.dataset = args.dataset if args.dataset in ("cifar10", "cifar100", "cinic10") else "cifar10"
        ds = B.load_data(a2, a2.dataset)
        loaders = [(ds[5][c], ds[6][c][:2]) for c in range(args.client_num_in_total)]
        if cmd == "fedgkt":
            from ..fl.split import FedML_FedGKT_distributed
            from ..models.resnet import resnet8_56, resnet56_server
            _, hist = FedML_FedGKT_distributed([resnet8_56(ds[7]) for _ in loaders], resnet56_server(ds[7]), loaders, dev, args)
            out["test_acc"] = hist[-1][1]["test_accTop1"]
        elif cmd == "split_nn":
            from ..fl.split import SplitNN_distributed, split_model
            from ..models.mobilenet import mobilenet
            bottom, top = split_model(mobilenet(1, ds[7]), 1)
            res = SplitNN_distributed([copy.deepcopy(bottom) for _ in loaders], top, loaders, dev, epochs=args.epochs, lr=args.lr)
            out["val_acc"] = res[-1]["acc"]
        else:
            from ..fl.fednas import FedML_FedNAS_distributed
            from ..models.darts import Network
            args.learning_rate, args.weight_decay = args.lr, args.weight_decay_nas
            agg, hist = FedML_FedNAS_distributed(Network(args.init_channels, ds[7], args.layers), loaders, ds[3][:2], dev, args)
            out["genotype"] = str(hist[-1][1])
    elif cmd == "vfl":
        from ..fl.split import FedML_VFL_distributed, VFLGuestTrainer, VFLHostTrainer
        from ..models.vfl import VFLClassifier, VFLFeatureExtractor
        Xtr, ytr, Xte, yte = B.load_vertical_parties(args.dataset, 2000, max(2, args.client_num_in_total))
        guest = VFLGuestTrainer(len(Xtr), dev, Xtr[0], ytr, Xte[0], yte, VFLFeatureExtractor(Xtr[0].shape[1], 10), VFLClassifier(10, 1), args)
        hosts = [VFLHostTrainer(i, dev, Xtr[i], Xte[i], VFLFeatureExtractor(Xtr[i].shape[1], 10), VFLClassifier(10, 1, bias=False), args)
                 for i in range(1, len(Xtr))]
        out["metrics"] = FedML_VFL_distributed(guest, hosts, args.comm_round)[-1]
    else:
        raise SystemExit(f"unknown experiment {cmd!r}")
    out.update({k: sink.last(k) for k in ("Train/Acc", "Test/Acc") if sink.last(k) is not None})
    print(json.dumps({"experiment": cmd, **out}, default=str))
    return out


if __name__ == "__main__":
    main()
