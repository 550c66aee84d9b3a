"""The non-drift FedML algorithm packages on CPU."""
import copy
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from feddrift_b200.data.drift import generate_drift_data, load_partition_data
from feddrift_b200.drift.fedavg_ens import FedML_init
from feddrift_b200.models import create_model
from feddrift_b200.utils.metrics import MetricsSink, set_sink


def _dataset(clients=4, n=60, batch=20):
    d = generate_drift_data("sine", 1, clients, n, 0.0, 1, np.zeros((2, clients), dtype=np.int64))
    tup = load_partition_data(d, batch, 0, "win-1", rng=np.random.RandomState(0))
    return list(tup[1:]), d


def _args(**kw):
    base = dict(client_num_in_total=4, client_num_per_round=4, comm_round=4, epochs=1, lr=0.05, wd=0.0, client_optimizer="sgd",
                frequency_of_the_test=1, dataset="sine", ci=0, report_client=0, is_mobile=0, batch_size=20, dummy_arg=0)
    base.update(kw)
    return SimpleNamespace(**base)


def test_standalone_fedavg_learns_and_matches_manual_average():
    from feddrift_b200.fl.standalone import FedAvgTrainer
    sink = set_sink(MetricsSink())
    ds, _ = _dataset()
    tr = FedAvgTrainer(ds, create_model("fnn", 2, 2), "cpu", _args(comm_round=20, client_optimizer="adam", epochs=3))
    tr.train()
    accs = sink.series("Train/Acc")
    assert accs[-1] > accs[0] and accs[-1] > 0.7
    rows, ns = torch.randn(3, tr.bank.P), [10.0, 30.0, 60.0]
    assert torch.allclose(tr.aggregate_rows(rows, ns), 0.1 * rows[0] + 0.3 * rows[1] + 0.6 * rows[2], atol=1e-6)


@pytest.mark.parametrize("opt", ["sgd", "adam", "adagrad", "yogi", "rmsprop"])
def test_fedopt_server_optimizers(opt):
    from feddrift_b200.fl.standalone import FedOptTrainer, OptRepo
    set_sink(MetricsSink())
    ds, _ = _dataset()
    a = _args(comm_round=3, server_optimizer=opt, server_lr=1.0 if opt == "sgd" else 0.05, server_momentum=0.0)
    tr = FedOptTrainer(ds, create_model("fnn", 2, 2), "cpu", a)
    before = tr.bank.theta[0].clone()
    tr.train()
    assert torch.isfinite(tr.bank.theta[0]).all() and not torch.equal(before, tr.bank.theta[0])
    assert OptRepo.name2cls("adam") is torch.optim.Adam and "lr" in OptRepo.supported_parameters("SGD")


def test_fedopt_sgd_lr1_equals_fedavg():
    from feddrift_b200.fl.standalone import FedAvgTrainer, FedOptTrainer
    set_sink(MetricsSink())
    ds, _ = _dataset()
    m = create_model("fnn", 2, 2)
    a1 = FedAvgTrainer(ds, copy.deepcopy(m), "cpu", _args(comm_round=2))
    a2 = FedOptTrainer(ds, copy.deepcopy(m), "cpu", _args(comm_round=2, server_optimizer="sgd", server_lr=1.0))
    a1.train(); a2.train()
    assert torch.allclose(a1.bank.theta[0], a2.bank.theta[0], atol=1e-6)


def test_hierarchical_fl_runs():
    from feddrift_b200.fl.standalone import HierarchicalTrainer
    sink = set_sink(MetricsSink())
    ds, _ = _dataset(clients=6)
    a = _args(client_num_in_total=6, client_num_per_round=6, group_num=2, group_method="random", global_comm_round=2,
              group_comm_round=2)
    np.random.seed(0)
    HierarchicalTrainer(ds, create_model("fnn", 2, 2), "cpu", a).train()
    assert len(sink.series("Test/Acc")) >= 2


@pytest.mark.parametrize("robust", [False, True])
def test_distributed_fedavg_inproc(robust):
    from feddrift_b200.fl.fedavg import FedML_FedAvg_distributed
    sink = set_sink(MetricsSink())
    ds, _ = _dataset()
    a = _args(comm_round=3, epochs=2, defense_type="weak_dp", norm_bound=0.5, stddev=0.01)
    comm, pid, size = FedML_init("INPROC", 5)
    srv = FedML_FedAvg_distributed(pid, size, "cpu", comm, create_model("fnn", 2, 2), ds[0], ds[2], ds[3], ds[4], ds[5],
                                   ds[6], a, robust=robust)
    assert srv.round_idx == 3 and len(sink.series("Test/Acc")) == 3
    assert torch.isfinite(srv.aggregator.bank.theta).all()


def test_distributed_fedavg_survives_a_lost_upload_with_the_round_watchdog():
    from feddrift_b200.fl.fedavg import FedML_FedAvg_distributed
    sink = set_sink(MetricsSink())
    ds, _ = _dataset()
    a = _args(comm_round=3, epochs=2, round_timeout_s=5.0, min_workers_per_round=2, fault_drop={1: [2]})
    comm, pid, size = FedML_init("INPROC", 5)
    srv = FedML_FedAvg_distributed(pid, size, "cpu", comm, create_model("fnn", 2, 2), ds[0], ds[2], ds[3], ds[4], ds[5], ds[6], a)
    assert srv.round_idx == 3 and srv.watchdog.timeouts == 1 and len(sink.series("Test/Acc")) == 3


def test_weak_dp_noise_is_standard_normal_and_deterministic():
    from feddrift_b200.ops import reference as ref
    z = ref.gauss_hash(7, 4, 50000)
    assert torch.equal(z, ref.gauss_hash(7, 4, 50000)) and not torch.equal(z, ref.gauss_hash(8, 4, 50000))
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01
    assert abs(torch.corrcoef(z[:2])[0, 1].item()) < 0.02        # rows are independent streams
    rows, g = torch.zeros(2, 1000), torch.zeros(1000)
    ref.robust_clip_(rows, g, 1.0, None, 0.5, 3)
    assert abs(rows.std().item() - 0.5) < 0.05


def test_robust_clipping_bounds_the_update():
    from feddrift_b200.fl.fedavg import FedAvgRobustAggregator
    ds, _ = _dataset()
    a = _args(defense_type="norm_diff_clipping", norm_bound=0.1, stddev=0.0)
    agg = FedAvgRobustAggregator(ds[2], ds[3], ds[0], ds[5], ds[6], ds[4], 2, "cpu", create_model("fnn", 2, 2), a)
    g = agg.bank.theta[0].clone()
    for i in range(2):
        sd = {k: v + 5.0 for k, v in agg.bank.state_dict(0).items()}
        agg.add_local_trained_result(i, sd, 10)
    agg.aggregate()
    assert (agg.bank.theta[0] - g).norm().item() <= 0.1 + 1e-5


def test_decentralized_dsgd_and_pushsum_reduce_regret():
    from feddrift_b200.fl.decentralized import DecentralizedSimulator
    set_sink(MetricsSink())
    rng = np.random.RandomState(0)
    n, T, d = 8, 300, 6
    w = rng.randn(d)
    data = [[{"x": (x := rng.randn(d).astype(np.float32)), "y": float(x @ w > 0)} for _ in range(T)] for _ in range(n)]
    for mode, sym in (("DOL", True), ("PUSHSUM", False), ("LOCAL", True)):
        a = SimpleNamespace(iteration_number=T, learning_rate=0.3, batch_size=1, weight_decay=0.0, epoch=1, mode=mode,
                            topology_neighbors_num_undirected=4, topology_neighbors_num_directed=2, latency=0,
                            b_symmetric=sym, time_varying=(mode == "PUSHSUM"), log_every=50, seed=1)
        r = DecentralizedSimulator(n, data, d, a).run()
        assert r[-1] < r[0] and r[-1] < 0.6, (mode, r)


def test_framework_templates():
    from feddrift_b200.fl.frameworks import FedML_Base_distributed, FedML_Decentralized_Demo_distributed
    a = SimpleNamespace(comm_round=3)
    comm, _, size = FedML_init("INPROC", 5)
    srv = FedML_Base_distributed(0, size, comm, a)
    assert srv.history == [0 + 1 + 2 + 3] * 3
    comm, _, size = FedML_init("INPROC", 6)
    mgrs = FedML_Decentralized_Demo_distributed(0, size, comm, a)
    assert all(m.completed == [0, 1, 2] for m in mgrs)


def _img_loaders(n_clients=2, n=24, classes=4, bs=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_clients):
        y = torch.randint(0, classes, (n,), generator=g)
        x = torch.randn(n, 3, 16, 16, generator=g) + y[:, None, None, None].float()
        tr = [(x[i:i + bs], y[i:i + bs]) for i in range(0, n, bs)]
        out.append((tr, tr[:1]))
    return out


def test_splitnn_round_robin():
    from feddrift_b200.fl.split import SplitNN_distributed, split_model
    set_sink(MetricsSink())
    net = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 16 * 16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))
    bottom, top = split_model(net, 2)
    loaders = _img_loaders()
    res = SplitNN_distributed([copy.deepcopy(bottom) for _ in loaders], top, loaders, "cpu", epochs=3, lr=0.05)
    assert len(res) == 6 and res[-1]["acc"] >= res[0]["acc"] and all(np.isfinite(r["loss"]) for r in res)


def test_fedgkt_two_stage_training(tmp_path):
    import json
    import os
    from feddrift_b200.fl.split import FedML_FedGKT_distributed, GKTServerTrainer
    from feddrift_b200.models.resnet import resnet8_56, ResNet, Bottleneck
    sink = set_sink(MetricsSink())
    cdir = str(tmp_path / "checkpoint")
    a = SimpleNamespace(comm_round=2, epochs_client=1, epochs_server=1, lr=0.01, wd=1e-4, optimizer="SGD", temperature=3.0,
                        alpha=1.0, whether_training_on_client=1, whether_distill_on_the_server=1, checkpoint_dir=cdir)
    loaders = _img_loaders()
    server_model = ResNet(Bottleneck, [1, 1, 1], 4, stem="features")
    srv, hist = FedML_FedGKT_distributed([resnet8_56(4) for _ in loaders], server_model, loaders, "cpu", a)
    assert len(hist) == 2 and "best" in srv.checkpoints and len(sink.series("Test/AccTop1")) == 2
    assert set(srv.get_global_logits(0).keys()) == {0, 1, 2}
    # file checkpoints like the reference (GKTServerTrainer.py:213-231): last.pth, best.pth, test_best_metrics.json
    assert sorted(os.listdir(cdir)) == ["best.pth", "last.pth", "test_best_metrics.json"]
    last = torch.load(os.path.join(cdir, "last.pth"), weights_only=True)
    assert last["epoch"] == 2 and set(last) >= {"state_dict", "optim_dict", "test_accTop1", "test_accTop5"}
    assert json.load(open(os.path.join(cdir, "test_best_metrics.json")))["epoch"] in (1, 2)
    fresh = GKTServerTrainer(len(loaders), "cpu", ResNet(Bottleneck, [1, 1, 1], 4, stem="features"), a)
    assert fresh.resume(os.path.join(cdir, "last.pth")) == 2
    for k, v in srv.model_global.state_dict().items():
        assert torch.equal(v, fresh.model_global.state_dict()[k])


def test_vertical_fl_distributed_and_standalone():
    from feddrift_b200.fl.split import (FedML_VFL_distributed, VFLGuestModel, VFLGuestTrainer, VFLHostModel, VFLHostTrainer,
                                        VerticalMultiplePartyLogisticRegressionFederatedLearning)
    from feddrift_b200.models.vfl import LocalModel, VFLClassifier, VFLFeatureExtractor
    set_sink(MetricsSink())
    rng = np.random.RandomState(0)
    n = 256
    Xa, Xb = rng.randn(n, 5).astype(np.float32), rng.randn(n, 4).astype(np.float32)
    y = ((Xa[:, 0] + Xb[:, 1]) > 0).astype(np.float32)
    a = SimpleNamespace(lr=0.05, batch_size=64, frequency_of_the_test=1)
    guest = VFLGuestTrainer(2, "cpu", Xa, y, Xa, y, VFLFeatureExtractor(5, 8), VFLClassifier(8, 1), a)
    host = VFLHostTrainer(1, "cpu", Xb, Xb, VFLFeatureExtractor(4, 8), VFLClassifier(8, 1, bias=False), a)
    hist = FedML_VFL_distributed(guest, [host], comm_round=6)
    assert hist[-1]["test_acc"] > 0.8 and hist[-1]["test_auc"] > 0.85
    fl = VerticalMultiplePartyLogisticRegressionFederatedLearning(VFLGuestModel(LocalModel(5, 6, 0.05), learning_rate=0.05))
    fl.add_party(id="B", party_model=VFLHostModel(LocalModel(4, 6, 0.05), learning_rate=0.05))
    losses = [fl.fit(Xa[i:i + 64], y[i:i + 64], {"B": Xb[i:i + 64]}, s) for s in range(8) for i in range(0, n, 64)]
    assert np.mean(losses[-4:]) < np.mean(losses[:4])
    pred = fl.predict(Xa, {"B": Xb})
    assert ((pred.flatten() > 0.5) == (y > 0.5)).mean() > 0.7


def test_fednas_search_round():
    from feddrift_b200.fl.fednas import FedML_FedNAS_distributed
    from feddrift_b200.models.darts import Genotype, Network, NetworkCIFAR, Network_GumbelSoftmax
    set_sink(MetricsSink())
    a = SimpleNamespace(comm_round=1, epochs=1, learning_rate=0.025, momentum=0.9, weight_decay=3e-4, grad_clip=5.0,
                        arch_learning_rate=3e-3, arch_weight_decay=1e-3, lambda_train_regularizer=1.0, lambda_valid_regularizer=1.0)
    loaders = _img_loaders(n=16)
    net = Network(4, 4, 3)
    assert net.alphas_normal.shape == (14, 8)
    a0 = net.alphas_normal.detach().clone()
    agg, hist = FedML_FedNAS_distributed(net, loaders, loaders[0][0], "cpu", a)
    g = hist[0][1]
    assert isinstance(g, Genotype) and len(g.normal) == 8 and not torch.equal(agg.model.alphas_normal.detach(), a0)
    ev = NetworkCIFAR(4, 4, 3, False, g)
    logits, aux = ev(torch.randn(2, 3, 16, 16))
    assert logits.shape == (2, 4) and aux is None
    gd = Network_GumbelSoftmax(4, 4, 3)
    assert gd(torch.randn(2, 3, 16, 16)).shape == (2, 4)


def test_turboaggregate_primitives_and_secure_average():
    from feddrift_b200.fl import turboaggregate as ta
    p, rng = 2 ** 31 - 1, np.random.RandomState(0)
    X = rng.randint(0, p, size=(4, 6))
    sh = ta.BGW_encoding(X, 7, 2, p, rng)
    idx = [1, 4, 6]
    assert (ta.BGW_decoding(sh[idx].reshape(3, -1), idx, p).reshape(4, 6) == X % p).all()
    assert ta.modular_inv(3, 7) == 5 and ta.divmod(6, 3, 7) == 2 and ta.PI([2, 3, 4], 5) == 4
    U = ta.gen_Lagrange_coeffs([1, 2], [3, 4, 5], p)
    assert U.shape == (2, 3) and (U.sum(1) % p == 1).all()          # Lagrange basis sums to one
    N, K, T = 8, 2, 1
    enc = ta.LCC_encoding(X, N, K, T, p, rng)
    assert enc.shape == (N, 2, 6)
    pts_a, pts_b = np.array([1, 2, 3]), np.array([10, 11, 12, 13])
    Y = rng.randint(0, p, size=(3, 5))
    enc2 = ta.LCC_encoding_with_points(Y, pts_a, pts_b, p)
    assert (ta.LCC_decoding_with_points(enc2[:3], pts_b[:3], pts_a, p) == Y).all()
    assert (ta.Gen_Additive_SS(5, 4, p, rng).sum(0) % p == 0).all()
    assert ta.my_key_agreement(12, ta.my_pk_gen(34, p, 5), p, 5) == ta.my_key_agreement(34, ta.my_pk_gen(12, p, 5), p, 5)
    agg = ta.TurboAggregator(6, 2)
    Ux, w = torch.randn(6, 40), torch.rand(6) + 0.1
    out = agg.aggregate(Ux, w, dropped=[0, 5])
    assert (out - (Ux * (w / w.sum())[:, None]).sum(0)).abs().max() < 1e-3


def test_darts_standalone_search_then_train(tmp_path):
    """``darts/train_search.py`` + ``darts/train.py`` parity CLI: search (MiLeNAS step) → genotype.json → train it."""
    from feddrift_b200.experiments import darts as cli
    common = ["--epochs", "1", "--layers", "3", "--init_channels", "4", "--n_train", "64", "--batch_size", "32",
              "--model_path", str(tmp_path)]
    out = cli.main(["search", "--optimization", "DARTS_V2"] + common)
    assert len(out["genotype"].normal) == 8 and (tmp_path / "genotype.json").exists()
    assert "digraph" in (tmp_path / "normal.dot").read_text()
    out = cli.main(["train", "--arch", str(tmp_path / "genotype.json"), "--auxiliary", "--cutout"] + common)
    assert len(out["history"]) == 1 and (tmp_path / "weights.pt").exists()


def test_darts_published_genotypes_and_imagenet_network():
    from feddrift_b200.models import darts
    for name in ("NASNet", "AmoebaNet", "DARTS_V1", "DARTS_V2", "FedNAS_V1"):
        g = getattr(darts, name)
        net = darts.NetworkCIFAR(8, 10, 3, False, g).eval()
        assert net(torch.randn(2, 3, 32, 32))[0].shape == (2, 10), name
    assert darts.DARTS is darts.DARTS_V2 and darts.FedNAS_V1.normal[2] == ("sep_conv_3x3", 2)
    inet = darts.NetworkImageNet(16, 20, 3, True, darts.DARTS_V2).train()
    logits, aux = inet(torch.randn(2, 3, 224, 224))
    assert logits.shape == (2, 20) and aux.shape == (2, 20)
