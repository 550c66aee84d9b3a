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
