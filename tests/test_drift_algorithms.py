"""Drift algorithms on the CPU oracle path: state machines, clustering decisions, every ``--concept_drift_algo``."""
import numpy as np
import pytest
import scipy.cluster.hierarchy as sch
import torch
from scipy.spatial.distance import squareform

from feddrift_b200.drift.hclust import complete_linkage_bipartition, linkage_fcluster
from feddrift_b200.drift.softcluster import SoftClusterState, parse_algo_arg
from feddrift_b200.drift.states import AdaState, KueState
from feddrift_b200.ops import reference as ref
from feddrift_b200.sim import DriftSim, make_args
from feddrift_b200.utils.metrics import MetricsSink


def test_linkage_matches_scipy_and_bipartition_matches_sklearn():
    from sklearn.cluster import AgglomerativeClustering
    rng = np.random.RandomState(1)
    for _ in range(60):
        L = rng.randint(2, 9)
        A = rng.rand(L, L)
        D = np.maximum(A, A.T)
        np.fill_diagonal(D, 0)
        for method in ("complete", "average"):
            thr = rng.rand() * 0.8
            T = sch.fcluster(sch.linkage(squareform(D), method=method), t=thr, criterion="distance")
            mine = linkage_fcluster(D, method, thr)
            assert all((T[i] == T[j]) == (mine[i] == mine[j]) for i in range(L) for j in range(L))
        S = (A + A.T) / 2
        lab = AgglomerativeClustering(metric="precomputed", linkage="complete").fit(-S).labels_
        g1, g2 = complete_linkage_bipartition(S)
        assert set(np.where(lab == 0)[0]) == set(g1) and set(np.where(lab == 1)[0]) == set(g2)


def test_arg_grammar():
    c = parse_algo_arg("H_A_C_1_10_0", "sea")
    assert (c["h_distance"], c["h_cluster"], c["h_w"]) == ("A", "C", 1) and abs(c["h_delta"] - 0.10) < 1e-9 \
        and abs(c["h_deltap"] - 0.10) < 1e-9
    assert abs(parse_algo_arg("H_B_D_2_0_0", "sea")["h_delta"] - 0.04) < 1e-9   # δ = 0 → dataset default
    assert abs(parse_algo_arg("mmacc_06", "sine")["mmacc_delta"] - 0.06) < 1e-9
    assert parse_algo_arg("softmax_3")["softmax_alpha"] == 3
    c = parse_algo_arg("cfl_0.1_win-1")
    assert c["cfl_gamma"] == 0.1 and c["cfl_retrain"] == "win-1"


def test_softcluster_primitives():
    sink = MetricsSink()
    st = SoftClusterState(4, 3, "hard", sink=sink)
    acc = np.array([[.9, .1, .5, .5], [.1, .9, .5, .4], [.2, .2, .5, .6]])
    st.cluster(acc, 0, 0)
    assert st.test_model_indices(0).tolist() == [0, 1, 0, 2]          # ties → first max (np.argmax)
    st2 = SoftClusterState(4, 3, "softmax_1", softmax_alpha=1, sink=sink)
    st2.cluster(acc, 0, 0)
    assert np.allclose(st2.W[0].sum(0), 1) and st2.W[0][0, 0] > st2.W[0][1, 0]
    # LRU allocation: fresh slots first, then least-recently-used, never a slot live at the current step
    st3 = SoftClusterState(2, 2, "H_A_C_1_10_0", sink=sink)
    st3.cluster_init()
    assert st3.find_unused_model_lru(1) == 1
    st3._new_step(1)[0, :] = 1
    st3.W[1][1, 0] = 1
    assert st3.find_unused_model_lru(1) == -1
    st3._new_step(2)[0, :] = 1
    assert st3.find_unused_model_lru(2) == 1 and st3.W[:, 1].sum() == 0


def test_ada_state_matches_closed_form():
    s = AdaState(init_lr=0.01)
    th = torch.tensor([1.0, 2.0, 3.0])
    s.update(th, 0)
    assert abs(s.eta - 0.01) < 1e-12 and torch.allclose(s.mu, 0.5 * th)
    s.update(th * 1.1, 1)
    assert 0 < s.eta <= 0.01


def test_reference_adam_equals_torch_optim():
    p0, g = torch.randn(50), [torch.randn(50) for _ in range(5)]
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([q], lr=0.01, weight_decay=1e-3, amsgrad=True)
    p, m, v, vm, step = p0.clone(), torch.zeros(50), torch.zeros(50), torch.zeros(50), 0
    for gi in g:
        q.grad = gi.clone()
        opt.step()
        step = ref.adam_amsgrad_update(p, gi, m, v, vm, step, 0.01, 1e-3)
    assert torch.allclose(p, q.detach(), rtol=1e-5, atol=1e-7)


def test_feddrift_recovers_ground_truth_concepts():
    """README-style run (sine, change-point matrix A, FedDrift H_A_C_1_δ_δ'): the hierarchical soft-clustering must
    (i) detect each client's drift, (ii) isolate it on a private model for W = 1 step, (iii) merge private models
    of the same concept — so that the cluster assignment equals the ground-truth concept at EVERY time step
    (clustering-trajectory golden test)."""
    sink = MetricsSink()
    sim = DriftSim(make_args(dataset="sine", concept_drift_algo_arg="H_A_C_1_0_0", comm_round=30, lr=0.05,
                             total_train_iteration=6, sample_num=100), device="cpu", sink=sink)
    sim.run()
    st, cps = sim.algo.state, sim.data_host.change_points
    assert sink.run.summary["num_models"] == 2
    for t in range(6):
        assert st.test_model_indices(t).tolist() == cps[t].tolist(), t
    assert sim.history[-1]["train_acc"] > 0.95
    # Plurality/CL-c logs exist for every client at every time step (round 0 of each step)
    assert len(sink.series("Plurality/CL-3")) >= 6


ALGOS = [
    ("softcluster", "H_A_C_1_10_0"), ("softcluster", "H_B_D_2_0_0"), ("softcluster", "H_A_E_1_10_0"),
    ("softcluster", "mmacc_06"), ("softcluster", "hard"), ("softclusterwin-1", "hard-r"), ("softcluster", "softmax_2"),
    ("softcluster", "geni"), ("softcluster", "cfl_0.1_win-1"), ("softclusterreset", "softmax_0"),
    ("win-1", ""), ("all", ""), ("weight-linear", ""), ("lin", ""), ("exp", ""), ("ada", "win-1_round"),
    ("ada", "all_iter"), ("aue", ""), ("auepc", ""), ("kue", ""), ("driftsurf", "0"), ("mmacc", ""),
    ("mmgeni", ""), ("mmgeniex", ""), ("clusterfl", "win-1"),
]


@pytest.mark.parametrize("algo,arg", ALGOS)
def test_every_algorithm_runs_three_time_steps(algo, arg):
    kw = dict(concept_drift_algo=algo, concept_drift_algo_arg=arg, comm_round=6, total_train_iteration=3,
              sample_num=60, concept_num=2 if algo in ("mmgeni", "mmgeniex", "clusterfl") else 4)
    sim = DriftSim(make_args(**kw), device="cpu", sink=MetricsSink())
    out = sim.run()
    assert len(out["history"]) >= 3
    for h in out["history"]:
        assert 0.0 <= h["train_acc"] <= 1.0 and 0.0 <= h["test_acc"] <= 1.0 and np.isfinite(h["train_loss"])
    assert torch.isfinite(sim.bank.theta).all()


def test_h_f_variant_starts_with_local_models():
    sim = DriftSim(make_args(concept_drift_algo_arg="H_A_F_1_10_0", concept_num=10, comm_round=3,
                             total_train_iteration=2, sample_num=60), device="cpu", sink=MetricsSink())
    sim.begin_time_step(0)
    assert sim.algo.state.test_model_indices(0).tolist() == list(range(10))
    sim.run_rounds(3)
    sim.end_time_step()
    sim.run_time_step(1)   # hierarchical clustering merges the identical-concept local models
    assert len(set(sim.algo.state.test_model_indices(1).tolist())) < 10


def test_kue_masks_only_grow():
    k = KueState(3, 5, np.random.RandomState(0))
    before = k.masks.copy()
    k.initialize_mask(1)
    assert (k.masks | before == k.masks).all() and k.masks.any(1).all()


def test_generic_path_equals_fused_reference_path():
    """The any-model executor and the fused-kernel reference semantics pick the same batches (shared RNG) and
    must produce the same models."""
    a = make_args(comm_round=3, total_train_iteration=2, sample_num=60)
    s1 = DriftSim(a, device="cpu", sink=MetricsSink())
    s1.run()
    s2 = DriftSim(make_args(comm_round=3, total_train_iteration=2, sample_num=60), device="cpu", sink=MetricsSink())
    s2.algo.fused_ok = lambda: False
    s2.run()
    assert torch.allclose(s1.bank.theta, s2.bank.theta, rtol=1e-4, atol=1e-5)
    assert abs(s1.history[-1]["test_acc"] - s2.history[-1]["test_acc"]) < 1e-6


def test_batched_acc_matrix_equals_per_client_inference_for_module_models():
    """Evaluator.acc_matrix on the nn.Module path (one batched forward per model) vs M·C separate `infer_client` calls."""
    from feddrift_b200.sim import DriftSim, make_args
    sim = DriftSim(make_args(model="cnn", dataset="MNIST", client_num_in_total=5, client_num_per_round=5, sample_num=12,
                             batch_size=4, comm_round=1, total_train_iteration=2, concept_drift_algo="win-1"), device="cpu")
    ev = sim.evaluator
    for m in range(sim.M):
        sim.bank.reset_parameters_random(m, torch.Generator().manual_seed(m))
    sim.data.nsamp[0, 2] = 7          # a partially filled client and an empty one
    sim.data.nsamp[0, 4] = 0
    got = ev.acc_matrix(list(range(sim.M)), 0)
    for m in range(sim.M):
        for c in range(5):
            k, n, _ = ev.infer_client(m, c, 0)
            assert abs(got[m, c] - (k / n if n else 0.0)) < 1e-6, (m, c)


@pytest.mark.parametrize("ens_mode", [1, 2])
def test_grouped_ensemble_evaluation_equals_per_client_votes(ens_mode):
    """generic._eval_ensemble_grouped (one batched forward per member) vs per-client ops.ensemble_vote / soft_vote."""
    from feddrift_b200 import ops
    from feddrift_b200.sim import DriftSim, generic, make_args
    sim = DriftSim(make_args(model="cnn", dataset="MNIST", client_num_in_total=5, client_num_per_round=5, sample_num=12,
                             batch_size=4, comm_round=1, total_train_iteration=2, concept_drift_algo="aue", ensemble_window=3),
                   device="cpu")
    for m in range(sim.M):
        sim.bank.reset_parameters_random(m, torch.Generator().manual_seed(10 + m))
    sim.data.nsamp[1, 3] = 5
    sim.data_host.nsamp[1, 3] = 5
    g = torch.Generator().manual_seed(0)
    w = torch.rand(5, sim.M, generator=g)
    w[2, 0] = 0.0
    out = torch.zeros(5, 4)
    generic._eval_ensemble_grouped(sim, {"ens_w": w}, 1, list(range(5)), out, ens_mode)
    for c in range(5):
        n1 = int(sim.data_host.nsamp[1, c])
        x1, y1 = sim.data.X[1, c, :n1], sim.data.Y[1, c, :n1]
        ks = [k for k in range(sim.M) if float(w[c, k]) > 0]
        with torch.no_grad():
            if ens_mode == 1:
                vote = ops.ensemble_vote(torch.stack([sim.bank.forward(k, x1).argmax(-1) for k in ks]), w[c, ks], sim.data.class_num)
            else:
                vote = ops.soft_vote(torch.stack([torch.softmax(sim.bank.forward(k, x1), 1) for k in ks]), w[c, ks])
        assert float(out[c, 2]) == float((vote == y1).sum()), (c, ens_mode)
