"""GPU numerics: the fused persistent round kernel vs the fp32 PyTorch reference of the same op."""
import copy

import pytest
import torch

from feddrift_b200 import ops
from feddrift_b200.ops import reference as ref

pytestmark = pytest.mark.gpu


def make_state(kind="fnn", din=3, hid=6, dout=2, T1=4, C=10, S=100, M=4, t=2, B=500, epochs=5, optimizer="adam", seed=3,
               mode="pool"):
    g = torch.Generator().manual_seed(seed)
    P = ref.mlp_param_count(kind, din, hid, dout)
    X = torch.rand(T1, C, S, din, generator=g) * (10.0 if din == 3 else 1.0)
    Y = (X[..., 1] + X[..., -1] > (8.0 if din == 3 else 1.0)).long() % dout
    nsamp = torch.full((T1, C), S, dtype=torch.int32)
    nsamp[1, 3] = S - 7
    W = torch.zeros(t + 1, M, C)
    for tt in range(t + 1):
        for c in range(C):
            W[tt, (c + tt) % 2 if tt else 0, c] = 1.0
    W[t, 2, 5] = 1.0
    W[t, 0, 5] = 0.0
    theta = torch.randn(M, P, generator=g) * 0.3
    st = dict(kind=kind, din=din, hid=hid, dout=dout, X=X, Y=Y, nsamp=nsamp, batch_size=B, W=W, theta=theta,
              opt_m=torch.zeros(C, M, P), opt_v=torch.zeros(C, M, P), opt_vmax=torch.zeros(C, M, P),
              opt_step=torch.zeros(C, M, dtype=torch.int32), lr=0.01, wd=0.001, epochs=epochs, optimizer=optimizer,
              seed=1234, round0=0, t_cur=t, sample_mode=mode)
    if mode == "time":
        st["W"] = torch.tensor([2.0 ** i for i in range(t + 1)])[:, None, None].expand(t + 1, M, C).contiguous().clone()
        st["W"][:, 1:, :] = 0
    if mode == "index":
        L = 2 * S
        ti = torch.zeros(M, C, L, dtype=torch.int32)
        tc = torch.zeros(M, C, dtype=torch.int32)
        for m in range(2):
            for c in range(C):
                idx = torch.cat([(t - 1) * S + torch.randperm(S, generator=g), t * S + torch.randperm(S, generator=g)])
                n = L if m == 0 else S
                ti[m, c, :n] = idx[:n].int()
                tc[m, c] = n
        st["train_index"], st["train_count"] = ti, tc
    return st


def to_cuda(st):
    out = {}
    for k, v in st.items():
        out[k] = v.cuda() if isinstance(v, torch.Tensor) else v
    out["Y"] = out["Y"].int()
    return out


@pytest.mark.parametrize("cfg", [
    dict(), dict(optimizer="sgd"), dict(kind="lr", hid=0), dict(din=2, hid=4), dict(B=32), dict(mode="time"),
    dict(mode="index", B=64), dict(C=37, M=4), dict(kind="fnn", din=4, hid=8, dout=3),
])
def test_fused_round_matches_reference(cfg):
    st_cpu = make_state(**cfg)
    st_gpu = to_cuda(copy.deepcopy(st_cpu))
    rounds = 3
    out_ref = ref.fed_round_small(st_cpu, rounds)
    out_gpu = ops.fed_round_small(st_gpu, rounds)
    torch.cuda.synchronize()
    assert torch.allclose(st_gpu["theta"].cpu(), st_cpu["theta"], rtol=2e-4, atol=2e-5), \
        (st_gpu["theta"].cpu() - st_cpu["theta"]).abs().max()
    assert torch.equal(st_gpu["opt_step"].cpu(), st_cpu["opt_step"])
    if st_cpu["optimizer"] == "adam":
        assert torch.allclose(st_gpu["opt_m"].cpu(), st_cpu["opt_m"], rtol=1e-3, atol=1e-6)
    mg, mr = out_gpu["metrics"].cpu(), out_ref["metrics"]
    assert (mg[..., 0] - mr[..., 0]).abs().max() <= 1.0  # correct counts (±1 sample at a decision boundary)
    assert torch.allclose(mg[..., 1], mr[..., 1], rtol=1e-3, atol=1e-2)
    assert (mg[..., 2] - mr[..., 2]).abs().max() <= 1.0
    assert torch.allclose(out_gpu["counts"].cpu(), out_ref["counts"])


def test_fused_round_ifca_recluster():
    st_cpu = make_state(M=3)
    st_cpu["recluster_hard"] = True
    st_gpu = to_cuda(copy.deepcopy(st_cpu))
    ref.fed_round_small(st_cpu, 2)
    ops.fed_round_small(st_gpu, 2)
    # after re-clustering every client is on exactly one model
    Wg = st_gpu["W"][st_gpu["t_cur"]].cpu()
    assert torch.all(Wg.sum(0) == 1)
    agree = (Wg.argmax(0) == st_cpu["W"][st_cpu["t_cur"]].argmax(0)).float().mean()
    assert agree >= 0.8


def test_fused_round_ensemble_and_eval_override():
    st_cpu = make_state()
    C, M = st_cpu["X"].shape[1], st_cpu["theta"].shape[0]
    st_cpu["ens_mode"] = 1
    st_cpu["ens_w"] = torch.rand(C, M)
    st_cpu["eval_train_model"] = torch.zeros(C, dtype=torch.int32)
    st_gpu = to_cuda(copy.deepcopy(st_cpu))
    o_ref = ref.fed_round_small(st_cpu, 2)
    o_gpu = ops.fed_round_small(st_gpu, 2)
    assert (o_gpu["metrics"].cpu()[..., 2] - o_ref["metrics"][..., 2]).abs().max() <= 2.0
    assert torch.allclose(o_gpu["metrics"].cpu()[..., 1], o_ref["metrics"][..., 1], rtol=1e-3, atol=1e-2)


def test_eval_matrix_kernel():
    st = make_state()
    theta, X, Y, ns = st["theta"], st["X"][1], st["Y"][1], st["nsamp"][1]
    c_ref, l_ref = ref.mlp_eval_matrix(theta, X, Y, ns, "fnn", 3, 6, 2)
    c_gpu, l_gpu = ops.mlp_eval_matrix(theta.cuda(), X.cuda(), Y.cuda(), ns.cuda(), "fnn", 3, 6, 2)
    assert (c_gpu.cpu() - c_ref).abs().max() <= 1.0
    assert torch.allclose(l_gpu.cpu(), l_ref, rtol=1e-4, atol=1e-3)


def test_drift_sim_runs_on_gpu_with_native_kernel():
    from feddrift_b200.ops import small_round
    from feddrift_b200.sim import DriftSim, make_args
    before = small_round.LAUNCH_COUNT["fed_round_small"]
    sim = DriftSim(make_args(comm_round=20, total_train_iteration=4), device="cuda")
    out = sim.run()
    assert small_round.LAUNCH_COUNT["fed_round_small"] > before
    assert out["history"][-1]["train_acc"] > 0.7


def test_end_to_end_round_graph_with_fused_host_io_matches_copy_path():
    """`run_round(host_inputs, use_graph=True)` (ONE kernel node: the kernel pulls the inputs from pinned host memory and
    mirrors the metrics into pinned host memory) must equal the explicit H2D-copy / D2H-copy path, and must really read
    the host buffers on every replay."""
    from feddrift_b200.sim import DriftSim, make_args

    def make():
        sim = DriftSim(make_args(comm_round=6, total_train_iteration=4), device="cuda")
        for t in range(2):
            sim.run_time_step(t, rounds=4)
        sim.begin_time_step(2)
        sim.args.rounds_per_launch = 1
        return sim

    a, b = make(), make()
    ha, hb = a.make_host_round_inputs(), b.make_host_round_inputs()
    assert ha["X"].is_pinned() and torch.equal(ha["X"], hb["X"])
    for _ in range(3):
        ra = a.run_round(ha, use_graph=True)
        rb = b.run_round(hb, use_graph=False)
        for k in ("train_acc", "train_loss", "test_acc", "test_loss"):
            assert abs(ra[k] - rb[k]) < 1e-5, (k, ra, rb)
    assert torch.allclose(a.bank.theta, b.bank.theta, atol=1e-6)
    # new data in the SAME pinned buffers must be picked up by the next replay (the graph holds only the pointers)
    ha["Y"].copy_(1 - ha["Y"])
    hb["Y"].copy_(1 - hb["Y"])
    ra = a.run_round(ha, use_graph=True)
    rb = b.run_round(hb, use_graph=False)
    assert abs(ra["train_acc"] - rb["train_acc"]) < 1e-5 and ra["train_acc"] < 0.5
    assert torch.allclose(a.bank.theta, b.bank.theta, atol=1e-6)


def test_generic_executor_graphed_pairs_match_eager(monkeypatch):
    """The CUDA-graphed per-pair local training (in-graph minibatch gathers, 8 concurrent streams) must reproduce the
    eager execution of the same ops (fnn-MNIST: no dropout, so both runs consume identical randomness)."""
    from feddrift_b200.sim import DriftSim, make_args
    from feddrift_b200.utils.metrics import MetricsSink
    kw = dict(model="fnn", dataset="MNIST", client_num_in_total=6, client_num_per_round=6, concept_drift_algo="softcluster",
              concept_drift_algo_arg="H_A_C_1_10_0", concept_num=2, change_points="A", sample_num=16, batch_size=8, comm_round=2,
              total_train_iteration=2, epochs=3)
    monkeypatch.setenv("FDB_NO_GRAPHS", "1")
    eager = DriftSim(make_args(**kw), device="cuda", sink=MetricsSink())
    out_e = eager.run()
    monkeypatch.delenv("FDB_NO_GRAPHS")
    graphed = DriftSim(make_args(**kw), device="cuda", sink=MetricsSink())
    out_g = graphed.run()
    assert any(g.indexed and g.launches > 0 for g in graphed.__dict__.get("_step_graphs", {}).values()), "per-pair graphs not used"
    assert (eager.bank.theta - graphed.bank.theta).abs().max().item() < 5e-3
    assert abs(out_e["history"][-1]["train_loss"] - out_g["history"][-1]["train_loss"]) < 5e-2
