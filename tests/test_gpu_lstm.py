"""Fused persistent 2-layer LSTM kernel (csrc/lstm_tc.cu) vs the plain fp32 nn.Embedding + nn.LSTM reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_and_fused(B, T, need_all, seed=0):
    from feddrift_b200.ops import lstm as fused
    torch.manual_seed(seed)
    emb = torch.nn.Embedding(90, 8, padding_idx=0).cuda()
    lstm = torch.nn.LSTM(8, 256, num_layers=2, batch_first=True).cuda()
    tok = torch.randint(0, 90, (B, T), device="cuda")
    assert fused.eligible(tok, emb.weight, lstm)
    params = [emb.weight] + list(lstm.parameters())
    # reference (fp32, cuDNN/ATen)
    out, _ = lstm(emb(tok))
    ref = out if need_all else out[:, -1]
    wgt = torch.randn_like(ref)
    (ref * wgt).sum().backward()
    gref = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    n0 = fused.CALLS["fwd"], fused.CALLS["bwd"]
    got = fused.lstm2_embed_forward(tok, emb, lstm, need_all=need_all)
    (got * wgt).sum().backward()
    torch.cuda.synchronize()
    assert fused.CALLS["fwd"] == n0[0] + 1 and fused.CALLS["bwd"] == n0[1] + 1
    ggot = [p.grad.clone() for p in params]
    return ref.detach(), got.detach(), gref, ggot, ["emb"] + [n for n, _ in lstm.named_parameters()]


@pytest.mark.parametrize("B,T,need_all", [(16, 12, False), (5, 7, False), (16, 80, False), (20, 9, True)])
def test_fused_lstm_matches_reference(B, T, need_all):
    ref, got, gref, ggot, names = _ref_and_fused(B, T, need_all)
    # bf16 operands / fp32 accumulation: compare against the fp32 reference with bf16-level tolerances
    err = (ref - got).abs().max().item()
    assert err < 2e-2, f"forward max abs err {err}"
    for n, a, b in zip(names, gref, ggot):
        denom = a.abs().max().item() + 1e-6
        rel = (a - b).abs().max().item() / denom
        assert rel < 4e-2, f"grad {n}: rel err {rel} (scale {denom})"


def test_rnn_model_uses_fused_kernel_and_trains():
    from feddrift_b200.models.rnn import RNN_OriginalFedAvg
    from feddrift_b200.ops import lstm as fused
    torch.manual_seed(1)
    m = RNN_OriginalFedAvg().cuda()
    x = torch.randint(1, 90, (16, 20), device="cuda")
    y = torch.randint(0, 90, (16,), device="cuda")
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    n0 = fused.CALLS["fwd"]
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert fused.CALLS["fwd"] == n0 + 8
    assert losses[-1] < losses[0] - 0.05, losses


def _run_rnn_federation(env):
    import os
    from feddrift_b200.sim import DriftSim, make_args
    from feddrift_b200.utils.metrics import MetricsSink
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        kw = dict(model="rnn", dataset="shakespeare", client_num_in_total=6, client_num_per_round=6, concept_drift_algo="win-1",
                  concept_drift_algo_arg="", concept_num=2, change_points="A", sample_num=32, batch_size=16, comm_round=2,
                  total_train_iteration=2, epochs=2, lr=0.05, client_optimizer="sgd", report_client=0)
        sim = DriftSim(make_args(**kw), device="cuda:0", sink=MetricsSink())
        init = sim.bank.theta.clone()
        out = sim.run_time_step(0, rounds=2)
        torch.cuda.synchronize()
        return init, sim.bank.theta.clone(), out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_batched_lstm_executor_matches_per_pair_cudnn_path():
    """All pairs in one launch (sim/lstm_exec.py) vs the per-pair nn.LSTM (cuDNN, fp32) executor: same federated update."""
    from feddrift_b200.ops import lstm as fused
    n0 = fused.CALLS["bwd"]
    init, th_b, out_b = _run_rnn_federation({"FDB_LSTM_BATCHED": "1"})
    assert fused.CALLS["bwd"] > n0, "batched executor did not run the fused BPTT kernel"
    _, th_r, out_r = _run_rnn_federation({"FDB_LSTM_BATCHED": "0", "FDB_NO_FUSED_LSTM": "1"})
    upd_b, upd_r = th_b - init, th_r - init
    scale = upd_r.abs().max().item()
    assert scale > 1e-4
    err = (upd_b - upd_r).abs().max().item() / scale
    assert err < 6e-2, f"federated update differs: rel {err}"
    assert abs(out_b["train_loss"] - out_r["train_loss"]) < 5e-2 * max(1.0, abs(out_r["train_loss"]))
