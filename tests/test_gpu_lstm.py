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
