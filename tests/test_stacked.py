"""Pair-stacked executor (sim/stacked.py): the channel-stacked network computes, for every (client, model) pair at once, exactly
the per-pair forward / gradients / buffer updates; end to end a round trained through it equals the per-pair executor."""
import copy
import os

import pytest
import torch
import torch.nn.functional as F

from feddrift_b200.models import cnn, resnet
from feddrift_b200.models.utils import flat_size, flat_spec, flat_view, flatten_state_dict, unflatten_to_state_dict
from feddrift_b200.sim import stacked as S


def _fresh_rows(tmpl, n):
    rows = []
    for i in range(n):
        m = copy.deepcopy(tmpl)
        torch.manual_seed(100 + i)
        for layer in m.modules():
            if layer is not m and hasattr(layer, "reset_parameters"):
                layer.reset_parameters()
        rows.append(flatten_state_dict(m.state_dict()))
    return torch.stack(rows)


def _check(tmpl, xshape, n=3, B=4, tol=1e-4):
    torch.manual_seed(0)
    spec, P = flat_spec(tmpl), flat_size(tmpl)
    rows = _fresh_rows(tmpl, n)
    assert rows.shape[1] == P
    x, y = torch.randn(n, B, *xshape), torch.randint(0, 10, (n, B))
    gref, outs, newrows = torch.zeros(n, P), [], rows.clone()
    for i in range(n):                                            # reference: one network per pair
        mod = copy.deepcopy(tmpl)
        mod.load_state_dict(unflatten_to_state_dict(rows[i].clone(), spec))
        mod.train()
        out = mod(x[i])
        outs.append(out)
        F.cross_entropy(out, y[i]).backward()
        grads = {k: p.grad for k, p in mod.named_parameters()}
        for k, _, _, off, numel in spec:
            if grads.get(k) is not None:
                gref[i, off:off + numel] = flat_view(grads[k])
        newrows[i] = flatten_state_dict(mod.state_dict())          # BN running statistics moved
    net = S.stack_module(tmpl, n)
    stage, G = rows.clone(), torch.zeros(n, P)
    sp = {k: (tuple(shape), off, numel) for k, shape, _, off, numel in spec}
    for name, mod in net.named_modules():
        if isinstance(mod, S._Stacked):
            mod.bind(name, sp, stage, G)
    net.train()
    st = S._Stage.__new__(S._Stage)
    st.params, st.spec, st.npairs = stage, sp, n
    st.bns = [(nm, m) for nm, m in net.named_modules() if isinstance(m, S.StackedBatchNorm2d) and m.track]
    st.load_buffers()
    logits = net(S.stack_input(tmpl, x))
    K = logits.shape[1] // n
    ref = torch.stack(outs, 1).reshape(B, n * K)
    assert (logits - ref).abs().max().item() <= tol * (1 + ref.abs().max().item())
    (F.cross_entropy(logits.reshape(B * n, K), y.t().reshape(-1), reduction="sum") / B).backward()
    assert (G - gref).abs().max().item() <= tol * (1 + gref.abs().max().item())
    st.store_buffers()
    assert (stage - newrows).abs().max().item() <= tol * (1 + newrows.abs().max().item())


def test_stacked_cnn_dropout_free_matches_per_pair():
    c = cnn.CNN_DropOut()
    c.dropout_1.p = c.dropout_2.p = 0.0
    _check(c, (784,))
    _check(cnn.CNN_OriginalFedAvg(), (28, 28))


def test_stacked_resnet_batchnorm_and_groupnorm_match_per_pair():
    _check(resnet.ResNet(resnet.BasicBlock, [1, 1, 1], 10, widths=(8, 16, 32)), (3, 16, 16))
    _check(resnet.ResNet(resnet.BasicBlock, [1, 1], 10, widths=(32, 64), norm=resnet._gn(2) if hasattr(resnet, "_gn") else None), (3, 8, 8),
           n=2, tol=1e-3)


def test_stackable_rejects_unsupported_layers():
    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(4, 4)

        def forward(self, x):
            return self.emb(x)
    assert not S.stackable(Odd())
    assert S.stackable(cnn.CNN_DropOut())
    cma = resnet.ResNet(resnet.BasicBlock, [1], 10, widths=(8,))
    cma.bn1.momentum = None                                   # cumulative moving average: semantics the stacked BN does not reproduce
    assert not S.stackable(cma)
    with pytest.raises(TypeError):
        S.stack_module(Odd(), 2)


def _run_round(stacked_on: bool, optimizer: str = "sgd", max_gb: str = None):
    from feddrift_b200.sim import DriftSim, make_args
    from feddrift_b200.utils.metrics import MetricsSink
    os.environ["FDB_STACKED"] = "1" if stacked_on else "0"
    if max_gb is not None:
        os.environ["FDB_STACKED_MAX_GB"] = max_gb
    try:
        a = make_args(model="cnn", dataset="MNIST", client_num_in_total=4, client_num_per_round=4, concept_drift_algo="win-1",
                      concept_drift_algo_arg="", concept_num=2, change_points="A", sample_num=16, batch_size=8, comm_round=1,
                      total_train_iteration=2, epochs=2, lr=0.05, report_client=0, client_optimizer=optimizer)
        sim = DriftSim(a, device="cpu", sink=MetricsSink())
        for mod in (sim.bank.template.dropout_1, sim.bank.template.dropout_2):
            mod.p = 0.0
        sim.run_time_step(0, rounds=1)
        return sim.bank.theta.clone(), sim.clients.params.clone(), sim.clients.step.clone()
    finally:
        os.environ.pop("FDB_STACKED", None)
        os.environ.pop("FDB_STACKED_MAX_GB", None)


def test_round_through_stacked_executor_equals_per_pair_executor():
    # SGD: the update is linear in the gradient, so the two executors agree to rounding.  (Adam's m/sqrt(v) turns rounding noise
    # on zero-gradient parameters — dead MNIST border pixels — into ±lr steps, so it is compared on the loss level only.)
    th1, cp1, s1 = _run_round(True)
    th0, cp0, s0 = _run_round(False)
    assert torch.equal(s1, s0)
    assert (cp1 - cp0).abs().max().item() < 1e-5
    assert (th1 - th0).abs().max().item() < 1e-5
    a1, _, st1 = _run_round(True, "adam")
    a0, _, st0 = _run_round(False, "adam")
    assert torch.equal(st1, st0) and int(st1.max()) == 2
    assert (a1 - a0).abs().max().item() < 0.05


def test_large_federations_are_stacked_in_memory_bounded_passes():
    """FDB_STACKED_MAX_GB caps the staged rows: 4 pairs of the 1.2 M-parameter CNN under a 0.05 GB budget → two passes of two pairs,
    same result as one pass."""
    from feddrift_b200.sim import stacked as S2
    os.environ["FDB_STACKED_MAX_GB"] = "0.05"
    try:
        assert S2.max_pairs_per_pass(1199884) == 2
    finally:
        os.environ.pop("FDB_STACKED_MAX_GB", None)
    th_one, cp_one, _ = _run_round(True)
    th_two, cp_two, _ = _run_round(True, max_gb="0.05")
    assert (cp_one - cp_two).abs().max().item() < 1e-5
    assert (th_one - th_two).abs().max().item() < 1e-5
