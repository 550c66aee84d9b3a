import numpy as np
import torch

from feddrift_b200.data import changepoints as cp
from feddrift_b200.data.drift import (DriftData, generate_drift_data, load_all_data, load_partition_data,
                                      select_iterations)
from feddrift_b200.models import create_model, reinitialize
from feddrift_b200.models.utils import flat_size, flatten_state_dict, unflatten_to_state_dict, flat_spec
from feddrift_b200.parallel.arena import ModelBank


def test_changepoint_tables_and_random():
    A = cp.named("A")
    assert A.shape == (11, 10) and A.max() == 1 and A[0].sum() == 0
    assert cp.named("B").max() == 3
    assert cp.load("A", 10, 25).shape == (11, 25)          # tiled clients
    assert cp.load("A", 20, 10).shape[0] >= 21              # stationary tail
    r = cp.load("rand", 10, 10, rng=np.random.RandomState(0))
    assert set(np.unique(r)) <= {0, 1} and (np.diff(r, axis=0) >= 0).all()


def test_generators_follow_concepts():
    d = generate_drift_data("sea", 4, 10, 200, 0.0, 1, "A", sea_label_noise=0.0)
    assert d.X.shape == (5, 10, 200, 3) and d.class_num == 2
    x, y = d.X[4, 1], d.Y[4, 1]   # A.cp: client 1 is on concept 1 (θ=9) at step 4
    assert torch.equal(y, (x[:, 1] + x[:, 2] > 9.0).long())
    for name in ("sine", "circle"):
        dd = generate_drift_data(name, 3, 4, 50, 0.1, 1, "rand")
        assert dd.X.shape == (4, 4, 50, 2) and 0 < dd.Y.float().mean() < 1
    m = generate_drift_data("MNIST", 2, 3, 40, 0.0, 1, "B")
    assert m.X.shape == (3, 3, 40, 784) and m.class_num == 10


def test_retrain_selectors_and_fedml_tuple(tmp_path):
    assert select_iterations("all", 3) == [0, 1, 2, 3]
    assert select_iterations("win-2", 3) == [2, 3]
    assert select_iterations("weight-linear", 2) == [0, 1, 1, 2, 2, 2]
    assert select_iterations("weight-exp", 2) == [0, 1, 1, 2, 2, 2, 2]
    assert select_iterations("sel-0,2", 5) == [0, 2]
    assert select_iterations("clientsel-[[0],[1,2]]", 5, 1) == [1, 2]
    d = generate_drift_data("sea", 3, 4, 30, 0.0, 1, "rand")
    tup = load_partition_data(d, 16, 1, "win-2", rng=np.random.RandomState(0))
    assert tup[0] == 4 and tup[1] == 4 * 60 and tup[2] == 4 * 30 and len(tup[6][0]) == 4 and tup[8] == 2
    allb = load_all_data(d, 16, 2)
    assert len(allb) == 4 and len(allb[0]) == 3 and allb[0][0][0][0].shape == (16, 3)
    d.to_csv_dir(str(tmp_path))
    back = DriftData.from_csv_dir(str(tmp_path), "sea", 4, 4, 2)
    assert torch.allclose(back.X, d.X, atol=1e-5) and torch.equal(back.Y, d.Y)


def test_models_param_counts_and_reinit_identity():
    assert sum(p.numel() for p in create_model("fnn", 2, 3).parameters()) == 38
    assert sum(p.numel() for p in create_model("lr", 10, 784).parameters()) == 7850
    assert sum(p.numel() for p in create_model("fnn", 10, 784).parameters()) == 1246570
    assert sum(p.numel() for p in create_model("cnn", 10).parameters()) == 1199882
    assert sum(p.numel() for p in create_model("cnn_fedavg", 10).parameters()) == 1663370
    assert sum(p.numel() for p in create_model("rnn", 90).parameters()) == 822570
    a, b = create_model("fnn", 2, 3), create_model("fnn", 2, 3)
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))  # reseeded → identical
    with torch.no_grad():
        a.fc1.weight.add_(1)
    reinitialize(a)
    assert torch.equal(a.fc1.weight, b.fc1.weight)
    x = torch.randn(5, 784)
    assert create_model("cnn", 10)(x).sum(1).allclose(torch.ones(5), atol=1e-5)  # softmax output (quirk)
    assert (create_model("lr", 2, 3)(torch.randn(4, 3)) >= 0).all()               # sigmoid output (quirk)


def test_model_bank_views_and_flat_roundtrip():
    tmpl = create_model("fnn", 2, 3)
    bank = ModelBank(tmpl, 3)
    assert bank.P == 38 and bank.stride % 32 == 0
    sd = bank.state_dict(1)
    sd["fc1.bias"].add_(1.0)                      # views alias the arena row
    assert torch.allclose(bank.theta[1][18:24], bank.init_row[18:24] + 1)
    mod = bank.module(1)
    x = torch.randn(7, 3)
    assert torch.allclose(mod(x), bank.forward(1, x), atol=1e-6)
    bank.copy(2, 1)
    bank.reinit(1)
    assert torch.equal(bank.theta[1], bank.init_row) and not torch.equal(bank.theta[2], bank.init_row)
    flat = flatten_state_dict(tmpl.state_dict())
    back = unflatten_to_state_dict(flat, flat_spec(tmpl))
    assert all(torch.equal(back[k], v) for k, v in tmpl.state_dict().items()) and flat_size(tmpl) == 38


def test_tcconv2d_is_a_drop_in_for_nn_conv2d():
    """Same state-dict keys / init law / CPU numerics as nn.Conv2d (the CUDA path is checked in tests/test_gpu_kernels.py)."""
    import torch
    from torch import nn
    from feddrift_b200.ops.conv import TcConv2d
    torch.manual_seed(3)
    a = TcConv2d(8, 16, 3, stride=2, padding=1)
    torch.manual_seed(3)
    b = nn.Conv2d(8, 16, 3, stride=2, padding=1)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    assert torch.equal(a.weight, b.weight) and torch.equal(a.bias, b.bias)
    x = torch.randn(2, 8, 9, 9)
    assert torch.allclose(a(x), b(x), atol=1e-6)
    from feddrift_b200.models.cnn import CNN_DropOut
    assert sum(p.numel() for p in CNN_DropOut().parameters()) == 1_199_882


def test_flat_arena_layout_aligns_big_tensors_and_round_trips():
    import torch
    from feddrift_b200.models import create_model
    from feddrift_b200.models import utils as mu
    for name, kw in (("cnn", {}), ("resnet56", {}), ("rnn", {})):
        net = create_model(name, 10, 784, **kw)
        spec = mu.flat_spec(net)
        assert all(off % 4 == 0 for _, _, _, off, n in spec if n >= 256), name
        sd = net.state_dict()
        flat = mu.flatten_state_dict(sd)
        assert flat.numel() == mu.flat_size(net)
        back = mu.unflatten_to_state_dict(flat, spec)
        assert all(torch.equal(back[k].float(), v.float()) for k, v in sd.items()), name
    mlp = create_model("fnn", 2, 3)   # SEA fnn: dense W1 | b1 | W2 | b2 (the fused kernel's layout), P = 38
    assert mu.flat_size(mlp) == 38 and [s[3] for s in mu.flat_spec(mlp)] == [0, 18, 24, 36]


def test_tensor_core_conv_weights_are_stored_channels_last_in_the_flat_rows():
    """models.utils.ohwi_stored: eligible conv weights live in the rows as (O, kh, kw, I); unflatten returns logical OIHW views with
    channels_last strides; stems / 1x1 / depthwise filters stay in logical order; ModelBank consumes them through TcConv2d."""
    import torch
    from torch import nn
    from feddrift_b200.models.utils import flat_spec, flat_view, flatten_state_dict, ohwi_stored, unflatten_to_state_dict
    from feddrift_b200.ops.conv import TcConv2d
    from feddrift_b200.parallel.arena import ModelBank
    assert ohwi_stored((64, 32, 3, 3)) and ohwi_stored((128, 64, 5, 5))
    assert not ohwi_stored((64, 3, 3, 3)) and not ohwi_stored((128, 64, 1, 1)) and not ohwi_stored((64, 1, 3, 3)) and not ohwi_stored((64, 32, 1, 7))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = nn.Conv2d(3, 32, 3, padding=1)
            self.body = nn.Conv2d(32, 64, 3, padding=1)
            self.head = nn.Linear(64, 5)

        def forward(self, x):
            return self.head(self.body(self.stem(x)).mean((2, 3)))
    torch.manual_seed(0)
    net = Net()
    sd = net.state_dict()
    flat = flatten_state_dict(sd)
    spec = {k: (shape, off, n) for k, shape, _, off, n in flat_spec(net)}
    shape, off, n = spec["body.weight"]
    assert torch.equal(flat[off:off + n], sd["body.weight"].permute(0, 2, 3, 1).reshape(-1))      # (O, kh, kw, I) in the row
    shape, off, n = spec["stem.weight"]
    assert torch.equal(flat[off:off + n], sd["stem.weight"].reshape(-1))                          # logical order
    back = unflatten_to_state_dict(flat, flat_spec(net))
    for k in sd:
        assert torch.equal(back[k], sd[k])
    assert back["body.weight"].is_contiguous(memory_format=torch.channels_last) and not back["body.weight"].is_contiguous()
    assert torch.equal(flat_view(back["body.weight"]), flat[spec["body.weight"][1]:spec["body.weight"][1] + spec["body.weight"][2]])
    bank = ModelBank(net, 2)
    assert isinstance(bank.template.body, TcConv2d) and isinstance(bank.template.stem, nn.Conv2d) and not isinstance(bank.template.stem, TcConv2d)
    assert list(bank.template.state_dict().keys()) == list(sd.keys())
    mod = bank.module(1)
    x = torch.randn(2, 3, 6, 6)
    net2 = Net()
    net2.load_state_dict(bank.state_dict(1))
    assert torch.allclose(mod(x), net2(x), atol=1e-5)
    mod(x).sum().backward()                                                                        # CPU backward through the channels_last view
    assert mod.body.weight.grad is not None


def test_checkpoint_v1_rows_are_upgraded_to_the_channels_last_row_layout():
    import torch
    from torch import nn
    from feddrift_b200.models.utils import flat_spec, flatten_state_dict
    from feddrift_b200.sim.checkpoint import upgrade_theta_v1

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Conv2d(3, 32, 3)
            self.b = nn.Conv2d(32, 64, 3)
            self.fc = nn.Linear(64, 4)
    net = Net()
    sd = net.state_dict()
    v1 = torch.cat([v.reshape(-1) for v in sd.values()])          # format 1: every tensor in logical order, dense (offsets aligned here)
    spec = flat_spec(net)
    assert all(spec[i][3] + spec[i][4] == spec[i + 1][3] for i in range(len(spec) - 1))
    assert torch.equal(upgrade_theta_v1(v1[None], spec)[0], flatten_state_dict(sd))
