from types import SimpleNamespace

import numpy as np
import torch

from feddrift_b200.data import benchmarks as B


def test_lda_partition_is_a_partition_and_skewed():
    rng = np.random.RandomState(0)
    y = rng.randint(0, 10, 5000)
    m = B.partition_indices(y, 8, "hetero", 0.3, rng, 10)
    allidx = np.concatenate(list(m.values()))
    assert len(allidx) == 5000 and len(np.unique(allidx)) == 5000 and min(len(v) for v in m.values()) >= 10
    stats = B.record_net_data_stats(y, m)
    assert max(max(s.values()) / sum(s.values()) for s in stats.values()) > 0.25      # non-iid
    h = B.partition_indices(y, 8, "homo", rng=rng)
    assert sorted(len(v) for v in h.values()) == [625] * 8


def test_loaders_return_fedml_tuples_of_the_named_shapes():
    a = SimpleNamespace(batch_size=8, client_num_in_total=4, partition_method="hetero", partition_alpha=0.5, data_dir=None)
    ds = B.load_data(a, "cifar10")
    assert ds[7] == 10 and ds[5][0][0][0].shape[1:] == (3, 32, 32) and len(ds[5]) == 4
    sh = B.load_data(a, "fed_shakespeare")
    x, y = sh[5][0][0]
    assert x.shape[1] == 80 and x.dtype == torch.int64 and y.dim() == 1 and sh[7] == 90
    xs, ys = B.load_partition_data_shakespeare(8, 2, 10, per_position=True)[6][0][0]
    assert ys.shape == (8, 80)
    lr = B.load_data(a, "stackoverflow_lr")
    assert lr[5][0][0][1].shape[1] == 500
    assert B.load_data(a, "mnist")[7] == 10 and B.load_data(a, "femnist")[7] == 62 and B.load_data(a, "fed_cifar100")[7] == 100
    assert B.load_data(a, "stackoverflow_nwp")[7] == 10004


def test_streaming_vertical_synthetic_and_cutout():
    s = B.load_streaming_susy_or_ro(3, 5, "SUSY")
    assert len(s) == 3 and len(s[0]) == 5 and s[0][0]["x"].shape == (18,)
    Xtr, ytr, Xte, yte = B.load_vertical_parties("lending_club_loan", 200, 3)
    assert len(Xtr) == 3 and len(ytr) == 160 and Xte[0].shape[0] == 40
    Xs, Ys = B.generate_synthetic(1.0, 1.0, False, num_user=5)
    assert len(Xs) == 5 and Xs[0].shape[1] == 60 and Ys[0].max() < 10
    img = torch.ones(3, 32, 32)
    out = B.Cutout(16, np.random.RandomState(0))(img)
    assert 0 < (out == 0).sum() <= 3 * 16 * 16 and out.shape == img.shape
