import os

import torch

from feddrift_b200.sim import DriftSim, checkpoint, make_args
from feddrift_b200.utils.metrics import MetricsSink


def test_checkpoint_resume_reproduces_uninterrupted_run(tmp_path):
    kw = dict(dataset="sine", concept_drift_algo_arg="H_A_C_1_0_0", comm_round=8, lr=0.05, total_train_iteration=4,
              sample_num=60)
    full = DriftSim(make_args(**kw), device="cpu", sink=MetricsSink())
    full.run()
    part = DriftSim(make_args(checkpoint_dir=str(tmp_path), **kw), device="cpu", sink=MetricsSink())
    part.run(0, 2)
    assert sorted(os.listdir(tmp_path)) == ["step_0000.fdck", "step_0001.fdck"]
    resumed = DriftSim(make_args(checkpoint_dir=str(tmp_path), **kw), device="cpu", sink=MetricsSink())
    nxt = checkpoint.resume(resumed, checkpoint.latest(str(tmp_path)))
    assert nxt == 2
    resumed.run(nxt)
    assert torch.allclose(resumed.bank.theta, full.bank.theta, atol=1e-6)
    assert (resumed.algo.state.W[:4] == full.algo.state.W[:4]).all()
    p = checkpoint.export_model_params(resumed, str(tmp_path / "model_params.pt"))
    blob = torch.load(p, weights_only=False)
    assert set(blob.keys()) == {0, 1, 2, 3} and "fc1.weight" in blob[0]
