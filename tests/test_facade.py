"""FedML-compatible message-passing façade: INPROC event loop, zero-copy DeviceRef, gloo world_size=2 plumbing
(BASELINE.json config 1: SEA-4 fnn, 10 clients, softcluster H_A_F FedDrift on CPU/gloo)."""
import json
import os
import subprocess
import sys

import pytest
import torch  # noqa: F401  (imported for its side effects before the facade modules)

from feddrift_b200.experiments.fedavg_cont_ens import add_args, run_facade
from feddrift_b200.utils.metrics import MetricsSink, set_sink

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(*extra):
    import argparse
    return add_args(argparse.ArgumentParser()).parse_args(["--engine", "facade", "--comm_round", "3", "--total_train_iteration",
                                                            "3", "--sample_num", "60", *extra])


@pytest.mark.parametrize("algo,arg", [("softcluster", "H_A_C_1_10_0"), ("softcluster", "hard-r"), ("aue", ""), ("kue", ""),
                                      ("driftsurf", "0"), ("mmacc", ""), ("ada", "win-1_round"), ("exp", ""),
                                      ("softcluster", "cfl_0.1_win-1")])
def test_facade_inproc_runs(algo, arg):
    sink = set_sink(MetricsSink())
    extra = ["--concept_drift_algo", algo]
    if arg:
        extra += ["--concept_drift_algo_arg", arg]
    out = run_facade(_args(*extra), sink)
    assert len(out["history"]) == 3 and 0 <= out["history"][-1]["test_acc"] <= 1
    assert len(sink.series("Train/Acc")) == 9 and len(sink.series("Test/Acc-CL-0")) == 9


def test_facade_zero_copy_device_ref_equals_state_dict_transport():
    a1, a2 = _args(), _args()
    a2.zero_copy = 1
    o1 = run_facade(a1, set_sink(MetricsSink()))
    o2 = run_facade(a2, set_sink(MetricsSink()))
    assert o1["history"][-1]["train_acc"] == o2["history"][-1]["train_acc"]


def test_gloo_world2_plumbing_h_a_f():
    """Two OS processes over gloo: rank 0 = server, rank 1 hosts all 10 logical workers (packed)."""
    code = r'''
import os, sys, json
sys.path.insert(0, %r)
from feddrift_b200.experiments.fedavg_cont_ens import add_args, run_facade
from feddrift_b200.utils.metrics import MetricsSink, set_sink
import argparse
a = add_args(argparse.ArgumentParser()).parse_args(["--engine", "facade", "--backend", "GLOO", "--comm_round", "2",
    "--total_train_iteration", "2", "--sample_num", "40", "--concept_drift_algo_arg", "H_A_F_1_10_0", "--concept_num", "10"])
a.pack_workers = 1
out = run_facade(a, set_sink(MetricsSink()))
if int(os.environ["RANK"]) == 0:
    print("RESULT " + json.dumps(out["history"][-1]))
''' % ROOT
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", "-c", code]
    # torchrun has no -c; write the worker to a temp file instead
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as fh:
        fh.write(code)
        path = fh.name
    cmd = cmd[:-2] + [path]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    os.unlink(path)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    h = json.loads(line[len("RESULT "):])
    assert h["iteration"] == 1 and 0 <= h["test_acc"] <= 1
