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


def test_round_watchdog_closes_rounds_without_a_crashed_worker():
    """Failure tolerance (the reference blocks forever on a missing upload): with ``round_timeout_s`` set, a worker whose
    upload is lost (fault injection) does not stall the experiment — the round closes with the 9 uploads that arrived,
    the straggler rejoins at the next broadcast, every round is still evaluated."""
    sink = set_sink(MetricsSink())
    a = _args("--concept_drift_algo", "softcluster", "--concept_drift_algo_arg", "H_A_C_1_10_0")
    a.round_timeout_s, a.min_workers_per_round, a.fault_drop = 5.0, 5, {1: [3], 2: [0, 7]}
    out = run_facade(a, sink)
    assert len(out["history"]) == 3 and len(sink.series("Train/Acc")) == 9
    assert a.watchdog_timeouts == 6            # rounds 1 and 2 of each of the 3 time steps
    # without the watchdog the same fault leaves the round open: the event loop drains and the experiment stops early
    sink2 = set_sink(MetricsSink())
    b = _args("--concept_drift_algo", "softcluster", "--concept_drift_algo_arg", "H_A_C_1_10_0")
    b.fault_drop = {1: [3]}
    run_facade(b, sink2)
    assert len(sink2.series("Train/Acc")) < 9


def test_round_watchdog_timer_posts_a_local_timeout_message():
    import time
    from feddrift_b200.core.managers import RoundWatchdog

    class FakeManager:
        backend, rank, posted = "DIST", 0, []

        def post_local(self, msg):
            self.posted.append(msg)
            return True

    m = FakeManager()
    wd = RoundWatchdog(m, timeout_s=0.05, min_workers=2)
    wd.arm(4)
    time.sleep(0.3)
    assert len(m.posted) == 1 and m.posted[0].get("round_idx") == 4 and m.posted[0].get_type() == RoundWatchdog.MSG_TYPE_ROUND_TIMEOUT
    wd.arm(5)
    wd.cancel()
    time.sleep(0.15)
    assert len(m.posted) == 1
