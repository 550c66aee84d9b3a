"""Phase timing of the batched LSTM executor (config 5): CUDA-event time of every stage of one round, plus host time."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from feddrift_b200.experiments.configs import CONFIGS  # noqa: E402
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402
from feddrift_b200.utils.metrics import MetricsSink  # noqa: E402
from feddrift_b200.ops import lstm as L  # noqa: E402
from feddrift_b200.sim import lstm_exec, generic  # noqa: E402

kw = dict(CONFIGS["cfg5_shakespeare_lstm_128clients_win1"])
kw.update(total_train_iteration=2, epochs=5, lr=0.01, report_client=0)
sim = DriftSim(make_args(**kw), device="cuda:0", sink=MetricsSink())
sim.run_time_step(0, rounds=1)
sim.begin_time_step(1)
sim.run_rounds(1)
torch.cuda.synchronize()

stamps = {}


def timed(name, fn):
    def wrap(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        r = fn(*a, **k)
        e1.record()
        stamps.setdefault(name, []).append((e0, e1, time.perf_counter() - t0))
        return r
    return wrap


L.lstm2_pairs_forward = timed("lstm_fwd", L.lstm2_pairs_forward)
L.lstm2_pairs_backward = timed("lstm_bwd", L.lstm2_pairs_backward)
L.lstm_head = timed("head", L.lstm_head)
L.lstm2_weight_grads_per_chunk = timed("dW_batched_gemm", L.lstm2_weight_grads_per_chunk)
lstm_exec.train_pairs = timed("train_pairs_total", lstm_exec.train_pairs)
generic._evaluate = timed("evaluate", generic._evaluate)
import feddrift_b200.ops as ops  # noqa: E402
ops.adam_amsgrad_rows_ = timed("adam_rows", ops.adam_amsgrad_rows_)
ops.cluster_aggregate_ = timed("aggregate", ops.cluster_aggregate_)

t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
R = 3
sim.run_rounds(R)
e1.record()
torch.cuda.synchronize()
out = {"rounds": R, "wall_ms_per_round": (time.perf_counter() - t0) * 1e3 / R, "gpu_ms_per_round": e0.elapsed_time(e1) / R}
for k, v in stamps.items():
    out[k] = {"calls_per_round": len(v) / R, "gpu_ms_per_round": sum(a.elapsed_time(b) for a, b, _ in v) / R,
              "host_ms_per_round": sum(h for _, _, h in v) * 1e3 / R}
print(json.dumps(out))
