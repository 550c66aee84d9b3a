"""Compare the generic executor's execution modes on one GPU (eager / graphs with 1 stream / graphs with 8 streams)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402
from feddrift_b200.utils.metrics import MetricsSink  # noqa: E402

kw = dict(model="fnn", dataset="MNIST", client_num_in_total=6, client_num_per_round=6, concept_drift_algo="softcluster",
          concept_drift_algo_arg="H_A_C_1_10_0", concept_num=2, change_points="A", sample_num=16, batch_size=8, comm_round=2,
          total_train_iteration=2, epochs=3)


def run(env):
    for k in ("FDB_NO_GRAPHS", "FDB_GRAPH_STREAMS", "FDB_NO_PAIR_GRAPH"):
        os.environ.pop(k, None)
    os.environ.update(env)
    sim = DriftSim(make_args(**kw), device="cuda", sink=MetricsSink())
    out = sim.run()
    torch.cuda.synchronize()
    return sim.bank.theta.clone(), out["history"][-1]


ref, h0 = run({"FDB_NO_GRAPHS": "1"})
ref2, _ = run({"FDB_NO_GRAPHS": "1"})
print("eager vs eager", (ref - ref2).abs().max().item(), h0)
for name, env in (("pair-graph 1 stream", {"FDB_GRAPH_STREAMS": "1"}), ("pair-graph 8 streams", {}),
                  ("step-graph 1 stream", {"FDB_GRAPH_STREAMS": "1", "FDB_NO_PAIR_GRAPH": "1"}),
                  ("step-graph 8 streams", {"FDB_NO_PAIR_GRAPH": "1"})):
    th, h = run(env)
    print(name, (ref - th).abs().max().item(), (ref - th).abs().mean().item(), h)
