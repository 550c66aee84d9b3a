"""cfg4 (ResNet-18, 32 clients) through the pair-stacked executor: rounds/s of steady-state rounds with the stacked step replayed
as a CUDA graph vs eager (FDB_STACKED_GRAPH from the environment); the first three rounds (warm-up, capture) are not timed."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from feddrift_b200.experiments.configs import CONFIGS  # noqa: E402
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402
from feddrift_b200.utils.metrics import MetricsSink  # noqa: E402

kw = dict(CONFIGS["cfg4_cifar_resnet18_32clients_aue"])
kw.update(total_train_iteration=2, epochs=5, lr=0.01, report_client=0)
sim = DriftSim(make_args(**kw), device="cuda", sink=MetricsSink())
sim.run_time_step(0, rounds=1)
sim.begin_time_step(1)
sim.run_rounds(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
R = 5
sim.run_rounds(R)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
st = sim.__dict__.get("_stack_stage")
print(json.dumps({"graph_env": os.environ.get("FDB_STACKED_GRAPH", "default"), "graphed": bool(st is not None and st.__dict__.get("graph")),
                  "broken": bool(sim.__dict__.get("_stack_graph_broken", False)), "rounds_per_s": R / dt, "ms_per_round": dt / R * 1e3}))
