"""torch.profiler view of one generic-executor round (top CUDA kernels + wall clock) for a BASELINE config."""
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402
from feddrift_b200.utils.metrics import MetricsSink  # noqa: E402

sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["cfg5_shakespeare_lstm_128clients_win1"])
exec(open("tools/config_bench.py").read().split("rank, world =")[0])   # CONFIGS
name = sys.argv[1]
kw = dict(CONFIGS[name])
kw.update(total_train_iteration=2, epochs=5, lr=0.01, report_client=0)
sim = DriftSim(make_args(**kw), device="cuda", sink=MetricsSink())
sim.run_time_step(0, rounds=1)
sim.begin_time_step(1)
sim.run_rounds(2)
torch.cuda.synchronize()
t0 = time.perf_counter()
sim.run_rounds(1)
torch.cuda.synchronize()
print("round wall s", time.perf_counter() - t0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sim.run_rounds(1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
