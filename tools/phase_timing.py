"""In-kernel phase timing of the fused round kernel (globaltimer stamps of cluster rank 0): local-training end,
aggregation end, evaluation end per round, in persistent mode (R rounds in one launch)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402
from bench import HEADLINE, BENCH_TIME_STEP  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sim = DriftSim(make_args(**HEADLINE), device="cuda")
for t in range(BENCH_TIME_STEP):
    sim.run_time_step(t, rounds=20)
sim.begin_time_step(BENCH_TIME_STEP)
st = sim._small_state()
sim.run_rounds_device(50)
timers = torch.zeros(R, 4, dtype=torch.int64, device="cuda")
st["timers"] = timers
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sim.run_rounds_device(R)
e1.record()
torch.cuda.synchronize()
tm = timers.cpu().double()
train = (tm[1:, 0] - tm[:-1, 2]).mean().item()
agg = (tm[:, 1] - tm[:, 0]).mean().item()
ev = (tm[:, 2] - tm[:, 1]).mean().item()
tot = (tm[-1, 2] - tm[0, 2]).item() / (R - 1)
print(json.dumps({"rounds": R, "launch_info": st.get("_launch_info"), "event_us_per_round": 1e3 * e0.elapsed_time(e1) / R,
                  "in_kernel_ns_per_round": tot, "train_ns": train, "aggregate_ns": agg, "eval_ns": ev,
                  "npairs_hint": int((st["W"][: st["t_cur"] + 1].sum(0) > 0).sum())}))
