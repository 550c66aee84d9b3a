"""Where does a multi-GPU fused round spend its time?  Run under torchrun (N >= 1).  Prints, on rank 0:
  * in-kernel globaltimer stamps (train end / aggregation end / eval end) of per-launch rounds,
  * device-timed per-launch and persistent rounds,
  * end-to-end run_round() timing in three loop shapes: bench-style (flush + sync between rounds), back-to-back, and with a
    host barrier before every round (removes inter-process launch skew — diagnostic only)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from bench import HEADLINE, BENCH_TIME_STEP, ClockSampler  # noqa: E402
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402

world, rank, lr_ = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sim = DriftSim(make_args(**HEADLINE), device=dev)
if world > 1:
    from feddrift_b200.parallel.symm import attach_multi_gpu
    attach_multi_gpu(sim, world, rank)
for t in range(BENCH_TIME_STEP):
    sim.run_time_step(t, rounds=20)
sim.begin_time_step(BENCH_TIME_STEP)
sim.args.rounds_per_launch = 1
K = 200
clk = ClockSampler(lr_)
clk.start()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = {"world": world}
# in-kernel stamps
st = sim._small_state()
timers = torch.zeros(1, 4, dtype=torch.int64, device=dev)
st["timers"] = timers
rows = []
for i in range(30):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run_round_device(); e1.record()
    torch.cuda.synchronize()
    tm = timers[0].tolist()
    rows.append((e0.elapsed_time(e1) * 1e3, (tm[1] - tm[0]) / 1e3, (tm[2] - tm[1]) / 1e3))
rows = rows[10:]
out["launch_us_avg"] = sum(r[0] for r in rows) / len(rows)
out["agg_phase_us_avg"] = sum(r[1] for r in rows) / len(rows)     # train end -> aggregation (incl. cross-GPU exchange) end
out["eval_phase_us_avg"] = sum(r[2] for r in rows) / len(rows)
st.pop("timers")
barrier()
# e2e variants
host_inputs = sim.make_host_round_inputs()
for _ in range(10):
    sim.run_round(host_inputs, use_graph=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in ("bench_style", "back_to_back", "host_barrier"):
    barrier()
    tot = 0.0
    for i in range(K):
        if name == "bench_style":
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
        elif name == "host_barrier" and world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim.run_round(host_inputs, use_graph=True)
        tot += time.perf_counter() - t0
    tt = torch.tensor([tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out[f"e2e_{name}_us"] = float(tt) / K * 1e6
out["clocks"] = clk.stop()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
