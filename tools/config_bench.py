"""BASELINE.json configs 2-5 on the device engine (synthetic data of the named shapes, random-init weights): rounds/s
at time step 1 after a short warm-up.  Small MLPs take the fused kernel, everything else the generic executor.
Under torchrun (one process per GPU) the clients are sharded over the ranks: the fused kernel uses its NVLink peer-inbox
mode, the generic executor `sim.shard_clients` + `PeerAggregator`; times are the max over ranks."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from feddrift_b200.sim import DriftSim, make_args  # noqa: E402
from feddrift_b200.utils.metrics import MetricsSink  # noqa: E402

from feddrift_b200.experiments.configs import CONFIGS  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(dev))


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


which = sys.argv[1:] or list(CONFIGS)
for name in which:
    kw = dict(CONFIGS[name])
    kw.update(total_train_iteration=2, epochs=5, lr=0.01, report_client=0)
    try:
        t0 = time.perf_counter()
        sim = DriftSim(make_args(**kw), device=dev, sink=MetricsSink())
        if world > 1:
            if sim.spec is not None and sim.algo.fused_ok():
                from feddrift_b200.parallel.symm import attach_multi_gpu
                attach_multi_gpu(sim, world, rank)
            else:
                sim.shard_clients = True
        sim.run_time_step(0, rounds=1)
        sim.begin_time_step(1)
        sim.run_rounds(1)
        sync()
        setup = time.perf_counter() - t0
        R = kw["comm_round"]
        t1 = time.perf_counter()
        out = sim.run_rounds(R)
        sync()
        dt = time.perf_counter() - t1
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt)
        if rank == 0:
            print(json.dumps({"config": name, "world": world, "rounds": R, "rounds_per_s": R / dt, "s_per_round": dt / R, "setup_s": setup,
                              "fused_kernel": bool(sim.spec is not None and sim.algo.fused_ok()), "P": sim.bank.P,
                              "last": {k: round(v, 4) for k, v in out.items() if isinstance(v, float)},
                              "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
        del sim
        torch.cuda.empty_cache()
    except Exception as e:  # keep going: a failing config must not hide the others
        import traceback
        print(json.dumps({"config": name, "rank": rank, "error": repr(e)[:300], "trace": traceback.format_exc()[-600:]}), flush=True)
if world > 1:
    dist.destroy_process_group()
