"""Bandwidth sweep of the fused multi-GPU aggregation+broadcast kernel (BASELINE config 5 style): per-round time of
``PeerAggregator.aggregate`` for ResNet-18 / CNN / char-LSTM sized cluster models, device-timed with CUDA events,
max over ranks, plus an NCCL baseline (local K1 + all_reduce) for the same job.
  torchrun --nproc-per-node N tools/peer_agg_bench.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from feddrift_b200 import ops  # noqa: E402
from feddrift_b200.parallel.peer_aggregate import PeerAggregator  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
from bench import ClockSampler  # noqa: E402
clk = ClockSampler(int(os.environ.get("LOCAL_RANK", 0)))
clk.start()
rows_out = []
CONFIGS = [("resnet18_2clusters_32clients", 2, 11_699_132, 32), ("cnn_4clusters_64clients", 4, 1_199_882, 64),
           ("charlstm_2clusters_128clients", 2, 822_570, 128)]


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for i in range(iters):
        flush.fill_(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for name, M, P, C in CONFIGS:
    Cl = C // world
    P4 = (P + 3) // 4 * 4
    cp = torch.randn(Cl, M, P4, device=dev)
    n = torch.randint(1, 5, (Cl, M), device=dev).float()
    agg = PeerAggregator(M, P, dev)
    ms = timed(lambda: agg.aggregate(cp, n))
    agg.check()
    nvls, ms_p2p = bool(getattr(agg, "nvls", False)), None
    if nvls:   # same kernel with the multicast path disabled: W peer loads + W peer stores per element
        agg.mc_part, agg.mc_theta = 0, 0
        ms_p2p = timed(lambda: agg.aggregate(cp, n))
        agg.check()
    theta = torch.zeros(M, P4, device=dev)

    def nccl_path():
        nn = n.clone()
        if world > 1:
            tot = nn.sum(0)
            dist.all_reduce(tot)
        ops.cluster_aggregate_(theta, cp, nn)        # local normalised partial (weights sum to the local total)
        if world > 1:
            dist.all_reduce(theta)
    ms_nccl = timed(nccl_path)
    wire = (world - 1) / world * M * P4 * 4
    if rank == 0:
        rows_out.append({"config": name, "world": world, "fused_ms": ms, "nvls": nvls, "fused_p2p_ms": ms_p2p, "nccl_path_ms": ms_nccl,
                          "local_hbm_bytes": Cl * M * P4 * 4, "nvlink_bytes_each_way": wire,
                          "nvlink_GBps_each_way": wire / ms / 1e6 if world > 1 else None,
                          "frac_of_770GBps": (wire / ms / 1e6) / 770.0 if world > 1 else None})
    del cp, agg, theta
    torch.cuda.empty_cache()
clocks = clk.stop()
for r_ in rows_out:
    r_["clocks"] = clocks
    print(json.dumps(r_))
if world > 1:
    dist.destroy_process_group()
