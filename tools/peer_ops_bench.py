"""Device-timed (CUDA events, max over ranks) micro-benchmarks of the NVLink peer-memory kernels that are not collectives
in the NCCL sense:
  * K2  peer-pull GEMM: Y = relu(X·Wᵀ + b) with W resident on the NEXT rank (TMA loads over NVLink inside the GEMM) vs W local;
  * K12 peer gossip: one ring mixing step x_i ← w_l·x_{i-1} + w_s·x_i + w_r·x_{i+1} for a ResNet-18 sized vector.
  torchrun --nproc-per-node N tools/peer_ops_bench.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from feddrift_b200.ops import _ext  # noqa: E402
from feddrift_b200.parallel.peer_gossip import PeerGossip  # noqa: E402
from feddrift_b200.parallel.peer_linear import PeerWeights  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
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


ext = _ext.load()
for name, M, N, K in (("fnn_mnist_fc1_batch500", 500, 1568, 784), ("square_4096", 4096, 4096, 4096)):
    store = PeerWeights({"w": (N, K)}, dev)
    store.publish("w", torch.randn(N, K, device=dev) * 0.02)
    torch.cuda.synchronize()
    store.fence()
    x = torch.randn(M, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    owner = (rank + 1) % world
    t_local = timed(lambda: ext.gemm_tn_bias_act_peer(x, store.ptr("w", rank), N, bias, True, False))
    t_peer = timed(lambda: ext.gemm_tn_bias_act_peer(x, store.ptr("w", owner), N, bias, True, False))
    if rank == 0:
        print(json.dumps({"op": "peer_pull_gemm", "shape": name, "world": world, "local_us": t_local * 1e3, "peer_us": t_peer * 1e3,
                          "weight_bytes": N * K * 2, "nvlink_GBps": (N * K * 2) / (t_peer * 1e-3) / 1e9 if world > 1 else None}), flush=True)
    if world > 1:
        dist.barrier()
    del store

P = 11_699_132
node = PeerGossip(P, dev)
node.x.normal_()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
w = [0.0] * world
w[rank] = 0.5 if world > 1 else 1.0
if world > 1:
    w[(rank - 1) % world] += 0.25
    w[(rank + 1) % world] += 0.25
t = timed(lambda: node.step(w))
node.check()
nbrs = sum(1 for i, v in enumerate(w) if v != 0 and i != rank)
if rank == 0:
    print(json.dumps({"op": "peer_gossip_step", "P": P, "world": world, "neighbours": nbrs, "ms": t,
                      "nvlink_in_bytes": nbrs * P * 4, "nvlink_in_GBps": nbrs * P * 4 / (t * 1e-3) / 1e9 if nbrs else None,
                      "local_GBps": (1 + 1) * P * 4 / (t * 1e-3) / 1e9}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
