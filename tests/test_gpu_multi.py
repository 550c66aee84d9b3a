"""Multi-GPU: the fused round kernel with NVLink peer-inbox aggregation must reproduce the single-GPU result."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDB_ROOT"])
from feddrift_b200.sim import DriftSim, make_args
from feddrift_b200.parallel.symm import attach_multi_gpu, check_error
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
args = make_args(comm_round=6, total_train_iteration=3)
sim = DriftSim(args, device=f"cuda:{rank}")
attach_multi_gpu(sim, world, rank)
out = sim.run()
check_error(sim)
ref = DriftSim(make_args(comm_round=6, total_train_iteration=3), device=f"cuda:{rank}")
oref = ref.run()
err = (sim.bank.theta - ref.bank.theta).abs().max().item()
ok = err < 1e-4 and abs(out["history"][-1]["train_acc"] - oref["history"][-1]["train_acc"]) < 0.02
print(json.dumps({"rank": rank, "err": err, "ok": bool(ok), "acc": out["history"][-1]["train_acc"]}))
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_matches_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


PEER_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDB_ROOT"])
from feddrift_b200.parallel.peer_aggregate import PeerAggregator
from feddrift_b200.ops import reference as ref
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
M, P, C = 3, 100003, 6
g = torch.Generator().manual_seed(0)
cp_all = torch.randn(C, M, P, generator=g)
n_all = torch.randint(0, 4, (C, M), generator=g).float()
n_all[:, 1] = 0   # an unused cluster must stay untouched
theta0 = torch.randn(M, P, generator=g)
mine = [c for c in range(C) if c % world == rank]
agg = PeerAggregator(M, P, f"cuda:{rank}", theta0.cuda())
ok = True
for it in range(3):
    th = agg.aggregate(cp_all[mine].cuda() + it, n_all[mine].cuda())
    torch.cuda.synchronize()
    agg.check()
    want = theta0.clone()
    ref.cluster_aggregate_(want, cp_all + it, n_all)
    err = (th.cpu() - want).abs().max().item()
    ok = ok and err < 1e-4
    theta0 = want
print(json.dumps({"rank": rank, "ok": bool(ok), "err": err}))
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_reduce_apply_broadcast_matches_reference(tmp_path):
    script = tmp_path / "peer_worker.py"
    script.write_text(PEER_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


GENERIC_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDB_ROOT"])
from feddrift_b200.sim import DriftSim, make_args
from feddrift_b200.utils.metrics import MetricsSink
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
kw = dict(model="cnn", dataset="MNIST", client_num_in_total=6, client_num_per_round=6, concept_drift_algo="win-1",
          concept_num=2, change_points="A", sample_num=16, batch_size=8, comm_round=2, total_train_iteration=2, epochs=2)
sim = DriftSim(make_args(**kw), device=f"cuda:{rank}", sink=MetricsSink())
sim.shard_clients = True
out = sim.run()
sim._peer_agg.check()
ref = DriftSim(make_args(**kw), device=f"cuda:{rank}", sink=MetricsSink())
oref = ref.run()
err = (sim.bank.theta - ref.bank.theta).abs().max().item()
ok = err < 5e-3 and abs(out["history"][-1]["train_loss"] - oref["history"][-1]["train_loss"]) < 5e-2
print(json.dumps({"rank": rank, "err": err, "ok": bool(ok)}))
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_generic_executor_sharded_clients_with_peer_aggregation(tmp_path):
    script = tmp_path / "generic_worker.py"
    script.write_text(GENERIC_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29521", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
