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
# end-to-end round graph (fused host I/O on every rank) vs the single-GPU explicit-copy path
def mk(multi):
    s_ = DriftSim(make_args(comm_round=6, total_train_iteration=4), device=f"cuda:{rank}")
    if multi:
        attach_multi_gpu(s_, world, rank)
    for t_ in range(2):
        s_.run_time_step(t_, rounds=4)
    s_.begin_time_step(2)
    s_.args.rounds_per_launch = 1
    return s_
a, b = mk(True), mk(False)
ha, hb = a.make_host_round_inputs(), b.make_host_round_inputs()
for _ in range(3):
    ra, rb = a.run_round(ha, use_graph=True), b.run_round(hb, use_graph=False)
    ok = ok and abs(ra["train_acc"] - rb["train_acc"]) < 1e-5 and abs(ra["test_loss"] - rb["test_loss"]) < 1e-3
check_error(a)
ok = ok and (a.bank.theta - b.bank.theta).abs().max().item() < 1e-4
print(json.dumps({"rank": rank, "err": err, "ok": bool(ok), "acc": out["history"][-1]["train_acc"], "e2e": ra, "e2e_ref": rb}))
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_matches_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root, PYTHONFAULTHANDLER="1")
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
M, P, C = 3, 300004, 6   # > 1 chunk per model (chunks are 2^16 float4)
g = torch.Generator().manual_seed(0)
cp_all = torch.randn(C, M, P, generator=g)
n_all = torch.randint(0, 4, (C, M), generator=g).float()
n_all[:, 1] = 0   # an unused cluster must stay untouched
theta0 = torch.randn(M, P, generator=g)
mine = [c for c in range(C) if c % world == rank]
agg = PeerAggregator(M, P, f"cuda:{rank}", theta0.cuda())
ok = True
arena = cp_all.cuda()
cidx = torch.tensor(mine, dtype=torch.int32, device="cuda")
for it in range(3):
    if it == 1:   # in-place form: the whole client arena + the list of this rank's rows (what the engine uses)
        th = agg.aggregate(arena + it, n_all[mine].cuda(), cidx)
    else:
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
    env = dict(os.environ, FDB_ROOT=root, PYTHONFAULTHANDLER="1")
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
# fnn (no dropout): the sharded and the single-GPU run must consume identical randomness
kw = dict(model="fnn", dataset="MNIST", client_num_in_total=6, client_num_per_round=6, concept_drift_algo="win-1",
          concept_num=2, change_points="A", sample_num=16, batch_size=8, comm_round=2, total_train_iteration=2, epochs=2)
sim = DriftSim(make_args(**kw), device=f"cuda:{rank}", sink=MetricsSink())
sim.shard_clients = True
out = sim.run()
sim._peer_agg.check()
ref = DriftSim(make_args(**kw), device=f"cuda:{rank}", sink=MetricsSink())
oref = ref.run()
err = (sim.bank.theta - ref.bank.theta).abs().max().item()
ok = err < 5e-3 and abs(out["history"][-1]["train_loss"] - oref["history"][-1]["train_loss"]) < 5e-2
print(json.dumps({"rank": rank, "err": err, "ok": bool(ok), "loss": out["history"][-1]["train_loss"],
                  "loss_ref": oref["history"][-1]["train_loss"]}))
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_generic_executor_sharded_clients_with_peer_aggregation(tmp_path):
    script = tmp_path / "generic_worker.py"
    script.write_text(GENERIC_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root, PYTHONFAULTHANDLER="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29521", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


PULL_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDB_ROOT"])
from feddrift_b200.parallel.peer_linear import PeerWeights
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
g = torch.Generator().manual_seed(5)
W1, b1 = torch.randn(1568, 784, generator=g) * 0.05, torch.randn(1568, generator=g)
W2 = torch.randn(16, 1568, generator=g) * 0.05
store = PeerWeights({"fc1": (1568, 784), "fc2": (16, 1568)}, f"cuda:{rank}")
if rank == 0:   # only the owner ever holds the weights
    store.publish("fc1", W1.cuda())
    store.publish("fc2", W2.cuda())
torch.cuda.synchronize()
store.fence()
x = torch.randn(500, 784, generator=torch.Generator().manual_seed(100 + rank)).cuda()
h = store.linear(x, "fc1", owner=0, bias=b1.cuda(), relu=True)
y = store.linear(h, "fc2", owner=0)
torch.cuda.synchronize()
xb, w1b, w2b = x.cpu().bfloat16().float(), W1.bfloat16().float(), W2.bfloat16().float()
h_ref = torch.relu(xb @ w1b.t() + b1)
y_ref = h_ref.bfloat16().float() @ w2b.t()
e1 = (h.cpu() - h_ref).abs().max().item()
e2 = (y.cpu() - y_ref).abs().max().item()
ok = e1 < 2e-2 and e2 < 5e-2
print(json.dumps({"rank": rank, "e1": e1, "e2": e2, "ok": ok}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_pull_gemm_reads_weights_from_owner_gpu(tmp_path):
    """K2: the tcgen05 GEMM's TMA producer pulls the weight tiles from rank 0's symmetric arena over NVLink."""
    script = tmp_path / "pull_worker.py"
    script.write_text(PULL_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root, PYTHONFAULTHANDLER="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


GOSSIP_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDB_ROOT"])
from feddrift_b200.parallel.peer_gossip import PeerGossip
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
P = 100003
g = torch.Generator().manual_seed(11)
X0 = torch.randn(world, P, generator=g)
W = torch.rand(world, world, generator=g) + 0.1
W = W / W.sum(1, keepdim=True)                  # row-stochastic (DSGD)
node = PeerGossip(P, f"cuda:{rank}")
node.x.copy_(X0[rank])
torch.cuda.synchronize(); dist.barrier()
X = X0.clone()
for k in range(5):
    node.step(W[rank].tolist())
    X = W @ X
torch.cuda.synchronize()
node.check()
err = (node.x.cpu() - X[rank]).abs().max().item()
# PushSum: column-stochastic weights, omega mixes along; x/omega converges to the average
Wc = (torch.rand(world, world, generator=g) + 0.1)
Wc = Wc / Wc.sum(0, keepdim=True)
ps = PeerGossip(P, f"cuda:{rank}")
ps.x.copy_(X0[rank])
torch.cuda.synchronize(); dist.barrier()
for k in range(40):
    ps.pushsum_step(Wc[rank].tolist())
torch.cuda.synchronize()
ps.check()
err2 = (ps.debiased().cpu() - X0.mean(0)).abs().max().item()
ok = err < 1e-4 and err2 < 1e-3
print(json.dumps({"rank": rank, "err": err, "err_pushsum": err2, "ok": ok}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_gossip_dsgd_and_pushsum(tmp_path):
    """K12: in-kernel neighbour exchange over peer memory reproduces W^k·X (DSGD) and the PushSum average."""
    script = tmp_path / "gossip_worker.py"
    script.write_text(GOSSIP_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_ROOT=root, PYTHONFAULTHANDLER="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
