"""Generic (any-model) round executor of the device engine.

Same plan semantics as the fused ``fed_round_small`` kernel (W-pool / time-weighted / index-list batch selection with
the same counter-based RNG, per-pair optimizer state that persists across rounds, per-cluster weighted aggregation,
train/test evaluation with optional ensembles) but built from separate native kernels so it works for every
architecture and for algorithms that must look at raw client updates before aggregating (CFL):

* local step: bank-bound ``nn.Module`` forward/backward (TcLinear → tcgen05 GEMM; convs/LSTMs → library kernels),
  gradients land in a flat scratch row, ``ops.adam_amsgrad_rows_`` / ``ops.sgd_rows_`` update the client row;
* aggregation: ``ops.cluster_aggregate_`` over the ``[C, M, P]`` client arena (K1);
* evaluation: ``ops.eval_logits`` device-side accumulation, ONE host copy per block of rounds.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from ..ops.reference import batch_hash, mix32


def _pair_sampler(st: Dict, c: int, m: int, t: int, nb: torch.Tensor, B: int):
    """(n_cm, sampler(h1) -> index tensor into the flattened [T1*S] sample axis of client c)."""
    from ..ops.reference import _pair_plan
    return _pair_plan(st, c, m, t, nb, B)


def run_rounds_generic(sim, rounds: int) -> Dict[str, torch.Tensor]:
    a, bank, data = sim.args, sim.bank, sim.data
    dev = sim.device
    plan = sim.current_plan()
    t, C, M, P = sim.t, sim.C, sim.M, bank.P
    T1, S = data.steps, data.X.shape[2]
    B, E = a.batch_size, a.epochs
    W = plan["W"]
    nsamp_h = sim.data_host.nsamp
    nb = (nsamp_h.to(torch.int64) + B - 1) // B
    st = {"W": W.cpu(), "nsamp": nsamp_h, "X": sim.data_host.X, "sample_mode": plan.get("sample_mode", "pool"),
          "n_mode": plan.get("n_mode", "batches"), "train_index": _cpu(plan.get("train_index")),
          "train_count": _cpu(plan.get("train_count"))}
    use_adam = plan.get("optimizer", "sgd" if a.client_optimizer == "sgd" else "adam") != "sgd"
    lr = float(plan.get("lr", a.lr))
    seed = int(a.dummy_arg) * 7919 + 13 + 1000003 * t
    feat_mask = plan.get("feat_mask")
    ens_mode = int(plan.get("ens_mode", 0) or 0)
    metrics = torch.zeros(rounds, C, 4, dtype=torch.float32, device=dev)
    cl = sim.clients
    Xc_all = data.X.reshape(T1, C, S, *data.X.shape[3:])
    for r in range(rounds):
        rnd = sim.round_in_step + r
        Wt = st["W"][t]
        active = (st["train_count"] > 0).any(dim=1) if st["sample_mode"] == "index" else (Wt != 0).any(dim=1)
        cl.n.zero_()
        world, rank = _world_rank(sim)
        for c in range(C):
            if world > 1 and c % world != rank:   # clients are sharded over the ranks (one process per GPU)
                continue
            Xc = Xc_all[:, c].reshape(T1 * S, *data.X.shape[3:])
            Yc = data.Y[:, c].reshape(T1 * S)
            for m in range(M):
                if not bool(active[m]):
                    continue
                n_cm, sampler = _pair_sampler(st, c, m, t, nb, B)
                if n_cm <= 0:
                    continue
                cl.params[c, m].copy_(bank.theta[m])
                _local_steps(sim, c, m, Xc, Yc, sampler, seed, rnd, E, use_adam, lr, a.wd, feat_mask)
                cl.n[c, m] = n_cm
        # raw-update hooks (CFL family) may veto the aggregation of this round
        skip = False
        if hasattr(sim.algo, "state") and "cfl" in getattr(sim.algo, "arg", ""):
            skip = sim.algo.state.cluster_cfl(t, rnd + 1, bank, cl.params, cl.n)
            if skip:
                plan["W"] = sim.algo.state.weights_tensor(t)
                st["W"] = plan["W"].cpu()
        if hasattr(sim.algo, "on_client_updates") and not sim.algo.split_done and rnd == sim.algo.split_round:
            sim.algo.on_client_updates(t, cl.params, cl.n)
        if not skip:
            if world > 1:
                _peer_aggregate(sim, world, rank)
            else:
                ops.cluster_aggregate_(bank.theta, cl.params, cl.n)
        if plan.get("recluster_hard"):
            acc = sim.evaluator.acc_matrix(list(range(M)), t)
            best = np.argmax(acc, axis=0)
            st["W"][t].zero_()
            st["W"][t][torch.from_numpy(best), torch.arange(C)] = 1.0
            plan["W"] = st["W"].clone()
            sim.algo.absorb_weights(t, st["W"])
        _evaluate(sim, plan, st, r, metrics, ens_mode)
    if _world_rank(sim)[0] > 1:   # cold path: every rank evaluated only its own clients
        import torch.distributed as dist
        dist.all_reduce(metrics)
    counts = torch.stack([data.nsamp[t], data.nsamp[t + 1] if t + 1 < T1 else torch.zeros_like(data.nsamp[t])], 1).float()
    return {"metrics": metrics, "counts": counts}


def _world_rank(sim):
    import torch.distributed as dist
    if getattr(sim, "shard_clients", False) and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _peer_aggregate(sim, world, rank):
    """Multi-GPU aggregation + broadcast of the cluster models in ONE kernel over NVLink peer memory
    (``parallel/peer_aggregate.py``): this rank contributes the rows of its own clients."""
    from ..parallel.peer_aggregate import PeerAggregator
    pa = getattr(sim, "_peer_agg", None)
    if pa is None:
        pa = sim._peer_agg = PeerAggregator(sim.M, sim.bank.P, sim.device, sim.bank.theta)
    mine = [c for c in range(sim.C) if c % world == rank]
    theta = pa.aggregate(sim.clients.params[mine], sim.clients.n[mine])
    sim.bank.theta.copy_(theta)


def _cpu(x: Optional[torch.Tensor]):
    return x.cpu() if isinstance(x, torch.Tensor) else x


def _local_steps(sim, c, m, Xc, Yc, sampler, seed, rnd, E, use_adam, lr, wd, feat_mask):
    bank, cl = sim.bank, sim.clients
    row = cl.params[c, m]
    mlp = bank.mlp
    if mlp is None:
        mod = _scratch_module(sim)
        _bind(mod, bank, row)
        mod.train()
    for step in range(E):
        h1 = batch_hash(seed, rnd, c, m, step)
        idx = sampler(h1, mix32(h1 ^ 0x68E31DA4)).to(Xc.device)
        xb, yb = Xc[idx], Yc[idx].long()
        if feat_mask is not None:
            xb = xb * feat_mask[m].reshape((1,) + tuple(xb.shape[1:]))
        if mlp is not None:
            th = row.detach().clone().requires_grad_(True)
            logits = ops.mlp_forward(th, xb.reshape(xb.shape[0], -1), mlp["kind"], mlp["in"], mlp["hidden"], mlp["out"])
            (g,) = torch.autograd.grad(F.cross_entropy(logits, yb), th)
        else:
            for p_ in mod.parameters():
                p_.grad = None
            F.cross_entropy(mod(xb), yb).backward()
            g = _flat_grads(mod, bank, row)
        r2 = row.reshape(1, -1)
        if use_adam:
            ops.adam_amsgrad_rows_(r2, g.reshape(1, -1), cl.m[c, m].reshape(1, -1), cl.v[c, m].reshape(1, -1),
                                   cl.vmax[c, m].reshape(1, -1), cl.step[c, m].reshape(1), lr, wd)
        else:
            ops.sgd_rows_(r2, g.reshape(1, -1), lr, 0.0)


def _scratch_module(sim):
    mod = getattr(sim, "_scratch_mod", None)
    if mod is None:
        import copy
        mod = sim._scratch_mod = copy.deepcopy(sim.bank.template).to(sim.device)
    return mod


def _bind(mod, bank, row):
    """Point the module's parameters/buffers at ``row`` (views, no copy)."""
    from ..parallel.arena import _set_tensor
    from ..models.utils import unflatten_to_state_dict
    views = unflatten_to_state_dict(row, bank.spec)
    for name, p in list(mod.named_parameters()):
        _set_tensor(mod, name, torch.nn.Parameter(views[name], requires_grad=True))
    for name, b in list(mod.named_buffers()):
        if name in views and views[name].dtype == b.dtype:
            _set_tensor(mod, name, views[name], buffer=True)


def _flat_grads(mod, bank, row):
    g = torch.zeros_like(row)
    grads = dict(mod.named_parameters())
    for k, shape, dtype, off, n in bank.spec:
        p = grads.get(k)
        if p is not None and p.grad is not None:
            g[off:off + n] = p.grad.reshape(-1)
    return g


def _evaluate(sim, plan, st, r, metrics, ens_mode):
    data, bank, t, C = sim.data, sim.bank, sim.t, sim.C
    pick = st["W"][t].argmax(dim=0)
    etr, ete = plan.get("eval_train_model"), plan.get("eval_test_model")
    acc = torch.zeros(3, dtype=torch.float32, device=sim.device)
    world, rank = _world_rank(sim)
    with torch.no_grad():
        for c in range(C):
            if world > 1 and c % world != rank:
                continue
            mtr = int(etr[c]) if etr is not None and int(etr[c]) >= 0 else int(pick[c])
            mte = int(ete[c]) if ete is not None and int(ete[c]) >= 0 else int(pick[c])
            n0 = int(sim.data_host.nsamp[t, c])
            if n0:
                acc.zero_()
                ops.eval_logits(bank.forward(mtr, data.X[t, c, :n0]), data.Y[t, c, :n0], acc)
                metrics[r, c, 0:2] = acc[0:2]
            if t + 1 < data.steps:
                n1 = int(sim.data_host.nsamp[t + 1, c])
                if n1 == 0:
                    continue
                x1, y1 = data.X[t + 1, c, :n1], data.Y[t + 1, c, :n1]
                if ens_mode == 0:
                    acc.zero_()
                    ops.eval_logits(bank.forward(mte, x1), y1, acc)
                    metrics[r, c, 2:4] = acc[0:2]
                else:
                    w = plan["ens_w"][c]
                    ks = [k for k in range(bank.num_models) if float(w[k]) > 0]
                    if ens_mode == 1:
                        preds = torch.stack([bank.forward(k, x1).argmax(-1) for k in ks])
                        vote = ops.ensemble_vote(preds, w[ks].to(sim.device), data.class_num)
                    else:
                        probs = torch.stack([torch.softmax(bank.forward(k, x1), 1) for k in ks])
                        vote = ops.soft_vote(probs, w[ks].to(sim.device))
                    metrics[r, c, 2] = (vote == y1).sum().float()
