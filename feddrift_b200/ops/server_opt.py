"""FedOpt server step fused into the K1 aggregation epilogue (native path)."""
from __future__ import annotations

from typing import Dict

import torch

from . import _ext

_KIND = {"sgd": 1, "adam": 2, "adagrad": 3, "yogi": 4}


def native_server_opt_step_(theta, avg, state: Dict, opt: str, lr: float, momentum=0.0, b1=0.9, b2=0.999, eps=1e-8):
    """``theta`` [P] or [M,P] is stepped in place using the pseudo-gradient ``theta - avg``.
    Implemented by running the K1 kernel on a single 'client' (= avg) with the optimizer epilogue."""
    ext = _ext.load(required=True)
    th = theta.reshape(1, -1) if theta.dim() == 1 else theta
    M, P = th.shape
    cp = avg.reshape(1, M, P).contiguous()
    n = torch.ones(1, M, dtype=torch.float32, device=th.device)
    state["step"] = state.get("step", 0) + 1
    s0 = s1 = None
    if opt == "sgd":
        if momentum:
            s0 = state.setdefault("momentum", torch.zeros_like(th))
    elif opt == "adagrad":
        s0 = state.setdefault("sum", torch.zeros_like(th))
    else:
        s0 = state.setdefault("m", torch.zeros_like(th))
        s1 = state.setdefault("v", torch.full_like(th, 1e-6) if opt == "yogi" else torch.zeros_like(th))
    ext.cluster_aggregate_opt(th, cp, n, _KIND[opt], float(lr), float(momentum), float(b1), float(b2), float(eps),
                              int(state["step"]), s0, s1)
    return theta
