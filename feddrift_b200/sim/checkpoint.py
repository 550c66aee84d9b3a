"""Per-time-step checkpoint / resume.

The reference carries state across time steps through six files in CWD (``model_params.pt`` written only at
the final round, ``sc_state.pkl`` & friends pickled during the last 4 rounds, two of them containing whole
``nn.Module``s — SURVEY §5).  Here ONE versioned blob per time step holds: the parameter arena, the drift
state (plain arrays/scalars, never modules), RNG states and the metric history, so a run can resume at any
time-step boundary in-process (`resume`), and ``export_model_params`` writes the reference's
``{m: state_dict}`` format for interoperability.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

FORMAT_VERSION = 2      # 2: tensor-core conv weights are stored (O, kh, kw, I) in the rows (models.utils.ohwi_stored); 1: logical order


def upgrade_theta_v1(theta: torch.Tensor, spec) -> torch.Tensor:
    """Rows written by format 1 (every tensor in logical order) → the current row layout."""
    from ..models.utils import ohwi_stored
    out = theta.clone()
    for _, shape, _, off, n in spec:
        if ohwi_stored(shape):
            seg = theta[:, off:off + n].reshape(theta.shape[0], *shape)
            out[:, off:off + n] = seg.permute(0, 1, 3, 4, 2).reshape(theta.shape[0], -1)
    return out


def save(sim, path: str) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    blob = {
        "version": FORMAT_VERSION,
        "t": sim.t,
        "global_round": sim.global_round,
        "args": {k: v for k, v in vars(sim.args).items() if isinstance(v, (int, float, str, bool, type(None)))},
        "theta": sim.bank.theta.detach().cpu().clone(),
        "algo": sim.algo.state_dict(),
        "history": list(sim.history),
        "summary": dict(sim.sink.run.summary),
        "np_rng": sim.rng.get_state(),
        "torch_rng": torch.get_rng_state(),
    }
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, path)  # atomic: a crash never leaves a half-written checkpoint
    return path


def _safe_globals():
    """Checkpoints hold tensors, numpy arrays and plain containers only — load them with the restricted unpickler
    (``weights_only=True``) plus an allow-list of the numpy reconstruction helpers, never arbitrary objects."""
    import numpy as np
    allow = [np.ndarray, np.dtype, np.random.RandomState]
    for mod, name in (("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
                      ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")):
        try:
            allow.append(getattr(__import__(mod, fromlist=[name]), name))
        except Exception:  # noqa: BLE001 — module layout differs between numpy 1.x / 2.x
            pass
    for t in ("bool_", "int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "float16", "float32", "float64"):
        allow.append(getattr(np, t))
    try:
        allow += [type(np.dtype(t)) for t in ("bool", "int32", "int64", "uint32", "float32", "float64")]
    except Exception:  # noqa: BLE001
        pass
    return allow


def load(path: str) -> Dict:
    try:
        with torch.serialization.safe_globals(_safe_globals()):
            return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:  # noqa: BLE001
        if os.environ.get("FDB_TRUSTED_CHECKPOINTS") == "1":   # legacy blobs with arbitrary pickled objects: opt-in only
            return torch.load(path, map_location="cpu", weights_only=False)
        raise RuntimeError(f"checkpoint {path} is not loadable with the restricted unpickler ({exc}); set "
                           "FDB_TRUSTED_CHECKPOINTS=1 to unpickle a checkpoint you trust") from exc


def resume(sim, path: str) -> int:
    """Restore ``sim`` from a checkpoint; returns the next time step to run."""
    blob = load(path)
    if blob["version"] not in (1, FORMAT_VERSION):
        raise ValueError(f"checkpoint version {blob['version']} != {FORMAT_VERSION}")
    theta = blob["theta"] if blob["version"] == FORMAT_VERSION else upgrade_theta_v1(blob["theta"], sim.bank.spec)
    sim.bank.theta.copy_(theta.to(sim.bank.device))
    sim.algo.load_state_dict(blob["algo"])
    sim.history = list(blob["history"])
    sim.global_round = blob["global_round"]
    sim.rng.set_state(blob["np_rng"])
    torch.set_rng_state(blob["torch_rng"])
    for k, v in blob["summary"].items():
        sim.sink.set_summary(k, v)
    sim.t = blob["t"]
    return blob["t"] + 1


def latest(cdir: str):
    files = sorted(f for f in os.listdir(cdir) if f.endswith(".fdck")) if os.path.isdir(cdir) else []
    return os.path.join(cdir, files[-1]) if files else None


def export_model_params(sim, path: str = "model_params.pt") -> str:
    """Reference-compatible ``torch.save({m: state_dict})`` (``FedAvgEnsServerManager.py:84-86``)."""
    torch.save({m: {k: v.detach().cpu().clone() for k, v in sim.bank.state_dict(m).items()}
                for m in range(sim.bank.num_models)}, path)
    return path


def import_model_params(sim, path: str = "model_params.pt") -> None:
    params = torch.load(path, map_location="cpu", weights_only=True)   # {m: state_dict} of plain tensors
    for m, sd in params.items():
        sim.bank.load_state_dict(int(m), sd)
