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

FORMAT_VERSION = 1


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


def load(path: str) -> Dict:
    return torch.load(path, map_location="cpu", weights_only=False)


def resume(sim, path: str) -> int:
    """Restore ``sim`` from a checkpoint; returns the next time step to run."""
    blob = load(path)
    if blob["version"] != FORMAT_VERSION:
        raise ValueError(f"checkpoint version {blob['version']} != {FORMAT_VERSION}")
    sim.bank.theta.copy_(blob["theta"].to(sim.bank.device))
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
    params = torch.load(path, map_location="cpu", weights_only=False)
    for m, sd in params.items():
        sim.bank.load_state_dict(int(m), sd)
