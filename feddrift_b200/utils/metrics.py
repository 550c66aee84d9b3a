"""Metrics sink with the wandb call surface the reference uses (``wandb.log({...})``,
``wandb.run.summary[...] = ...``, ``wandb.init(project, name, config)``) — SURVEY §5.

Default is an in-memory + JSONL sink (offline box).  ``MetricsSink(use_wandb=True)`` forwards to the
real ``wandb`` package (offline mode) when importable.  Keys are identical to the reference so plots
line up: ``Train/Acc``, ``Train/Loss``, ``Test/Acc``, ``Test/Loss``, ``Train/Acc-CL-c``,
``Test/Acc-CL-c``, ``Plurality/CL-c``, ``Weight-All/CL-c``, summaries ``num_models``, ``local_models``,
``Contribute/CL-c``, ``Merge``, ``Reset-m``; CFL ``Max_Norm``/``Mean_Norm``.

The device engine accumulates per-round metrics ON DEVICE and flushes them here once per time step
(``log_rounds``), instead of 2N+4 synchronous ``wandb.log`` calls per round.
"""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, List, Optional


class _Run:
    def __init__(self) -> None:
        self.summary: Dict[str, Any] = {}
        self.name = ""
        self.config: Dict[str, Any] = {}


class MetricsSink:
    def __init__(self, path: Optional[str] = None, use_wandb: bool = False, keep: bool = True):
        self.run = _Run()
        self.records: List[Dict[str, Any]] = []
        self.path = path
        self.keep = keep
        self._fh = open(path, "a") if path else None
        self._wandb = None
        if use_wandb:
            try:
                os.environ.setdefault("WANDB_MODE", "offline")
                import wandb
                self._wandb = wandb
            except Exception:
                self._wandb = None

    def init(self, project: str = "fedml", name: str = "", config: Any = None) -> "MetricsSink":
        self.run.name = name
        cfg = vars(config) if hasattr(config, "__dict__") else dict(config or {})
        self.run.config = {k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool, type(None)))}
        if self._wandb is not None:
            self._wandb.init(project=project, name=name, config=self.run.config)
        self._write({"_event": "init", "name": name, "config": self.run.config, "ts": time.time()})
        return self

    def log(self, data: Dict[str, Any], step: Optional[int] = None) -> None:
        rec = dict(data)
        if step is not None:
            rec["_step"] = step
        if self.keep:
            self.records.append(rec)
        self._write(rec)
        if self._wandb is not None:
            self._wandb.log(data)

    def log_rounds(self, rounds: List[int], columns: Dict[str, List[float]]) -> None:
        """Bulk flush of device-accumulated per-round metrics (one record per round per key group)."""
        for i, r in enumerate(rounds):
            rec = {k: v[i] for k, v in columns.items()}
            rec["round"] = r
            self.log(rec)

    def set_summary(self, key: str, value: Any) -> None:
        self.run.summary[key] = value
        if self._wandb is not None and self._wandb.run is not None:
            self._wandb.run.summary[key] = value

    def series(self, key: str) -> List[Any]:
        return [r[key] for r in self.records if key in r]

    def last(self, key: str, default=None):
        for r in reversed(self.records):
            if key in r:
                return r[key]
        return default

    def _write(self, rec) -> None:
        if self._fh is not None:
            self._fh.write(json.dumps(rec, default=_default) + "\n")
            self._fh.flush()

    def finish(self) -> None:
        self._write({"_event": "summary", "summary": self.run.summary})
        if self._fh is not None:
            self._fh.close()
            self._fh = None


def _default(o):
    try:
        import numpy as np
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        if isinstance(o, np.ndarray):
            return o.tolist()
    except Exception:
        pass
    return str(o)


_GLOBAL = MetricsSink()


def get_sink() -> MetricsSink:
    return _GLOBAL


def set_sink(sink: MetricsSink) -> MetricsSink:
    global _GLOBAL
    _GLOBAL = sink
    return sink
