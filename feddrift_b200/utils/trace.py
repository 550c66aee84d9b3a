"""Tracing / profiling helpers (the reference only has integer-second ``time.time()`` prints).

* :func:`cuda_timer` — CUDA-event timing context on the current stream (what bench.py / tools use);
* :func:`nvtx_range` — NVTX ranges around engine phases (visible in ncu / nsys when available);
* :class:`PhaseClock` — host wall-clock accounting per phase (cluster / rounds / flush), printed by the CLI;
* in-kernel phase stamps: pass ``st["timers"] = int64[R,4]`` to the fused round kernel (``tools/phase_timing.py``);
* :func:`torch_profile` — ``torch.profiler`` trace export for the generic executor.
"""
from __future__ import annotations

import contextlib
import time
from collections import defaultdict
from typing import Dict

import torch


@contextlib.contextmanager
def cuda_timer(out: Dict[str, float], key: str = "ms"):
    if not torch.cuda.is_available():
        t0 = time.perf_counter()
        yield
        out[key] = 1e3 * (time.perf_counter() - t0)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    yield
    b.record()
    torch.cuda.synchronize()
    out[key] = a.elapsed_time(b)


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class PhaseClock:
    def __init__(self):
        self.t = defaultdict(float)
        self.n = defaultdict(int)

    @contextlib.contextmanager
    def phase(self, name: str, sync: bool = False):
        if sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with nvtx_range(name):
            yield
        if sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.t[name] += time.perf_counter() - t0
        self.n[name] += 1

    def report(self) -> Dict[str, Dict[str, float]]:
        return {k: {"seconds": v, "calls": self.n[k]} for k, v in self.t.items()}


@contextlib.contextmanager
def torch_profile(path: str):
    acts = [torch.profiler.ProfilerActivity.CPU] + ([torch.profiler.ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
    with torch.profiler.profile(activities=acts) as prof:
        yield prof
    prof.export_chrome_trace(path)
