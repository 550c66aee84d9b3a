"""Symmetric-memory plumbing for the multi-GPU fused round kernel.

Every rank allocates one identically-shaped buffer through ``torch.distributed._symmetric_memory`` (CUDA VMM +
fabric handles under the hood); after the rendezvous each rank holds a device pointer to EVERY peer's buffer, so
the kernel can ``st.global`` partial sums straight into the peers' inboxes over NVLink and publish
``st.release.sys`` epoch flags — no NCCL call on the aggregate/broadcast path (BASELINE.json north star).

Layout of the per-rank buffer (floats):   inbox [2 (parity)] [world] [M·P]   |   flags (u32) [2] [world]
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def _rendezvous(numel: int, device) -> Dict:
    import torch.distributed._symmetric_memory as symm_mem
    buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD.group_name)
    return {"buf": buf, "hdl": hdl, "ptrs": [int(p) for p in hdl.buffer_ptrs]}


def attach_multi_gpu(sim, world: int, rank: int) -> None:
    """Give ``sim`` the symmetric inbox/flag pointers; clients are then sharded ``c % world == rank``."""
    MP = sim.M * sim.bank.P
    inbox_floats = 2 * world * MP
    inbox_floats = (inbox_floats + 31) // 32 * 32  # keep the flag words on their own 128-byte line
    flag_words = 3 * world
    flag_floats = max((flag_words + 31) // 32 * 32, 32)
    max_rounds = int(getattr(sim.args, "max_rounds_per_launch", 0) or max(sim.args.comm_round, 256))
    metric_floats = max_rounds * sim.C * 4
    r = _rendezvous(inbox_floats + flag_floats + metric_floats, sim.device)
    base = r["ptrs"]
    moff = inbox_floats + flag_floats
    sim.multi = {
        "world": world, "rank": rank, "flag_base": 0,
        "inbox_ptrs": base, "flag_ptrs": [p + 4 * inbox_floats for p in base],
        "metrics_ptrs": [p + 4 * moff for p in base], "metrics_buf": r["buf"][moff:moff + metric_floats],
        "metrics_rounds": max_rounds,
        "error_flag": torch.zeros(1, dtype=torch.int32, device=sim.device),
        "_keepalive": r,
    }
    torch.cuda.synchronize()
    dist.barrier()


def check_error(sim) -> None:
    m = getattr(sim, "multi", None)
    if m is not None and int(m["error_flag"].item()) != 0:
        raise RuntimeError("fed_round_small: a peer rank never published its round flag (spin timeout)")
