"""Where does a k-block of the tcgen05 mainloop spend its time?  FDB_GEMM_DBG switches off the operand loads (1), the MMAs (2) and
the output stores (4) of gemm_tn_kernel; this prints the kernel time of a conv forward and a plain GEMM under the combination
given in the environment (run once per setting; results of the switched-off runs are numerically garbage by design)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200.ops import _ext  # noqa: E402

ext = _ext.load(required=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return round(tot / n * 1e3, 1)


r = {"dbg": int(os.environ.get("FDB_GEMM_DBG", "0"))}
for name, (N, HW, C, Co) in {"conv64": (32, 56, 64, 64), "conv128": (32, 28, 128, 128), "conv512": (64, 7, 512, 512)}.items():
    xb = torch.randn(N, HW, HW, C, device="cuda").bfloat16()
    wq = (torch.randn(Co, 3, 3, C, device="cuda") / (9 * C) ** 0.5).bfloat16()
    r[name + "_us"] = timeit(lambda: ext.conv_tma_fwd(xb, wq, None, 1, 1, False, 1))
    if r["dbg"] & 8:
        c = ext.gemm_debug_counters()
        r[name + "_cycles"] = {"producer_wait_empty": c[0], "producer_total": c[1], "kblocks": c[2], "mma_wait_full": c[3],
                               "mma_wait_acc": c[4], "mma_total": c[5], "tiles": c[6], "epi_wait_acc_full": c[7], "epi_total": c[8]}
for n in (1024, 4096):
    A = torch.randn(n, n, device="cuda").bfloat16()
    B = torch.randn(n, n, device="cuda").bfloat16()
    r[f"gemm{n}_us"] = timeit(lambda: ext.gemm_tn_bias_act(A, B, None, False, False))
    if r["dbg"] & 8:
        c = ext.gemm_debug_counters()
        r[f"gemm{n}_cycles"] = {"producer_wait_empty": c[0], "producer_total": c[1], "kblocks": c[2], "mma_wait_full": c[3],
                                "mma_wait_acc": c[4], "mma_total": c[5], "tiles": c[6], "epi_wait_acc_full": c[7], "epi_total": c[8]}
print(json.dumps(r), flush=True)
