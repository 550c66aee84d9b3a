"""Cold (L2 flushed) / warm timings of ops.gram_cosine (2 launches: gram_kernel + gram_finish_kernel) over CFL-sized inputs."""
import json
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def cold(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for i in range(iters):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6482.7
for n, P in ((10, 1_199_882), (22, 1_199_882), (8, 11_183_582), (32, 4_000_000), (16, 8_000_000), (10, 13_000_000), (24, 6_000_000)):
    U = torch.randn(n, P, device="cuda")
    c = cold(lambda: ops.gram_cosine(U))
    gbs = n * P * 4 / (c * 1e-3) / 1e9
    print(json.dumps({"kernel": f"gram_cosine/{n}x{P}", "cold_us": round(c * 1e3, 1), "GBps": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / peak, 3)}), flush=True)
