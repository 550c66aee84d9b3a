"""Latency / throughput probe of the tcgen05 GEMM and the Gram kernel.

For every shape prints the cold single-launch time (L2 flushed, one launch between two events), the warm
back-to-back time (50 launches between two events, launch overhead amortised) and the same for cuBLAS
(`F.linear` + relu).  Used to separate fixed per-launch cost from steady-state tile throughput."""
import json
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200 import ops  # noqa: E402
from feddrift_b200.ops import _ext  # noqa: E402

ext = _ext.load()
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


def warm(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


shapes = [(128, 128, 64), (128, 256, 64), (128, 256, 832), (500, 1568, 784), (512, 128, 9216), (1024, 1024, 1024),
          (2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 8192)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
for M, N, K in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    bb = bias.bfloat16()
    mine = lambda: ext.gemm_tn_bias_act(A, B, bias, True, False)  # noqa: E731
    lib = lambda: torch.relu(torch.nn.functional.linear(A, B, bb))  # noqa: E731
    fl = 2.0 * M * N * K
    r = {"shape": f"{M}x{N}x{K}", "cold_us": cold(mine) * 1e3, "warm_us": warm(mine) * 1e3, "cublas_cold_us": cold(lib) * 1e3,
         "cublas_warm_us": warm(lib) * 1e3}
    r["warm_TFLOPs"] = fl / (r["warm_us"] * 1e-6) / 1e12
    r["cublas_warm_TFLOPs"] = fl / (r["cublas_warm_us"] * 1e-6) / 1e12
    print(json.dumps(r), flush=True)

if len(sys.argv) == 1:
    for n, P in ((10, 1_199_882), (22, 1_199_882), (8, 11_183_582), (32, 4_000_000)):
        U = torch.randn(n, P, device="cuda")
        f = lambda: ops.gram_cosine(U)  # noqa: E731
        c, w = cold(f), warm(f, 20)
        print(json.dumps({"gram": f"{n}x{P}", "cold_us": c * 1e3, "warm_us": w * 1e3, "cold_GBps": n * P * 4 / (c * 1e-3) / 1e9}), flush=True)
