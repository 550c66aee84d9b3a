"""Roofline micro-benchmarks of the streaming / GEMM kernels (CUDA events, warm-up, L2 flush between iterations).
Prints one JSON line per kernel with achieved bandwidth / FLOP rate and the fraction of the MEASURED peaks
(MEASURED_PEAKS.json: STREAM-copy HBM bandwidth and cuBLAS bf16 throughput)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200 import ops  # noqa: E402
from feddrift_b200.ops import _ext  # noqa: E402

peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
if os.path.exists("MEASURED_PEAKS.json"):
    p = json.load(open("MEASURED_PEAKS.json"))
    peaks = {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "src": "measured"}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, bytes_=None, flops=None, **extra):
    out = {"kernel": name, "ms": ms, **extra}
    if bytes_:
        out["GBps"] = bytes_ / ms / 1e6
        out["frac_of_%s_hbm" % peaks["src"]] = out["GBps"] / peaks["hbm_gbs"]
    if flops:
        out["TFLOPs"] = flops / ms / 1e9
        out["frac_of_%s_bf16" % peaks["src"]] = out["TFLOPs"] / peaks["bf16_tflops"]
    print(json.dumps(out))


# K1: per-cluster weighted FedAvg over the client arena (ResNet-18-sized rows, 32 clients, 2 clusters; CNN 64 clients)
for name, C, M, P in (("resnet18_32clients", 32, 2, 11_699_132 // 4 * 4), ("cnn_64clients", 64, 4, 1_199_882 // 4 * 4)):
    cp = torch.randn(C, M, P, device="cuda")
    n = torch.randint(1, 5, (C, M), device="cuda").float()
    theta = torch.zeros(M, P, device="cuda")
    ms = timeit(lambda: ops.cluster_aggregate_(theta, cp, n))
    report("cluster_aggregate/" + name, ms, bytes_=(C * M * P + M * P) * 4, C=C, M=M, P=P)
    del cp, theta
    torch.cuda.empty_cache()

# fused Adam(amsgrad) over arena rows: 16 clients × CNN
R, P = 16, 1_199_880
p_, g_ = torch.randn(R, P, device="cuda"), torch.randn(R, P, device="cuda")
m_, v_, x_ = torch.zeros_like(p_), torch.zeros_like(p_), torch.zeros_like(p_)
st = torch.zeros(R, dtype=torch.int32, device="cuda")
ms = timeit(lambda: ops.adam_amsgrad_rows_(p_, g_, m_, v_, x_, st, 1e-3, 1e-3))
report("adam_amsgrad_rows/16xCNN", ms, bytes_=R * P * 4 * (5 + 4), R=R, P=P)
del p_, g_, m_, v_, x_

# robust clip (norm + apply) over 32 ResNet-18 rows
R, P = 32, 11_699_132
rows, g = torch.randn(R, P, device="cuda"), torch.randn(P, device="cuda")
ms = timeit(lambda: ops.robust_clip_(rows, g, 5.0), iters=10)
report("robust_clip/32xResNet18", ms, bytes_=R * P * 4 * 3, R=R, P=P)
del rows, g
torch.cuda.empty_cache()

# Gram matrix (CFL) : 10 updates of CNN size
U = torch.randn(10, 1_199_882, device="cuda")
ms = timeit(lambda: ops.gram_cosine(U))
report("gram_cosine/10xCNN", ms, bytes_=U.numel() * 4)

# tcgen05 GEMM (TcLinear shapes + a large square for the tensor-core ceiling of this 1-CTA design)
ext = _ext.load()
for M_, N_, K_ in ((500, 1568, 784), (512, 128, 9216), (4096, 4096, 4096), (8192, 8192, 8192)):
    A = torch.randn(M_, K_, device="cuda").bfloat16()
    B = torch.randn(N_, K_, device="cuda").bfloat16()
    bias = torch.randn(N_, device="cuda")
    ms = timeit(lambda: ext.gemm_tn_bias_act(A, B, bias, True, False), iters=10)
    ms_ref = timeit(lambda: torch.relu(torch.nn.functional.linear(A, B, bias.bfloat16())), iters=10)
    report(f"gemm_tn_tcgen05/{M_}x{N_}x{K_}", ms, flops=2.0 * M_ * N_ * K_, cublas_ms=ms_ref)
