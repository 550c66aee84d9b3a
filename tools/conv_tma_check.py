"""TMA-im2col convolution path (gemm_tc.cu conv modes): numerics vs F.conv2d on bf16-rounded operands + kernel times.
Prints one JSON line per geometry; any mismatch is reported, not raised, so one run shows every failing direction."""
import json
import sys

import torch
import torch.nn.functional as F

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


def rel(a, b):
    return round((a - b).abs().max().item() / (b.abs().max().item() + 1e-9), 5)


GEOMS = [(2, 64, 8, 8, 64, 3, 1, 1), (2, 64, 16, 16, 128, 3, 2, 1), (2, 64, 16, 16, 128, 1, 2, 0), (3, 128, 9, 11, 256, 3, 1, 1),
         (2, 64, 13, 13, 64, 3, 1, 0), (2, 64, 15, 15, 64, 3, 2, 1),
         (32, 64, 56, 56, 64, 3, 1, 1), (32, 128, 28, 28, 128, 3, 1, 1), (32, 256, 14, 14, 256, 3, 1, 1), (32, 64, 56, 56, 128, 3, 2, 1),
         (64, 512, 7, 7, 512, 3, 1, 1)]
if len(sys.argv) > 1:
    GEOMS = GEOMS[: int(sys.argv[1])]
for (N, Ci, H, W, Co, k, st, pad) in GEOMS:
    torch.manual_seed(0)
    x = torch.randn(N, H, W, Ci, device="cuda")
    w = torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5
    b = torch.randn(Co, device="cuda")
    xb = ext.conv_cast_bf16(x, None)
    assert torch.equal(xb, x.bfloat16())
    wq = ext.conv_cast_bf16(w.permute(0, 2, 3, 1).contiguous(), None)
    xr = xb.float().permute(0, 3, 1, 2)
    wr = w.bfloat16().float()
    ref = F.conv2d(xr, wr, b, st, pad)
    r = {"geom": [N, Ci, H, W, Co, k, st, pad]}
    y = ext.conv_tma_fwd(xb, wq, b, st, pad, False, 1)
    torch.cuda.synchronize()
    r["fwd_err"] = rel(y.permute(0, 3, 1, 2), ref)
    dy = torch.randn_like(y)
    dyb = ext.conv_cast_bf16(dy, None)
    dyr = dyb.float().permute(0, 3, 1, 2)
    gx, gw = torch.autograd.grad(F.conv2d(xr.requires_grad_(True), wr.requires_grad_(True), None, st, pad), (xr, wr), dyr)
    if st == 1 and Co % 64 == 0:
        dx = ext.conv_tma_dgrad(dyb, wq, pad, 1)
        torch.cuda.synchronize()
        r["dgrad_err"] = rel(dx.permute(0, 3, 1, 2), gx) if dx.shape[1:3] == (H, W) else f"shape {tuple(dx.shape)}"
    dw = torch.zeros(Co, k, k, Ci, device="cuda")
    ext.conv_tma_wgrad(xb, dyb, dw, k, st, pad, 1)
    torch.cuda.synchronize()
    r["wgrad_err"] = rel(dw.permute(0, 3, 1, 2), gw)
    if N >= 32:
        r["cast_x_us"] = timeit(lambda: ext.conv_cast_bf16(x, None))
        r["fwd_us"] = timeit(lambda: ext.conv_tma_fwd(xb, wq, b, st, pad, False, 1))
        if st == 1:
            r["dgrad_us"] = timeit(lambda: ext.conv_tma_dgrad(dyb, wq, pad, 1))
        r["wgrad_us"] = timeit(lambda: ext.conv_tma_wgrad(xb, dyb, dw, k, st, pad, 1))
        w_ohwi = w.permute(0, 2, 3, 1).contiguous()
        r["cast_w_us"] = timeit(lambda: ext.conv_cast_bf16(w_ohwi, None))
        xcl = xb.permute(0, 3, 1, 2)
        wcl = w.bfloat16().contiguous(memory_format=torch.channels_last)
        r["cudnn_bf16cl_fwd_us"] = timeit(lambda: F.conv2d(xcl, wcl, None, st, pad))
    print(json.dumps(r), flush=True)
