"""Is the TMA-im2col producer row-rate bound?  The same 1×1 convolution (a plain GEMM [pixels, C]·[Cout, C]ᵀ) through the im2col
producer (conv_tma_fwd) and through the tiled-TMA producer (gemm_tn_bias_act) of the SAME mainloop."""
import json
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


for (N, HW, C, Co) in ((32, 56, 64, 64), (32, 56, 576, 64), (32, 28, 128, 128), (32, 28, 1152, 128), (32, 14, 2304, 256)):
    xb = torch.randn(N, HW, HW, C, device="cuda").bfloat16()
    wq = (torch.randn(Co, 1, 1, C, device="cuda") / C ** 0.5).bfloat16()
    y1 = ext.conv_tma_fwd(xb, wq, None, 1, 0, False, 1)
    y2 = ext.gemm_tn_bias_act(xb.view(-1, C), wq.view(Co, C), None, False, True)
    err = (y1.view(-1, Co) - y2).abs().max().item()
    r = {"pixels": N * HW * HW, "C": C, "Cout": Co, "err": err,
         "im2col_us": timeit(lambda: ext.conv_tma_fwd(xb, wq, None, 1, 0, False, 1)),
         "tiled_us": timeit(lambda: ext.gemm_tn_bias_act(xb.view(-1, C), wq.view(Co, C), None, False, True))}
    print(json.dumps(r), flush=True)
