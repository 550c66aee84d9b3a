"""Forward+backward time of TcConv2d (im2col + tcgen05 GEMM) vs nn.Conv2d (cuDNN, fp32 and bf16-autocast) on the conv
shapes of the model zoo (CNN_DropOut conv2, ResNet-18 stages)."""
import json
import sys

import torch
from torch import nn

sys.path.insert(0, ".")
from feddrift_b200.ops.conv import TcConv2d  # noqa: E402


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for B, cin, cout, k, stride, pad, hw in ((50, 32, 64, 3, 1, 0, 26), (32, 64, 64, 3, 1, 1, 56), (32, 128, 128, 3, 1, 1, 28),
                                          (32, 256, 256, 3, 1, 1, 14), (32, 64, 128, 3, 2, 1, 56), (64, 512, 512, 3, 1, 1, 7)):
    torch.manual_seed(0)
    tc = TcConv2d(cin, cout, k, stride=stride, padding=pad).cuda()
    ref = nn.Conv2d(cin, cout, k, stride=stride, padding=pad).cuda()
    x = torch.randn(B, cin, hw, hw, device="cuda", requires_grad=True)
    xcl = x.detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)

    def run(layer, inp, autocast=False):
        def f():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                y = layer(inp)
            y.float().square().mean().backward()
        return f
    r = {"shape": f"B{B} {cin}->{cout} k{k} s{stride} p{pad} {hw}x{hw}", "tcconv_us": timeit(run(tc, x)),
         "cudnn_fp32_us": timeit(run(ref, x)), "cudnn_bf16_cl_us": timeit(run(ref.to(memory_format=torch.channels_last), xcl, True))}
    Ho = (hw + 2 * pad - k) // stride + 1
    r["GFLOP_fwd_bwd"] = 3 * 2.0 * B * Ho * Ho * cout * cin * k * k / 1e9
    print(json.dumps(r), flush=True)
