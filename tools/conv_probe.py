"""Implicit-GEMM conv kernels (TMA-im2col GEMM modes of csrc/gemm_tc.cu, gather kernels of csrc/conv_igemm.cu) vs cuDNN on the conv shapes of the model zoo (CNN_DropOut conv2, ResNet-18
stages): per-direction kernel time (forward / dgrad / wgrad, CUDA events, L2-flushing 256 MiB write between timed launches) and
module-level forward+backward time (TcConv2d vs nn.Conv2d fp32 NCHW vs nn.Conv2d bf16-autocast channels_last).  Clocks recorded."""
import json
import sys

import torch
import torch.nn.functional as F
from torch import nn

sys.path.insert(0, ".")
from bench import ClockSampler  # noqa: E402
from feddrift_b200.ops import _ext  # noqa: E402
from feddrift_b200.ops.conv import TcConv2d  # noqa: E402

ext = _ext.load(required=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20, do_flush=True):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        if do_flush:
            flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


clk = ClockSampler(0)
clk.start()
rows = []
for B, cin, cout, k, stride, pad, hw in ((50, 32, 64, 3, 1, 0, 26), (32, 64, 64, 3, 1, 1, 56), (32, 128, 128, 3, 1, 1, 28),
                                          (32, 256, 256, 3, 1, 1, 14), (32, 64, 128, 3, 2, 1, 56), (64, 512, 512, 3, 1, 1, 7)):
    torch.manual_seed(0)
    tc = TcConv2d(cin, cout, k, stride=stride, padding=pad).cuda()
    ref = nn.Conv2d(cin, cout, k, stride=stride, padding=pad).cuda()
    x = torch.randn(B, cin, hw, hw, device="cuda")
    xcl = x.contiguous(memory_format=torch.channels_last)
    xh = xcl.permute(0, 2, 3, 1)                                   # NHWC view
    w = tc.weight.detach().contiguous()
    w_ohwi = tc.weight.detach().permute(0, 2, 3, 1).contiguous()        # free view: TcConv2d stores its weight channels_last
    wq = ext.conv_cast_bf16(w_ohwi, None)
    wq_t = ext.conv_pack_t(wq)
    dwbuf = torch.zeros(cout, k, k, cin, device="cuda")
    Ho = (hw + 2 * pad - k) // stride + 1
    dyh = torch.randn(B, Ho, Ho, cout, device="cuda")
    dy_nchw = dyh.permute(0, 3, 1, 2).contiguous()
    dy_cl = dyh.permute(0, 3, 1, 2)                                # channels_last view
    wcl = w.contiguous(memory_format=torch.channels_last)
    xb, wb, dyb = xcl.bfloat16(), wcl.bfloat16(), dy_cl.bfloat16()
    mask = [True, True, False]
    r = {"shape": f"B{B} {cin}->{cout} k{k} s{stride} p{pad} {hw}x{hw}", "GFLOP_per_dir": 2.0 * B * Ho * Ho * cout * cin * k * k / 1e9}
    r["gather_fwd_us"] = timeit(lambda: ext.conv_igemm_fwd(xh, ext.conv_cast_bf16(w_ohwi, None), tc.bias.detach(), stride, pad, pad, False))
    r["gather_dgrad_us"] = timeit(lambda: ext.conv_igemm_dgrad(dyh, wq_t, hw, hw, stride, pad, pad))
    r["gather_wgrad_us"] = timeit(lambda: ext.conv_igemm_wgrad(xh, dyh, k, k, stride, pad, pad, dwbuf, True))
    r["ours_fwd_us"], r["ours_dgrad_us"], r["ours_wgrad_us"] = r["gather_fwd_us"], r["gather_dgrad_us"], r["gather_wgrad_us"]
    if cin % 64 == 0:      # TMA-im2col GEMM path: bf16 NHWC operands; the casts are charged to the directions that need them
        xh_c = xh.contiguous()
        xbh, dybh = ext.conv_cast_bf16(xh_c, None), ext.conv_cast_bf16(dyh, None)
        r["cast_x_us"] = timeit(lambda: ext.conv_cast_bf16(xh_c, None))
        r["cast_dy_us"] = timeit(lambda: ext.conv_cast_bf16(dyh, None))
        r["tma_fwd_us"] = timeit(lambda: ext.conv_tma_fwd(xbh, ext.conv_cast_bf16(w_ohwi, None), tc.bias.detach(), stride, pad, False, 1))
        r["tma_wgrad_us"] = timeit(lambda: ext.conv_tma_wgrad(xbh, dybh, dwbuf, k, stride, pad, 1))
        r["ours_fwd_us"] = r["tma_fwd_us"] + r["cast_x_us"]
        r["ours_wgrad_us"] = r["tma_wgrad_us"] + r["cast_dy_us"]
        if stride == 1 and cout % 64 == 0:
            r["tma_dgrad_us"] = timeit(lambda: ext.conv_tma_dgrad(dybh, wq, pad, 1))
            r["ours_dgrad_us"] = r["tma_dgrad_us"]
    r["cudnn_fp32_fwd_us"] = timeit(lambda: F.conv2d(x, w, tc.bias.detach(), stride, pad))
    r["cudnn_fp32_bwd_us"] = timeit(lambda: torch.ops.aten.convolution_backward(dy_nchw, x, w, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1, mask))
    r["cudnn_bf16cl_fwd_us"] = timeit(lambda: F.conv2d(xb, wb, None, stride, pad))
    r["cudnn_bf16cl_bwd_us"] = timeit(lambda: torch.ops.aten.convolution_backward(dyb, xb, wb, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1, mask))
    r["ours_total_us"] = r["ours_fwd_us"] + r["ours_dgrad_us"] + r["ours_wgrad_us"]
    r["cudnn_fp32_total_us"] = r["cudnn_fp32_fwd_us"] + r["cudnn_fp32_bwd_us"]
    r["cudnn_bf16cl_total_us"] = r["cudnn_bf16cl_fwd_us"] + r["cudnn_bf16cl_bwd_us"]
    r["ours_tflops"] = 3 * r["GFLOP_per_dir"] / r["ours_total_us"] * 1e3 / 1e3

    # module level, warm caches (what an autograd training step pays)
    xr = x.clone().requires_grad_(True)
    xclr = xcl.clone().requires_grad_(True)

    def run(layer, inp, autocast=False):
        def f():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                y = layer(inp)
            y.backward(dy_cl if y.is_contiguous(memory_format=torch.channels_last) else dy_nchw)
        return f
    r["module_ours_us"] = timeit(run(tc, xclr), do_flush=False)
    r["module_cudnn_fp32_us"] = timeit(run(ref, xr), do_flush=False)
    rows.append(r)
c = clk.stop()
for r in rows:
    r["clocks"] = c
    print(json.dumps({k_: (round(v, 2) if isinstance(v, float) else v) for k_, v in r.items()}), flush=True)
