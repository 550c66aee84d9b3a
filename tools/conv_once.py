"""One forward / dgrad / wgrad launch of the implicit-GEMM conv kernels on a ResNet-18 layer1 shape (for ncu).
Default: the TMA-im2col GEMM modes (gemm_tn_kernel); `gather` = the software-gather kernels; `s2` = the strided layer-2 shape."""
import sys
import torch
sys.path.insert(0, ".")
from feddrift_b200.ops import _ext
ext = _ext.load(required=True)
B, cin, cout, k, stride, pad, hw = 32, 64, 64, 3, 1, 1, 56
if "s2" in sys.argv[1:]:
    B, cin, cout, k, stride, pad, hw = 32, 64, 128, 3, 2, 1, 56
if "deep" in sys.argv[1:]:
    B, cin, cout, k, stride, pad, hw = 64, 512, 512, 3, 1, 1, 7
torch.manual_seed(0)
x = torch.randn(B, hw, hw, cin, device="cuda")
w = (torch.randn(cout, cin, k, k, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
Ho = (hw + 2 * pad - k) // stride + 1
dy = torch.randn(B, Ho, Ho, cout, device="cuda")
dw = torch.zeros(cout, k, k, cin, device="cuda")
for _ in range(2):
    wq = ext.conv_cast_bf16(w.permute(0, 2, 3, 1), None)
    if "gather" in sys.argv[1:]:
        y = ext.conv_igemm_fwd(x, wq, None, stride, pad, pad, False)
        dx = ext.conv_igemm_dgrad(dy, ext.conv_pack_t(wq), hw, hw, stride, pad, pad)
        ext.conv_igemm_wgrad(x, dy, k, k, stride, pad, pad, dw, True)
    else:
        xb, dyb = ext.conv_cast_bf16(x, None), ext.conv_cast_bf16(dy, None)
        y = ext.conv_tma_fwd(xb, wq, None, stride, pad, False, 1)
        if stride == 1:
            dx = ext.conv_tma_dgrad(dyb, wq, pad, 1)
        ext.conv_tma_wgrad(xb, dyb, dw, k, stride, pad, 1)
torch.cuda.synchronize()
