"""One forward / dgrad / wgrad launch of the implicit-GEMM conv kernels on a ResNet-18 layer1 shape (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from feddrift_b200.ops import _ext
ext = _ext.load(required=True)
B, cin, cout, k, stride, pad, hw = 32, 64, 64, 3, 1, 1, 56
if len(sys.argv) > 1 and sys.argv[1] == "s2":
    B, cin, cout, k, stride, pad, hw = 32, 64, 128, 3, 2, 1, 56
torch.manual_seed(0)
x = torch.randn(B, hw, hw, cin, device="cuda")
w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
Ho = (hw + 2 * pad - k) // stride + 1
dy = torch.randn(B, Ho, Ho, cout, device="cuda")
for _ in range(2):
    wpk = ext.conv_pack_weights(w)
    y = ext.conv_igemm_fwd(x, wpk[0], None, stride, pad, pad, False)
    dx = ext.conv_igemm_dgrad(dy, wpk[1], hw, hw, stride, pad, pad)
    dw = ext.conv_igemm_wgrad(x, dy, k, k, stride, pad, pad, None)
torch.cuda.synchronize()
