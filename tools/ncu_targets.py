"""Tiny driver for `ncu` captures of single kernels (one GPU, few launches).

    ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 2 -c 1 -o gpurun_out/gemm_tn \
        python tools/ncu_targets.py gemm 4096
    ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/gram \
        python tools/ncu_targets.py gram 16 8000000
"""
import sys

import torch

sys.path.insert(0, ".")
from feddrift_b200 import ops  # noqa: E402
from feddrift_b200.ops import _ext  # noqa: E402

what = sys.argv[1]
if what == "gemm":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    A = torch.randn(n, n, device="cuda").bfloat16()
    B = torch.randn(n, n, device="cuda").bfloat16()
    bias = torch.randn(n, device="cuda")
    ext = _ext.load()
    for _ in range(4):
        ext.gemm_tn_bias_act(A, B, bias, True, False)
elif what == "gram":
    n, P = int(sys.argv[2]), int(sys.argv[3])
    U = torch.randn(n, P, device="cuda")
    for _ in range(4):
        ops.gram_cosine(U)
torch.cuda.synchronize()
