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
elif what == "aggregate":     # K1: per-cluster weighted reduce + apply over the client arena (ResNet-18 sized rows)
    C, M, P = 32, 2, 11183644
    theta = torch.zeros(M, P, device="cuda")
    params = torch.randn(C, M, P, device="cuda")
    n = torch.rand(C, M, device="cuda")
    for _ in range(3):
        ops.cluster_aggregate_(theta, params, n)
elif what == "adam":          # fused arena optimizer on 32 ResNet-18 rows (the stacked executor's update)
    R, P = 32, 11183644
    p_, g, m, v, vm = (torch.randn(R, P, device="cuda") for _ in range(5))
    v.abs_(); vm.abs_()
    st = torch.zeros(R, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ops.adam_amsgrad_rows_(p_, g, m, v, vm, st, 0.01, 0.0)
elif what == "bn":            # NHWC training BatchNorm over 32·64 stacked channels
    x = torch.randn(32, 2048, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w, b = torch.ones(2048, device="cuda", requires_grad=True), torch.zeros(2048, device="cuda", requires_grad=True)
    rm, rv = torch.zeros(2048, device="cuda"), torch.ones(2048, device="cuda")
    for _ in range(3):
        ops.batch_norm_train_nhwc(x, w, b, rm, rv, 1e-5, 0.1).sum().backward()
torch.cuda.synchronize()
