"""Implicit-GEMM convolution kernels (csrc/conv_igemm.cu) vs F.conv2d in fp32 (cuDNN): forward, dgrad, wgrad."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GEOMS = [
    # N, Cin, H, W, Cout, k, stride, pad
    (4, 32, 26, 26, 64, 3, 1, 0),      # CNN_DropOut conv2
    (2, 64, 16, 16, 64, 3, 1, 1),      # ResNet-18 layer1
    (2, 64, 16, 16, 128, 3, 2, 1),     # ResNet-18 layer2 downsampling 3x3
    (2, 64, 16, 16, 128, 1, 2, 0),     # ResNet-18 1x1 shortcut
    (3, 128, 8, 8, 256, 3, 1, 1),
    (2, 256, 4, 4, 512, 3, 2, 1),
    (5, 32, 9, 11, 32, 5, 1, 2),       # odd sizes, 5x5, small channels
    (2, 96, 7, 7, 96, 3, 1, 1),        # Cout multiple of 32 only
    (2, 512, 4, 4, 512, 3, 1, 1),      # deep layer, 32 output pixels: split-K forward / dgrad
    (3, 128, 9, 11, 192, 3, 1, 1),     # TMA path, odd sizes, partial pixel tile, Cout = 1.5 N tiles
    (2, 64, 13, 13, 64, 3, 1, 0),      # TMA path, no padding
    (2, 64, 15, 15, 64, 3, 2, 1),      # TMA forward / wgrad with traversal stride 2 on an odd image
    (2, 64, 12, 12, 64, 5, 1, 2),      # 5x5 taps
]


def test_conv_tma_ops_direct_and_gated_cast():
    """conv_cast_bf16 (+ ReLU gate) and the three TMA-im2col GEMM modes against F.conv2d on the same bf16-rounded operands."""
    from feddrift_b200.ops import _ext
    ext = _ext.load(required=True)
    torch.manual_seed(1)
    N, Ci, H, W, Co, k, st, pad = 3, 128, 10, 7, 64, 3, 1, 1
    x = torch.randn(N, H, W, Ci, device="cuda")
    gate = torch.randn(N, H, W, Ci, device="cuda")
    assert torch.equal(ext.conv_cast_bf16(x, None), x.bfloat16())
    assert torch.equal(ext.conv_cast_bf16(x, gate), (x * (gate > 0)).bfloat16())
    odd = torch.randn(1003, device="cuda")
    assert torch.equal(ext.conv_cast_bf16(odd, None), odd.bfloat16())
    w = torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5
    xb = ext.conv_cast_bf16(x, None)
    wq = ext.conv_cast_bf16(w.permute(0, 2, 3, 1).contiguous(), None)            # bf16 [Cout][kh][kw][Cin]
    xr = xb.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.bfloat16().float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, st, pad)
    y = ext.conv_tma_fwd(xb, wq, None, st, pad, True, 1)
    assert (y.permute(0, 3, 1, 2) - F.relu(ref)).abs().max().item() < 1e-3
    dy = torch.randn_like(y)
    dyb = ext.conv_cast_bf16(dy, None)
    gx, gw = torch.autograd.grad(ref, (xr, wr), dyb.float().permute(0, 3, 1, 2))
    dx = ext.conv_tma_dgrad(dyb, wq, pad, 1)                                          # reads the forward pack MN-major
    assert dx.shape == (N, H, W, Ci)
    assert (dx.permute(0, 3, 1, 2) - gx).abs().max().item() < 1e-3 * gx.abs().max().item() + 1e-4
    dw = torch.full((Co, k, k, Ci), 2.0, device="cuda")                            # the kernel ADDS into the channels_last buffer
    ext.conv_tma_wgrad(xb, dyb, dw, k, st, pad, 1)
    assert (dw.permute(0, 3, 1, 2) - 2.0 - gw).abs().max().item() < 1e-3 * gw.abs().max().item() + 1e-4
    # software-gather kernels with the same operand conventions
    wq_t = ext.conv_pack_t(wq)
    assert torch.equal(wq_t, wq.permute(3, 1, 2, 0).contiguous())
    dw2 = torch.zeros(Co, k, k, Ci, device="cuda")
    ext.conv_igemm_wgrad(xb.float(), dyb.float(), k, k, st, pad, pad, dw2, True)
    assert (dw2.permute(0, 3, 1, 2) - gw).abs().max().item() < 1e-3 * gw.abs().max().item() + 1e-4


def test_channels_last_parameter_grad_is_accumulated_in_place():
    """TcConv2d stores its weight channels_last; a preset channels_last .grad (the executor's flat gradient row) is the
    buffer the wgrad kernel reduce-adds into; a contiguous OIHW weight (plain tensor) still works through one transposing copy."""
    from feddrift_b200.ops import conv as C
    from feddrift_b200.ops.conv import TcConv2d
    torch.manual_seed(0)
    m = TcConv2d(64, 64, 3, padding=1, bias=False).cuda()
    assert m.weight.is_contiguous(memory_format=torch.channels_last)
    ref = torch.nn.Conv2d(64, 64, 3, padding=1, bias=False)
    torch.manual_seed(0)
    ref.reset_parameters()
    torch.manual_seed(0)
    m2 = TcConv2d(64, 64, 3, padding=1, bias=False)
    assert torch.equal(m2.weight.detach(), ref.weight.detach())                    # same values as nn.Conv2d for a given seed
    x = torch.randn(2, 64, 9, 9, device="cuda", requires_grad=True)
    g0 = torch.full_like(m.weight, 0.5)                                            # preserves channels_last
    m.weight.grad = g0
    y = m(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert m.weight.grad is g0
    xr = x.detach().bfloat16().float()
    wr = m.weight.detach().bfloat16().float().contiguous().requires_grad_(True)
    gw, = torch.autograd.grad(F.conv2d(xr, wr, None, 1, 1), wr, dy.bfloat16().float())
    assert (g0 - 0.5 - gw).abs().max().item() < 2e-3 * gw.abs().max().item() + 1e-4
    wc = m.weight.detach().contiguous().clone().requires_grad_(True)               # contiguous OIHW weight
    y2 = C._ConvIgemmFn.apply(x, wc, None, (1, 1), (1, 1), False)
    assert (y2 - y).abs().max().item() < 1e-5
    gwc, = torch.autograd.grad(y2, wc, dy)
    assert (gwc - gw).abs().max().item() < 2e-3 * gw.abs().max().item() + 1e-4


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_igemm_matches_conv2d(geom):
    from feddrift_b200.ops import conv as C
    N, Ci, H, W, Co, k, st, pad = geom
    torch.manual_seed(0)
    x = torch.randn(N, Ci, H, W, device="cuda", requires_grad=True)
    w = (torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Co, device="cuda", requires_grad=True)
    ref = F.conv2d(x, w, b, st, pad)
    dy = torch.randn_like(ref)
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
    n0 = dict(C.IGEMM_CALLS)
    got = C._ConvIgemmFn.apply(x, w, b, (st, st), (pad, pad), False)
    hx, hw, hb = torch.autograd.grad(got, (x, w, b), dy)
    torch.cuda.synchronize()
    assert C.IGEMM_CALLS["fwd"] == n0["fwd"] + 1 and C.IGEMM_CALLS["dgrad"] == n0["dgrad"] + 1 and C.IGEMM_CALLS["wgrad"] == n0["wgrad"] + 1
    assert got.shape == ref.shape
    # the GEMM-mainloop path with the TMA-im2col producer must be the one that ran wherever the shape allows it
    exp_tma = (int(Ci % 64 == 0), int(Co % 64 == 0), int(Ci % 64 == 0))      # strided data gradients run on zero-dilated dY
    assert (C.IGEMM_CALLS["tma_fwd"] - n0["tma_fwd"], C.IGEMM_CALLS["tma_dgrad"] - n0["tma_dgrad"],
            C.IGEMM_CALLS["tma_wgrad"] - n0["tma_wgrad"]) == exp_tma

    def rel(a, b_):
        return (a - b_).abs().max().item() / (b_.abs().max().item() + 1e-6)

    assert rel(got, ref) < 2e-2, f"forward rel err {rel(got, ref)}"
    assert rel(hx, gx) < 2e-2, f"dgrad rel err {rel(hx, gx)}"
    assert rel(hw, gw) < 2e-2, f"wgrad rel err {rel(hw, gw)}"
    assert rel(hb, gb) < 1e-4


def test_tcconv2d_module_relu_and_cnn_model_use_igemm():
    from feddrift_b200.ops import conv as C
    from feddrift_b200.ops.conv import TcConv2d
    torch.manual_seed(0)
    m = TcConv2d(32, 64, 3, padding=1, activation="relu").cuda()
    x = torch.randn(2, 32, 10, 10, device="cuda", requires_grad=True)
    n0 = C.IGEMM_CALLS["fwd"]
    y = m(x)
    assert C.IGEMM_CALLS["fwd"] == n0 + 1
    ref = F.relu(F.conv2d(x, m.weight, m.bias, 1, 1))
    assert (y - ref).abs().max().item() < 3e-2
    y.sum().backward()
    gx = x.grad.clone()
    x.grad = None
    ref.sum().backward()
    assert (gx - x.grad).abs().max().item() / (x.grad.abs().max().item() + 1e-6) < 3e-2


def test_unsupported_channel_counts_fall_back_to_library_conv():
    from feddrift_b200.ops import _ext
    from feddrift_b200.ops.conv import TcConv2d
    m = TcConv2d(3, 64, 3, padding=1).cuda()          # 3-channel stem → F.conv2d
    y = m(torch.randn(2, 3, 8, 8, device="cuda"))
    assert y.shape == (2, 64, 8, 8)


def test_torchvision_resnet18_body_runs_on_igemm_and_matches_library_convs():
    import os
    from feddrift_b200.models.resnet_tv import resnet18
    from feddrift_b200.ops import conv as C
    torch.manual_seed(0)
    m = resnet18(10, small_input=True).cuda().train()
    x = torch.randn(4, 3, 32, 32, device="cuda")
    n0 = C.IGEMM_CALLS["fwd"]
    y = m(x)
    loss = y.square().mean()
    loss.backward()
    assert C.IGEMM_CALLS["fwd"] - n0 >= 16          # every body conv (stem excluded)
    g_ours = {k: p.grad.clone() for k, p in m.named_parameters()}
    for p in m.parameters():
        p.grad = None
    os.environ["FDB_NO_TC_CONV"] = "1"
    try:
        y2 = m(x)
        y2.square().mean().backward()
    finally:
        os.environ.pop("FDB_NO_TC_CONV")
    def cos(a, b):
        return torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()

    # bf16 tensor-core operands through 17 BatchNorm'd layers at batch 4: compare directions, not digits
    assert cos(y, y2) > 0.99, cos(y, y2)
    k = "layer1.0.conv1.weight"
    assert cos(g_ours[k], m.get_parameter(k).grad) > 0.9, cos(g_ours[k], m.get_parameter(k).grad)



@pytest.mark.parametrize("geom", [(3, 4, 64, 9, 9, 64, 3, 1, 1), (2, 3, 64, 8, 8, 128, 3, 2, 1), (4, 2, 128, 6, 6, 64, 1, 2, 0),
                                  (5, 2, 64, 7, 5, 192, 3, 1, 1)])
def test_grouped_tma_conv_matches_per_group(geom):
    """Grouped mode (one group per stacked pair): forward, data gradient (incl. zero-dilated strided layers) and the weight
    gradient written through the 3-D tensor map into strided gradient rows, against per-group F.conv2d."""
    from feddrift_b200.ops import _ext
    from feddrift_b200.ops import conv as C
    ext = _ext.load(required=True)
    G, N, Ci, H, W, Co, k, st, pad = geom
    torch.manual_seed(2)
    x = torch.randn(N, H, W, G * Ci, device="cuda")
    w = torch.randn(G, Co, k, k, Ci, device="cuda") / (Ci * k * k) ** 0.5           # per-group (O, kh, kw, I)
    bias = torch.randn(G * Co, device="cuda")
    stage = torch.zeros(G, Co * k * k * Ci + 40, device="cuda")                       # rows with a stride ≠ numel, like the parameter rows
    stage[:, :Co * k * k * Ci] = w.reshape(G, -1)
    wq = ext.conv_cast_rows_bf16(stage[:, :Co * k * k * Ci]).view(G * Co, k, k, Ci)
    assert torch.equal(wq.view(G, -1), w.reshape(G, -1).bfloat16())
    xb = ext.conv_cast_bf16(x, None)
    y = ext.conv_tma_fwd(xb, wq, bias, st, pad, False, G)
    xr = xb.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.bfloat16().float().permute(0, 1, 4, 2, 3).reshape(G * Co, Ci, k, k).requires_grad_(True)
    ref = F.conv2d(xr, wr, bias, st, pad, 1, G)
    assert (y.permute(0, 3, 1, 2) - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    dy = torch.randn_like(y)
    dyb = ext.conv_cast_bf16(dy, None)
    gx, gw = torch.autograd.grad(ref, (xr, wr), dyb.float().permute(0, 3, 1, 2))
    dx = ext.conv_tma_dgrad(C._dilate(dyb, st, H, W, k, pad), wq, pad, G)
    assert dx.shape == (N, H, W, G * Ci)
    assert (dx.permute(0, 3, 1, 2) - gx).abs().max().item() < 2e-3 * gx.abs().max().item()
    grows = torch.full((G, Co * k * k * Ci + 40), 0.25, device="cuda")
    ext.conv_tma_wgrad(xb, dyb, grows[:, :Co * k * k * Ci], k, st, pad, G)
    got = (grows[:, :Co * k * k * Ci] - 0.25).view(G, Co, k, k, Ci).permute(0, 1, 4, 2, 3).reshape(G * Co, Ci, k, k)
    assert (got - gw).abs().max().item() < 2e-3 * gw.abs().max().item()
    assert torch.all(grows[:, Co * k * k * Ci:] == 0.25)                               # nothing spilled past a group's segment


def test_stacked_resnet_on_gpu_uses_grouped_kernels_and_matches_per_pair():
    import copy
    from feddrift_b200.models import resnet
    from feddrift_b200.models.utils import flat_size, flat_spec, flat_view, flatten_state_dict, unflatten_to_state_dict
    from feddrift_b200.ops import conv as C
    from feddrift_b200.sim import stacked as S
    torch.manual_seed(0)
    tmpl = resnet.ResNet(resnet.BasicBlock, [1, 1], 10, widths=(64, 128))
    n, B = 3, 4
    spec, P = flat_spec(tmpl), flat_size(tmpl)
    rows = []
    for i in range(n):
        m = copy.deepcopy(tmpl)
        torch.manual_seed(10 + i)
        for layer in m.modules():
            if layer is not m and hasattr(layer, "reset_parameters"):
                layer.reset_parameters()
        rows.append(flatten_state_dict(m.state_dict()))
    rows = torch.stack(rows).cuda()
    x, y = torch.randn(n, B, 3, 12, 12, device="cuda"), torch.randint(0, 10, (n, B), device="cuda")
    gref, outs = torch.zeros(n, P, device="cuda"), []
    for i in range(n):
        mod = copy.deepcopy(tmpl).cuda()
        mod.load_state_dict(unflatten_to_state_dict(rows[i].clone(), spec))
        mod.train()
        out = mod(x[i])
        outs.append(out)
        F.cross_entropy(out, y[i]).backward()
        grads = {k: p.grad for k, p in mod.named_parameters()}
        for k, _, _, off, numel in spec:
            if grads.get(k) is not None:
                gref[i, off:off + numel] = flat_view(grads[k])
    net = S.stack_module(tmpl, n).cuda()
    stage, G = rows.clone(), torch.zeros(n, P, device="cuda")
    sp = {k: (tuple(shape), off, numel) for k, shape, _, off, numel in spec}
    for name, mod in net.named_modules():
        if isinstance(mod, S._Stacked):
            mod.bind(name, sp, stage, G)
    net.train()
    n0 = C.IGEMM_CALLS.get("stacked_fwd", 0)
    logits = net(S.stack_input(tmpl, x))
    assert C.IGEMM_CALLS.get("stacked_fwd", 0) - n0 == 5          # every 64/128-channel body conv incl. the strided one + its 1×1 shortcut
    ref = torch.stack(outs, 1).reshape(B, -1)
    assert (logits - ref).abs().max().item() < 3e-2 * (1 + ref.abs().max().item())
    (F.cross_entropy(logits.reshape(B * n, -1), y.t().reshape(-1), reduction="sum") / B).backward()
    cos = F.cosine_similarity(G.flatten(), gref.flatten(), dim=0).item()
    assert cos > 0.995, cos
