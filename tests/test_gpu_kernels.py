"""GPU numerics of the streaming / reduction / GEMM kernels vs plain PyTorch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

from feddrift_b200 import ops
from feddrift_b200.ops import reference as ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(10, 4, 38), (32, 3, 4096), (7, 2, 100003)])
def test_cluster_aggregate(shape):
    C, M, P = shape
    cp = torch.randn(C, M, P)
    n = torch.randint(0, 5, (C, M)).float()
    n[:, 0] = 0  # an unused model must stay untouched
    theta = torch.randn(M, P)
    th_ref = theta.clone()
    ref.cluster_aggregate_(th_ref, cp, n)
    th = theta.cuda()
    ops.cluster_aggregate_(th, cp.cuda(), n.cuda())
    assert torch.allclose(th.cpu(), th_ref, rtol=1e-5, atol=1e-5)


def test_weighted_average_and_merge_and_gossip():
    rows, w = torch.randn(9, 5000), torch.rand(9)
    assert torch.allclose(ops.weighted_average(rows.cuda(), w.cuda()).cpu(), ref.weighted_average(rows, w), atol=1e-5)
    th = torch.randn(4, 777)
    a = th.clone()
    ref.merge_axpby_(a, 0, 2, 0.3, 0.7)
    b = th.cuda()
    ops.merge_axpby_(b, 0, 2, 0.3, 0.7)
    assert torch.allclose(b.cpu(), a, atol=1e-6)
    X, Wm = torch.randn(8, 3000), torch.rand(8, 8)
    Wm = Wm / Wm.sum(1, keepdim=True)
    assert torch.allclose(ops.gossip_mix(X.cuda(), Wm.cuda()).cpu(), Wm @ X, atol=1e-5)


def test_robust_clip_and_ada_stats():
    rows, g = torch.randn(6, 10000) * 3, torch.randn(10000)
    a = rows.clone()
    n_ref = ref.robust_clip_(a, g, 5.0)
    b = rows.cuda()
    n_gpu = ops.robust_clip_(b, g.cuda(), 5.0)
    assert torch.allclose(b.cpu(), a, rtol=1e-4, atol=1e-5)
    assert torch.allclose(n_gpu.cpu(), n_ref, rtol=1e-4)
    # weak-DP: clip + counter-hash Gaussian noise fused in the same pass, masked entries pass through untouched
    mask = torch.rand(10000) > 0.1
    a, b = rows.clone(), rows.cuda()
    ref.robust_clip_(a, g, 5.0, mask, 0.05, 1234)
    ops.robust_clip_(b, g.cuda(), 5.0, mask.cuda(), 0.05, 1234)
    assert torch.allclose(b.cpu(), a, rtol=1e-4, atol=2e-4)
    assert torch.equal(b.cpu()[:, ~mask], rows[:, ~mask])
    x, y = torch.randn(12345), torch.randn(12345)
    assert abs(ops.ada_stats(x.cuda(), y.cuda()) - ref.ada_stats(x, y)) < 1e-5


def test_adam_amsgrad_rows_matches_torch_optim():
    R, P = 5, 1000
    p0 = torch.randn(R, P)
    ps = [torch.nn.Parameter(p0[r].clone()) for r in range(R)]
    opts = [torch.optim.Adam([q], lr=0.01, weight_decay=1e-3, amsgrad=True) for q in ps]
    p = p0.cuda()
    m, v, vm = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    steps = torch.zeros(R, dtype=torch.int32, device="cuda")
    for it in range(4):
        g = torch.randn(R, P)
        for r in range(R):
            ps[r].grad = g[r].clone()
            opts[r].step()
        ops.adam_amsgrad_rows_(p, g.cuda(), m, v, vm, steps, 0.01, 1e-3)
    want = torch.stack([q.detach() for q in ps])
    assert torch.allclose(p.cpu(), want, rtol=1e-4, atol=1e-6)
    assert int(steps[0]) == 4


def test_eval_reductions():
    logits, y = torch.randn(777, 90), torch.randint(0, 90, (777,))
    acc = ops.eval_logits(logits.cuda(), y.cuda()).cpu()
    r = ref.eval_logits(logits, y)
    assert acc[0] == r[0] and abs(acc[1] - r[1]) < 1e-2 * 777 * 1e-2 + 1e-1 and acc[2] == 777
    assert abs(float(ops.aue_sqerr(logits.cuda(), y.cuda())) - float(ref.aue_sqerr(logits, y))) < 1e-2
    preds, w = torch.randint(0, 5, (4, 300)), torch.rand(4)
    assert torch.equal(ops.ensemble_vote(preds.cuda(), w.cuda(), 5).cpu(), ref.ensemble_vote(preds, w, 5))
    pr, tg = torch.randint(0, 10, (5000,)), torch.randint(0, 10, (5000,))
    assert torch.equal(ops.confusion_matrix(pr.cuda(), tg.cuda(), 10).cpu(), ref.confusion_matrix(pr, tg, 10))


def test_gram_cosine_and_modp_and_misc():
    U = torch.randn(7, 50000)
    S, nrm = ops.gram_cosine(U.cuda())
    S2, nrm2 = ref.gram_cosine(U)
    assert torch.allclose(S.cpu(), S2, atol=1e-5) and torch.allclose(nrm.cpu(), nrm2, rtol=1e-5)
    for n, P in ((1, 33), (9, 70001), (16, 5000), (23, 12345), (32, 3001), (40, 999)):   # 1/3/6/10 block pairs + library fallback
        U = torch.randn(n, P)
        S, nrm = ops.gram_cosine(U.cuda())
        S2, nrm2 = ref.gram_cosine(U)
        assert torch.allclose(S.cpu(), S2, atol=2e-5), (n, P)
        assert torch.allclose(nrm.cpu(), nrm2, rtol=1e-5), (n, P)
    p = 2 ** 31 - 1
    A, B = torch.randint(0, p, (17, 33)), torch.randint(0, p, (33, 9))
    want = torch.tensor([[sum(int(A[i, k]) * int(B[k, j]) for k in range(33)) % p for j in range(9)] for i in range(17)])
    assert torch.equal(ops.modp_matmul(A.cuda(), B.cuda(), p).cpu(), want)
    s, t = torch.randn(64, 10, requires_grad=True), torch.randn(64, 10)
    l_ref = ref.kd_kl_loss(s, t, 3.0)
    l_ref.backward()
    sg = s.detach().cuda().requires_grad_(True)
    l_gpu = ops.kd_kl_loss(sg, t.cuda(), 3.0)
    l_gpu.backward()
    assert abs(float(l_gpu.detach()) - float(l_ref.detach())) < 1e-4 and torch.allclose(sg.grad.cpu(), s.grad, atol=1e-5)
    parts, y = torch.randn(3, 128, 1), torch.randint(0, 2, (128, 1)).float()
    l1, g1 = ref.vfl_bce_grad(parts, y)
    l2, g2 = ops.vfl_bce_grad(parts.cuda(), y.cuda())
    assert abs(float(l1) - float(l2)) < 1e-5 and torch.allclose(g2.cpu(), g1, atol=1e-6)
    x, w, b = torch.randn(4, 32, 8, 8), torch.randn(32), torch.randn(32)
    with torch.no_grad():
        assert torch.allclose(ops.group_norm(x.cuda(), 4, w.cuda(), b.cuda()).cpu(), F.group_norm(x, 4, w, b), atol=1e-4)


@pytest.mark.parametrize("mnk", [(128, 128, 64), (500, 1568, 784), (64, 10, 1568), (333, 128, 9216), (256, 512, 3136)])
def test_tcgen05_gemm(mnk):
    from feddrift_b200.ops import _ext
    M, N, K = mnk
    A = (torch.randn(M, K) * 0.5).bfloat16().cuda()
    B = (torch.randn(N, K) * 0.5).bfloat16().cuda()
    bias = torch.randn(N).cuda()
    D = _ext.load().gemm_tn_bias_act(A, B, bias, True, True)
    want = torch.relu(A.float() @ B.float().t() + bias)
    err = (D - want).abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("batch", [100, 104, 50])   # 104: dW on tcgen05 too (B % 8 == 0); 100 / 50: library dW
def test_tclinear_forward_backward(batch):
    from feddrift_b200.ops import linear
    from feddrift_b200.ops.linear import TcLinear
    torch.manual_seed(0)
    lin = TcLinear(784, 1568, activation="relu").cuda()   # K = 784 is not a multiple of the 64-wide k-block: TMA zero-fill
    x = torch.randn(batch, 784, device="cuda", requires_grad=True)
    before = linear.TC_CALLS
    y = lin(x)
    assert linear.TC_CALLS == before + 1, "TcLinear did not take the tcgen05 path"
    y.square().mean().backward()
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.relu(F.linear(xr, lin.weight.detach(), lin.bias.detach()))
    wr = lin.weight.detach().clone().requires_grad_(True)
    yr2 = torch.relu(F.linear(xr, wr, lin.bias.detach()))
    yr2.square().mean().backward()
    assert torch.allclose(y, yr, rtol=3e-2, atol=3e-2)
    assert torch.allclose(x.grad, xr.grad, rtol=5e-2, atol=2e-3)
    assert torch.allclose(lin.weight.grad, wr.grad, rtol=5e-2, atol=2e-3)


@pytest.mark.parametrize("geom", [(32, 64, 3, 1, 0, 28), (64, 128, 3, 2, 1, 17), (32, 64, 5, 1, 2, 14), (64, 64, 1, 1, 0, 8)])
def test_tcconv2d_forward_backward_matches_conv2d(geom):
    """im2col + tcgen05 GEMM conv vs F.conv2d in fp32 (bf16 operand tolerance), incl. stride 2 / padding / 1×1."""
    from feddrift_b200.ops import conv
    cin, cout, k, stride, pad, hw = geom
    torch.manual_seed(1)
    layer = conv.TcConv2d(cin, cout, k, stride=stride, padding=pad, activation="relu").cuda()
    x = torch.randn(10, cin, hw, hw, device="cuda", requires_grad=True)
    before = conv.TC_CONV_CALLS
    y = layer(x)
    assert conv.TC_CONV_CALLS == before + 1, "TcConv2d did not take the tcgen05 path"
    y.square().mean().backward()
    xr = x.detach().clone().requires_grad_(True)
    wr = layer.weight.detach().clone().requires_grad_(True)
    br = layer.bias.detach().clone().requires_grad_(True)
    yr = torch.relu(F.conv2d(xr, wr, br, stride, pad))
    yr.square().mean().backward()
    assert y.shape == yr.shape
    assert torch.allclose(y, yr, rtol=3e-2, atol=3e-2)
    scale = xr.grad.abs().max().item()
    assert (x.grad - xr.grad).abs().max().item() < 3e-2 * scale + 1e-5
    assert (layer.weight.grad - wr.grad).abs().max().item() < 3e-2 * wr.grad.abs().max().item() + 1e-5
    assert (layer.bias.grad - br.grad).abs().max().item() < 3e-2 * br.grad.abs().max().item() + 1e-5


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(128, 256, 64), (200, 320, 136), (504, 1568, 104), (64, 288, 28800)])
def test_tcgen05_gemm_operand_layouts(shape, a_mn, b_mn):
    """K-major and MN-major (transposed, row-contiguous) A / B operands of the tcgen05 GEMM vs an fp32 matmul."""
    from feddrift_b200.ops import _ext
    ext = _ext.load(required=True)
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    Ab, Bb = A.bfloat16(), B.bfloat16()
    want = Ab.float() @ Bb.float().t() + bias
    a_dev = (Ab.t().contiguous() if a_mn else Ab).cuda()
    b_dev = (Bb.t().contiguous() if b_mn else Bb).cuda()
    got = ext.gemm_bias_act(a_dev, b_dev, a_mn, b_mn, bias.cuda(), False, True).cpu()
    tol = 2e-3 * (K ** 0.5) + 1e-2
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < tol, (shape, a_mn, b_mn, (got - want).abs().max().item())


@pytest.mark.parametrize("shape,groups", [((4, 32, 8, 8), 8), ((3, 16, 5, 7), 2), ((2, 64, 4, 4, 4), 32)])
def test_group_norm_backward_kernel_matches_autograd(shape, groups):
    """K16 training path: fused forward (mean/rstd saved) + fused backward kernel vs F.group_norm autograd in fp32."""
    import torch.nn.functional as F
    from feddrift_b200 import ops
    torch.manual_seed(0)
    C = shape[1]
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    w = torch.randn(C, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    dy = torch.randn(*shape, device="cuda")
    ref = F.group_norm(x, groups, w, b, 1e-5)
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
    got = ops.group_norm(x, groups, w, b, 1e-5)
    hx, hw, hb = torch.autograd.grad(got, (x, w, b), dy)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)
    assert torch.allclose(hx, gx, atol=2e-4, rtol=1e-3)
    assert torch.allclose(hw, gw, atol=2e-3, rtol=1e-3)
    assert torch.allclose(hb, gb, atol=2e-3, rtol=1e-3)


def test_group_norm_module_trains_through_fused_kernels():
    from feddrift_b200.models.group_norm import GroupNorm2d
    m = GroupNorm2d(16, num_groups=4).cuda()
    x = torch.randn(2, 16, 6, 6, device="cuda", requires_grad=True)
    m(x).square().sum().backward()
    assert x.grad is not None and m.weight.grad is not None and torch.isfinite(x.grad).all()


@pytest.mark.parametrize("p", [2 ** 31 - 1, 2 ** 61 - 1, (1 << 62) + 135, 65537, 97, 1 << 20, (1 << 40) + 2])
def test_modp_matmul_montgomery_exact(p):
    """K13: Montgomery (odd p) / shift-subtract (even p) finite-field GEMM is bit-exact vs python big-int arithmetic."""
    g = torch.Generator().manual_seed(p % 1000)
    M, K, N = 37, 83, 21
    hi = min(p, 2 ** 62)
    A = torch.randint(0, hi, (M, K), generator=g, dtype=torch.int64)
    B = torch.randint(0, hi, (K, N), generator=g, dtype=torch.int64)
    A[0, :] = p - 1
    B[:, 0] = p - 1           # worst-case residues
    A[1, 0] = -5              # negative inputs are reduced into [0, p)
    want = [[sum((int(A[i, k]) % p) * (int(B[k, j]) % p) for k in range(K)) % p for j in range(N)] for i in range(M)]
    got = ops.modp_matmul(A.cuda(), B.cuda(), p).cpu()
    assert got.tolist() == want



@pytest.mark.parametrize("shape", [(4, 96, 5, 7), (32, 2048, 8, 8), (3, 4100, 1, 1)])
def test_batch_norm_nhwc_train_matches_torch(shape):
    """misc.cu::bn_nhwc_*: forward, running statistics and backward of training-mode BatchNorm on channels_last tensors."""
    from feddrift_b200 import ops
    torch.manual_seed(0)
    N, C, H, W = shape
    x = (torch.randn(N, C, H, W, device="cuda") * 2 + 0.5).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = torch.randn(C, device="cuda").requires_grad_(True)
    b = torch.randn(C, device="cuda").requires_grad_(True)
    rm0, rv0 = torch.randn(C, device="cuda"), torch.rand(C, device="cuda") + 0.5
    rm_ref, rv_ref, rm, rv = rm0.clone(), rv0.clone(), rm0.clone(), rv0.clone()
    ref = F.batch_norm(x, rm_ref, rv_ref, w, b, True, 0.1, 1e-5)
    dy = torch.randn_like(ref)
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
    w2 = w.detach().view(2, C // 2).clone().requires_grad_(True) if C % 2 == 0 else w.detach().clone().requires_grad_(True)
    got = ops.batch_norm_train_nhwc(x, w2, b, rm, rv, 1e-5, 0.1)
    hx, hw, hb = torch.autograd.grad(got, (x, w2, b), dy)
    tol = 2e-4
    assert (got - ref).abs().max().item() < tol * (1 + ref.abs().max().item())
    assert (rm - rm_ref).abs().max().item() < 1e-5 and (rv - rv_ref).abs().max().item() < 1e-4 * (1 + rv_ref.abs().max().item())
    assert (hx - gx).abs().max().item() < tol * (1 + gx.abs().max().item())
    assert (hw.reshape(-1) - gw).abs().max().item() < tol * (1 + gw.abs().max().item())
    assert (hb - gb).abs().max().item() < tol * (1 + gb.abs().max().item())
