"""Op surface of the framework.  Every function dispatches to the hand-written
sm_100a kernel when its tensors live on CUDA (the extension is then mandatory)
and to the fp32 PyTorch reference (``ops.reference``) on CPU.

Kernel map (SURVEY §2.9 numbering):
  K1  cluster_aggregate_ / weighted_average      csrc/aggregate.cu
  K3  fed_round_small (fused local step+K1+K4)    csrc/fed_round_small.cu
      adam_amsgrad_rows_ (arena optimizer)        csrc/optim.cu
      TcLinear GEMM (tcgen05/TMEM/TMA)            csrc/gemm_tc.cu
  K4  mlp_eval_matrix / eval_logits_              csrc/eval.cu
  K5  merge_axpby_ / cluster_distance             csrc/aggregate.cu / host (L ≤ #models)
  K6  gram_cosine                                 csrc/cluster_ops.cu
  K7  aue_sqerr / ensemble_vote / confusion       csrc/eval.cu
  K8  ada_stats (fused mean-square)               csrc/aggregate.cu
  K10 robust_clip_                                csrc/aggregate.cu
  K11 server_opt_step_                            csrc/aggregate.cu
  K12 gossip_mix                                  csrc/aggregate.cu
  K13 modp_matmul                                 csrc/mpc.cu
  K14 kd_kl_loss   K15 vfl_bce_grad   K16 group_norm   csrc/misc.cu
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _ext, reference as ref
from .reference import (batch_hash, cluster_distance, cohen_kappa, hash_choice, mix32, mlp_forward,  # noqa: F401
                        mlp_param_count, mlp_unpack)

KIND_ID = {"lr": 0, "fnn": 1}
OPT_ID = {"sgd": 0, "adam": 1}


def native(*tensors) -> bool:
    return _ext.use_native(*tensors)


# ----------------------------------------------------------------------------- K3 fused small round
def fed_round_small(st: Dict, rounds: int = 1) -> Dict[str, torch.Tensor]:
    """Run ``rounds`` complete FL rounds (broadcast → local steps → per-cluster aggregate →
    optional IFCA re-cluster → train/test evaluation of every client) for a small-MLP federation.
    See ``reference.fed_round_small`` for the exact semantics."""
    if native(st["theta"]):
        from .small_round import run_native
        return run_native(st, rounds)
    return ref.fed_round_small(st, rounds)


# ----------------------------------------------------------------------------- K1
def cluster_aggregate_(theta, client_params, n):
    if native(theta, client_params):
        return _ext.load().cluster_aggregate(theta, client_params.contiguous(), n.float().contiguous())
    return ref.cluster_aggregate_(theta, client_params, n)


def weighted_average(rows, weights, out=None):
    if native(rows):
        res = _ext.load().weighted_average(rows.contiguous(), weights.float().contiguous())
    else:
        res = ref.weighted_average(rows, weights)
    if out is not None:
        out.copy_(res)
        return out
    return res


def robust_clip_(rows, global_row, bound: float, weight_mask=None, stddev: float = 0.0, seed: int = 0):
    """K10: clip every row around ``global_row`` and (``stddev > 0``) add counter-hash Gaussian noise in the same pass."""
    if native(rows):
        mask = weight_mask.to(torch.uint8).contiguous() if weight_mask is not None else None
        return _ext.load().robust_clip(rows, global_row.contiguous(), float(bound), mask, float(stddev), int(seed) & 0xFFFFFFFF)
    return ref.robust_clip_(rows, global_row, bound, weight_mask, stddev, seed)


def server_opt_step_(theta, avg, state: Dict, opt: str, lr: float, **kw):
    if native(theta) and opt in ("sgd", "adam", "adagrad", "yogi"):
        from .server_opt import native_server_opt_step_
        return native_server_opt_step_(theta, avg, state, opt, lr, **kw)
    return ref.server_opt_step_(theta, avg, state, opt, lr, **kw)


def ada_stats(theta, prev_muh) -> float:
    if native(theta):
        return float(_ext.load().mean_sq_diff(theta.contiguous(), prev_muh.contiguous()))
    return ref.ada_stats(theta, prev_muh)


def gossip_mix(X, Wmix):
    if native(X):
        return _ext.load().gossip_mix(X.contiguous(), Wmix.float().contiguous())
    return ref.gossip_mix(X, Wmix)


def merge_axpby_(theta, base: int, second: int, w1: float, w2: float):
    if native(theta):
        _ext.load().merge_axpby(theta, int(base), int(second), float(w1), float(w2))
    else:
        ref.merge_axpby_(theta, base, second, w1, w2)


# ----------------------------------------------------------------------------- K4 / K7
def mlp_eval_matrix(theta, X, Y, nsamp, kind, din, hid, dout):
    if native(theta, X):
        out = _ext.load().mlp_eval_matrix(theta.contiguous(), X.contiguous(), Y.int().contiguous(),
                                          nsamp.int().contiguous(), KIND_ID[kind], din, hid, dout)
        return out[0], out[1]
    return ref.mlp_eval_matrix(theta, X, Y, nsamp, kind, din, hid, dout)


def eval_logits(logits, target, acc=None):
    """Accumulate (correct, loss_sum, count) into ``acc`` [3] on device without a host sync."""
    if native(logits):
        if acc is None:
            acc = torch.zeros(3, dtype=torch.float32, device=logits.device)
        _ext.load().eval_logits(logits.float().contiguous(), target.int().contiguous(), acc)
        return acc
    r = ref.eval_logits(logits, target)
    if acc is not None:
        acc += r
        return acc
    return r


def aue_sqerr(logits, target):
    if native(logits):
        return _ext.load().aue_sqerr(logits.float().contiguous(), target.int().contiguous())
    return ref.aue_sqerr(logits, target)


def ensemble_vote(preds, weights, num_classes: int):
    if native(preds):
        return _ext.load().ensemble_vote(preds.int().contiguous(), weights.float().contiguous(), int(num_classes))
    return ref.ensemble_vote(preds, weights, num_classes)


def soft_vote(probs, weights):
    return ref.soft_vote(probs, weights)


def confusion_matrix(pred, target, num_classes: int):
    if native(pred):
        return _ext.load().confusion_matrix(pred.int().contiguous(), target.int().contiguous(),
                                            int(num_classes)).double()
    return ref.confusion_matrix(pred, target, num_classes)


# ----------------------------------------------------------------------------- K6
def gram_cosine(U, eps: float = 1e-12):
    if native(U):
        out = _ext.load().gram_cosine(U.float().contiguous(), float(eps))
        return out[0], out[1]
    return ref.gram_cosine(U, eps)


# ----------------------------------------------------------------------------- arena optimizer
def adam_amsgrad_rows_(p, g, m, v, vmax, steps, lr: float, wd: float, b1=0.9, b2=0.999, eps=1e-8, row_mask=None):
    """Fused Adam(amsgrad, L2 wd) over arena rows [R,P]; ``steps`` [R] int32 is incremented in place."""
    if native(p):
        _ext.load().adam_amsgrad_rows(p, g.contiguous(), m, v, vmax, steps, float(lr), float(wd), float(b1),
                                      float(b2), float(eps), row_mask)
        return p
    for r in range(p.shape[0]):
        if row_mask is not None and not bool(row_mask[r]):
            continue
        steps[r] = ref.adam_amsgrad_update(p[r], g[r], m[r], v[r], vmax[r], int(steps[r]), lr, wd, b1, b2, eps)
    return p


def sgd_rows_(p, g, lr: float, wd: float = 0.0):
    if native(p):
        _ext.load().sgd_rows(p, g.contiguous(), float(lr), float(wd))
        return p
    p.add_(g + wd * p, alpha=-lr)
    return p


# ----------------------------------------------------------------------------- K13..K16
def modp_matmul(A, B, p: int):
    if native(A):
        return _ext.load().modp_matmul(A.long().contiguous(), B.long().contiguous(), int(p))
    return ref.modp_matmul(A, B, p)


class _KDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, t, T):
        out = _ext.load().kd_kl_fwd_bwd(s.float().contiguous(), t.float().contiguous(), float(T))
        ctx.save_for_backward(out[1])
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (gs,) = ctx.saved_tensors
        return gs * g, None, None


def kd_kl_loss(student_logits, teacher_logits, temperature: float = 1.0):
    if native(student_logits):
        return _KDLoss.apply(student_logits, teacher_logits.detach(), temperature)
    return ref.kd_kl_loss(student_logits, teacher_logits, temperature)


def vfl_bce_grad(logit_parts, y):
    if native(logit_parts):
        out = _ext.load().vfl_bce_grad(logit_parts.float().contiguous(), y.float().contiguous())
        return out[0], out[1]
    return ref.vfl_bce_grad(logit_parts, y)


class _GroupNormFn(torch.autograd.Function):
    """K16 with autograd: fused forward (saves per-group mean / rstd) + fused backward kernel (csrc/misc.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        xc = x.float().contiguous()
        y, mean, rstd = _ext.load().group_norm_fwd_train(xc, int(groups), weight, bias, float(eps))
        ctx.save_for_backward(xc, weight, mean, rstd)
        ctx.groups, ctx.has_bias = int(groups), bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, dg, db = _ext.load().group_norm_bwd(x, dy.float().contiguous(), weight, mean, rstd, ctx.groups)
        return dx, (dg if weight is not None else None), (db if ctx.has_bias else None), None, None


class _BatchNormNhwcFn(torch.autograd.Function):
    """Training-mode BatchNorm2d on a channels_last activation (csrc/misc.cu::bn_nhwc_*): 2 coalesced passes per direction, the
    running statistics are updated in place by the forward kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, run_mean, run_var, eps, momentum):
        xh = x.float().contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)       # NHWC view (free for channels_last)
        w = weight.detach().reshape(-1).contiguous() if weight is not None else None
        b = bias.detach().reshape(-1).contiguous() if bias is not None else None
        # the result is a fresh NCHW-logical channels_last tensor (NOT a view of a kernel output: ReLU(inplace=True) follows in
        # the torchvision blocks); the kernel writes through its NHWC view
        y = torch.empty_like(x, dtype=torch.float32, memory_format=torch.channels_last)
        mean, rstd = _ext.load().bn_nhwc_fwd(xh, y.permute(0, 2, 3, 1), w, b, run_mean, run_var, float(eps), float(momentum))
        ctx.save_for_backward(xh, w, mean, rstd)
        ctx.wshape = weight.shape if weight is not None else None
        ctx.has_bias = bias is not None
        ctx.bshape = bias.shape if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, gy):
        xh, w, mean, rstd = ctx.saved_tensors
        g = gy.float().contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
        dx, dw, db = _ext.load().bn_nhwc_bwd(xh, g, w, mean, rstd)
        return (dx.permute(0, 3, 1, 2), dw.view(ctx.wshape) if ctx.wshape is not None else None,
                db.view(ctx.bshape) if ctx.has_bias else None, None, None, None, None)


def batch_norm_train_nhwc(x, weight, bias, run_mean, run_var, eps: float, momentum: float):
    """Training-mode batch norm of a 4-D CUDA tensor through the NHWC kernels (weight / bias may be any shape with C elements)."""
    return _BatchNormNhwcFn.apply(x, weight, bias, run_mean, run_var, float(eps), float(momentum))


def group_norm(x, groups: int, weight=None, bias=None, eps: float = 1e-5):
    if native(x):
        if torch.is_grad_enabled() and (x.requires_grad or (weight is not None and weight.requires_grad)):
            return _GroupNormFn.apply(x, weight, bias, int(groups), float(eps))
        return _ext.load().group_norm_fwd(x.float().contiguous(), int(groups), weight, bias, float(eps))
    return ref.group_norm(x, groups, weight, bias, eps)
