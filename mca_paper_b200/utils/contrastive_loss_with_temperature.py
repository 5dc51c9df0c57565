"""ContrastiveLossWithTemperature with the reference's surface (utils/contrastive_loss_with_temperature.py:114-195,
runtime import model.py:6 from torchmultimodal), backed by the fused all-pairs InfoNCE kernels.

Inside MCA the module is only the holder of the shared `logit_scale` parameter (model.py:152-153: one instance for
every pair); MCA.forward evaluates all pairs in one launch.  Called on its own — `loss_fn(a, b, mask=...)` — it runs
the same kernels on a one-pair plan, with the all-gather of utils/contrastive_loss_with_temperature.py:21-37.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, Optional, Union

import numpy as np
import torch
from torch import Tensor, nn

from .. import ops
from ..ops import P, S, call
from .distributed import BackpropType, get_rank


@dataclass
class ContrastiveLossOutput:
    loss: Tensor
    logits_a: Tensor
    logits_b: Tensor
    loss_a: Tensor
    loss_b: Tensor


_PAIR_DTYPE = np.dtype([("a", "<i4"), ("b", "<i4"), ("all", "<u4"), ("any", "<u4"), ("fcl", "<i4")])


class _PairLoss(torch.autograd.Function):
    """(a, b, logit_scale) -> scalar loss through mca_contrastive_allpairs_{fwd,bwd} with R = 2."""

    @staticmethod
    def forward(ctx, a, b, logit_scale, mask, lo, hi):
        dev = a.device
        Bn, d = a.shape
        world, rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        pooled = torch.stack([a.float(), b.float()], dim=1).contiguous()
        if world > 1:
            pooled_all = torch.empty(world * Bn, 2, d, device=dev, dtype=torch.float32)
            torch.distributed.all_gather_into_tensor(pooled_all, pooled)
        else:
            pooled_all = pooled
        present = torch.ones(Bn, 1, device=dev, dtype=torch.uint8) if mask is None else mask.to(torch.uint8).view(Bn, 1).contiguous()
        plan = np.zeros(1, dtype=_PAIR_DTYPE)
        plan[0] = (0, 1, 1, 0, 0)
        plan_dev = torch.from_numpy(plan.view(np.uint8).copy()).to(dev)
        losses = torch.empty(1, device=dev)
        summary = torch.empty(4, device=dev)
        w = torch.empty(1, device=dev)
        call("mca_contrastive_allpairs_fwd", P(pooled_all), P(present), P(plan_dev), 1, P(logit_scale), Bn, world * Bn, 2,
             d, 1, rank, float(lo), float(hi), P(losses), P(summary), P(w), S())
        ctx.save_for_backward(pooled_all, present, plan_dev, logit_scale)
        ctx.meta = (Bn, d, world, rank)
        return losses[0].clone()

    @staticmethod
    def backward(ctx, g):
        pooled_all, present, plan_dev, logit_scale = ctx.saved_tensors
        Bn, d, world, rank = ctx.meta
        dev = pooled_all.device
        dall = torch.zeros_like(pooled_all)
        dscale = torch.zeros(1, device=dev)
        w = g.reshape(1).float().contiguous()
        call("mca_contrastive_allpairs_bwd", P(pooled_all), P(present), P(plan_dev), 1, P(logit_scale), Bn, world * Bn, 2,
             d, 1, rank, P(w), P(dall), P(dscale), S())
        if world > 1:
            dloc = torch.empty(Bn, 2, d, device=dev)
            torch.distributed.reduce_scatter_tensor(dloc, dall)
        else:
            dloc = dall
        return dloc[:, 0], dloc[:, 1], dscale.view(()), None, None, None


def _dist():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size(), torch.distributed.get_rank()
    return 1, 0


def _ce_options(cross_entropy_kwargs):
    """(label_smoothing, reduction) from the kwargs F.cross_entropy would receive (:94-100)."""
    kw = dict(cross_entropy_kwargs or {})
    eps = float(kw.pop("label_smoothing", 0.0))
    reduction = kw.pop("reduction", "mean")
    if reduction not in ("mean", "sum", "none"):
        raise ValueError(f"{reduction} is not a valid value for reduction")
    if not 0.0 <= eps <= 1.0:
        raise ValueError(f"label_smoothing must be between 0.0 and 1.0. Got: {eps}")
    for k in ("weight", "ignore_index", "size_average", "reduce"):
        v = kw.pop(k, None)
        if v is not None and not (k == "ignore_index" and v == -100):
            raise NotImplementedError(f"cross_entropy_kwargs[{k!r}] is not supported by the InfoNCE kernels "
                                      "(label_smoothing and reduction are)")
    if kw:
        raise TypeError(f"cross_entropy() got an unexpected keyword argument {next(iter(kw))!r}")
    return eps, reduction


class _GeneralLoss(torch.autograd.Function):
    """The functional form with its options: logits matrices as outputs, label_smoothing / reduction, and the three
    BackpropType modes of the all-gather (utils/distributed.py:23-56).  Kernels: mca_scaled_logits_f32,
    mca_cross_entropy_{fwd,bwd}, mca_small_gemm_f32 for the operand gradients."""

    @staticmethod
    def forward(ctx, a, b, logit_scale, mask, backprop, eps, reduction):
        dev = a.device
        Bn, d = a.shape
        world, rank = _dist()
        a32, b32 = a.detach().float().contiguous(), b.detach().float().contiguous()
        if world > 1:
            a_all = torch.empty(world * Bn, d, device=dev, dtype=torch.float32)
            b_all = torch.empty(world * Bn, d, device=dev, dtype=torch.float32)
            torch.distributed.all_gather_into_tensor(a_all, a32)
            torch.distributed.all_gather_into_tensor(b_all, b32)
        else:
            a_all, b_all = a32, b32
        GB = world * Bn
        s32 = logit_scale.detach().float().reshape(1).contiguous()
        la = torch.empty(Bn, GB, device=dev, dtype=torch.float32)
        lb = torch.empty(Bn, GB, device=dev, dtype=torch.float32)
        call("mca_scaled_logits_f32", P(a32), P(b_all), P(s32), Bn, GB, d, P(la), S())
        call("mca_scaled_logits_f32", P(b32), P(a_all), P(s32), Bn, GB, d, P(lb), S())
        labels = Bn * rank + torch.arange(Bn, device=dev, dtype=torch.int64)
        sel = None
        if mask is not None:
            sel = mask.to(device=dev, dtype=torch.bool).nonzero(as_tuple=True)[0]
            la, lb, labels = la.index_select(0, sel).contiguous(), lb.index_select(0, sel).contiguous(), labels.index_select(0, sel)
        n = int(la.shape[0])
        rl = torch.empty(2, max(n, 1), device=dev, dtype=torch.float32)
        lse = torch.empty(2, max(n, 1), device=dev, dtype=torch.float32)
        call("mca_cross_entropy_fwd", P(la), GB, P(labels), n, GB, eps, P(rl[0]), P(lse[0]), S())
        call("mca_cross_entropy_fwd", P(lb), GB, P(labels), n, GB, eps, P(rl[1]), P(lse[1]), S())
        rl = rl[:, :n]
        if reduction == "none":
            loss_a, loss_b = rl[0].clone(), rl[1].clone()
        elif reduction == "sum":
            loss_a, loss_b = rl[0].sum(), rl[1].sum()
        else:  # mean over the selected rows; an empty selection is NaN like F.cross_entropy on an empty batch
            loss_a, loss_b = rl[0].mean(), rl[1].mean()
        loss = (loss_a + loss_b) / 2
        ctx.save_for_backward(a32, b32, a_all, b_all, s32, la, lb, labels, lse, sel if sel is not None else labels)
        ctx.meta = (Bn, d, world, rank, n, eps, reduction, backprop, sel is not None, a.dtype, b.dtype)
        ctx.mark_non_differentiable(la, lb)
        return loss, la, lb, loss_a, loss_b

    @staticmethod
    def backward(ctx, g_loss, g_la, g_lb, g_loss_a, g_loss_b):
        a32, b32, a_all, b_all, s32, la, lb, labels, lse, sel = ctx.saved_tensors
        Bn, d, world, rank, n, eps, reduction, backprop, has_sel, adt, bdt = ctx.meta
        dev = a32.device
        GB = world * Bn
        da = torch.zeros(Bn, d, device=dev, dtype=torch.float32)
        db = torch.zeros(Bn, d, device=dev, dtype=torch.float32)
        dscale = torch.zeros(1, device=dev, dtype=torch.float32)
        if n > 0:
            def row_grad(g_dir):
                g = g_loss / 2 if g_loss is not None else 0.0
                if g_dir is not None:
                    g = g + g_dir
                if reduction == "mean":
                    g = g / n
                return (torch.zeros(n, device=dev) + g).float().contiguous()   # scalar or per-row ('none') -> [n]

            da_all = torch.zeros(GB, d, device=dev, dtype=torch.float32)
            db_all = torch.zeros(GB, d, device=dev, dtype=torch.float32)
            a_sel = a32.index_select(0, sel) if has_sel else a32
            b_sel = b32.index_select(0, sel) if has_sel else b32
            dl = torch.empty(n, GB, device=dev, dtype=torch.float32)
            dq = torch.empty(n, d, device=dev, dtype=torch.float32)
            for (z, zl, g_dir, q_sel, k_all, dq_out, dk_all) in ((la, lse[0], g_loss_a, a_sel, b_all, da, db_all),
                                                                 (lb, lse[1], g_loss_b, b_sel, a_all, db, da_all)):
                g_row = row_grad(g_dir)
                call("mca_cross_entropy_bwd", P(z), GB, P(labels), n, GB, eps, P(zl), P(g_row), P(s32), P(dl), P(dscale), S())
                # d q[sel] = dlogits k_all ; d k_all += dlogits^T q[sel]   (dlogits already carries the temperature)
                ops.small_gemm(dl, GB, 1, k_all, 1, d, dq, d, n, d, GB)
                if has_sel:
                    dq_out.index_add_(0, sel, dq)
                else:
                    dq_out.add_(dq)
                ops.small_gemm(dl, 1, GB, q_sel, 1, d, dk_all, d, GB, d, n, accumulate=True)
            # gradient of the gathered copies (utils/distributed.py:43-56)
            if world == 1:
                da.add_(da_all)
                db.add_(db_all)
            elif backprop == BackpropType.GLOBAL:
                ra, rb = torch.empty_like(da), torch.empty_like(db)
                torch.distributed.reduce_scatter_tensor(ra, da_all)
                torch.distributed.reduce_scatter_tensor(rb, db_all)
                da.add_(ra)
                db.add_(rb)
            elif backprop == BackpropType.LOCAL:
                da.add_(da_all[rank * Bn:(rank + 1) * Bn])
                db.add_(db_all[rank * Bn:(rank + 1) * Bn])
            # BackpropType.NONE: the gathered copies carry no gradient
        return da.to(adt), db.to(bdt), dscale.view(()), None, None, None, None


def contrastive_loss_with_temperature(embeddings_a: Tensor, embeddings_b: Tensor, logit_scale: nn.Parameter,
                                      mask: Optional[Tensor] = None, backprop_type: BackpropType = BackpropType.GLOBAL,
                                      cross_entropy_kwargs: Optional[Dict[str, Any]] = None) -> ContrastiveLossOutput:
    """Functional form (utils/contrastive_loss_with_temperature.py:40-108): temperature exp(logit_scale) applied as
    given (no clamp — the module clamps before calling), logits of both directions returned with the losses.  The
    logits are outputs only (not differentiable); gradients flow through `loss`, `loss_a` and `loss_b`."""
    if not embeddings_a.is_cuda:
        from .. import _lib
        raise _lib.MCAKernelError("contrastive_loss_with_temperature runs on CUDA only (no CPU fallback)")
    eps, reduction = _ce_options(cross_entropy_kwargs)
    loss, la, lb, loss_a, loss_b = _GeneralLoss.apply(embeddings_a, embeddings_b, logit_scale, mask, backprop_type, eps,
                                                      reduction)
    return ContrastiveLossOutput(loss=loss, logits_a=la, logits_b=lb, loss_a=loss_a, loss_b=loss_b)


class ContrastiveLossWithTemperature(nn.Module):
    def __init__(self, logit_scale: Union[float, nn.Parameter] = math.log(1 / 0.07),
                 logit_scale_min: Optional[float] = math.log(1), logit_scale_max: Optional[float] = math.log(100)):
        super().__init__()
        if not logit_scale_min and not logit_scale_max:
            raise ValueError("Only one of `logit_scale_min` and `logit_scale_max` can be None.")
        self.logit_scale_min = logit_scale_min
        self.logit_scale_max = logit_scale_max
        if isinstance(logit_scale, nn.Parameter):
            self.logit_scale = logit_scale
        else:
            self.logit_scale = nn.Parameter(logit_scale * torch.ones([]))

    def forward(self, embeddings_a: Tensor, embeddings_b: Tensor, backprop_type: BackpropType = BackpropType.GLOBAL,
                cross_entropy_kwargs: Optional[Dict[str, Any]] = None, mask: Optional[Tensor] = None) -> Tensor:
        """:178-195.  Default options (what MCAPretrainingLoss uses): the fused one-pair form of the all-pairs kernels;
        cross_entropy_kwargs or BackpropType.LOCAL / NONE: the general form above after the in-place clamp of :187."""
        if not cross_entropy_kwargs and backprop_type == BackpropType.GLOBAL:
            lo = self.logit_scale_min if self.logit_scale_min is not None else -1e30
            hi = self.logit_scale_max if self.logit_scale_max is not None else 1e30
            return _PairLoss.apply(embeddings_a, embeddings_b, self.logit_scale, mask, lo, hi)
        self.logit_scale.data.clamp_(self.logit_scale_min, self.logit_scale_max)
        return contrastive_loss_with_temperature(embeddings_a, embeddings_b, self.logit_scale, mask=mask,
                                                 backprop_type=backprop_type,
                                                 cross_entropy_kwargs=cross_entropy_kwargs).loss
