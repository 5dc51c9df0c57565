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
        if cross_entropy_kwargs:
            raise NotImplementedError("cross_entropy_kwargs are not supported by the fused InfoNCE kernel")
        if backprop_type != BackpropType.GLOBAL:
            raise NotImplementedError("only BackpropType.GLOBAL (the reference's default and only use) is implemented")
        lo = self.logit_scale_min if self.logit_scale_min is not None else -1e30
        hi = self.logit_scale_max if self.logit_scale_max is not None else 1e30
        return _PairLoss.apply(embeddings_a, embeddings_b, self.logit_scale, mask, lo, hi)
