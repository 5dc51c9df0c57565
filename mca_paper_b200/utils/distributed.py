"""Mirror of the reference's utils/distributed.py:11-63 API (BackpropType, gather_tensor, get_rank).

The fused loss path does ONE all-gather of the whole pooled block per step (engine.loss_forward) instead of two per
contrastive pair; these helpers keep the reference's module-level surface for callers that gather by hand.
"""
from enum import Enum
from typing import List

import torch
from torch import Tensor


class BackpropType(Enum):
    GLOBAL = 0
    LOCAL = 1
    NONE = 2


def get_rank() -> int:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


def gather_tensor(tensor: Tensor, backprop_type: BackpropType = BackpropType.GLOBAL) -> List[Tensor]:
    """utils/distributed.py:23-56: autograd-aware all-gather (GLOBAL), or a plain one (LOCAL/NONE)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return [tensor]
    world_size = torch.distributed.get_world_size()
    if world_size == 1:
        return [tensor]
    if backprop_type == BackpropType.GLOBAL:
        from torch.distributed.nn.functional import all_gather

        return list(all_gather(tensor))
    out = [torch.zeros_like(tensor) for _ in range(world_size)]
    torch.distributed.all_gather(out, tensor)
    if backprop_type == BackpropType.LOCAL:
        out[get_rank()] = tensor
    return out
