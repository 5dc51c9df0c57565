"""MCA / MMA model classes with the reference's public surface (model.py:282-478), executed by hand-written sm_100a
kernels through mca_paper_b200.engine.Engine.

Drop-in contract: same constructor kwargs (extra kwargs such as `eao` are swallowed), same
dict-of-modality-tensors batch, same output dict keys in the same order (model.py:181-233,477), same state_dict
names/shapes (so the reference's checkpoints load), `outputs['loss']` differentiable.  MMA is MCA(zorro=True)
(utils/config.py:49).  Differences, all deliberate: no per-forward `torch.save` (model.py:94), no per-loss host
synchronisation (model.py:225), non-finite input detection through one device flag instead of 16 host syncs
(encoders.py:197-213) — the flag is read once per forward when `check_finite` is on.

There is no CPU / eager fallback: the modules hold parameters, every FLOP runs in libmca_b200.so.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from . import _lib
from .encoders import encoders_dict
from .engine import D, Engine
from .plan import FUSION_TOKEN, GLOBAL_TOKEN, EAOPlan, StaticPlan, fusion_channel_sets
from .utils.contrastive_loss_with_temperature import ContrastiveLossWithTemperature


def adjusted_powerset(unique_tokens, powers=(2, 3)):
    """model.py:11-12."""
    for c in fusion_channel_sets(len(list(unique_tokens)), powers):
        yield tuple(sorted(c))


MAX_BATCH_SIZE = 32  # LOSS_MAXB of csrc/loss.cu: local rows the all-pairs loss kernels keep in registers


def _check_batch_size(batch_size):
    if not 1 <= int(batch_size) <= MAX_BATCH_SIZE:
        raise ValueError(f"batch_size={batch_size}: the fused contrastive-loss kernels hold the local batch in registers "
                         f"and support 1..{MAX_BATCH_SIZE} samples per GPU (the shipped configs use 8 or 32); use more "
                         "GPUs (data parallel) for a larger global batch")


class LayerNorm(nn.Module):
    """gamma is learnable, beta is a constant zero buffer kept in the state_dict (model.py:24-31)."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))

    def forward(self, x):
        from .standalone import layer_norm

        return layer_norm(x, self.gamma, self.beta)


class GEGLU(nn.Module):
    """Keeps the `feedforward.{0,2}` state_dict indices (model.py:35-38).  Inside FeedForward / MCA it is fused into
    the FF1 GEMM epilogue; there is no stand-alone GEGLU kernel (the activation never exists as a separate pass)."""

    def forward(self, x):
        raise NotImplementedError("GEGLU is fused into the feed-forward GEMM epilogue (mca_gemm_bf16, MCA_EPI_GEGLU); "
                                  "call FeedForward, which runs Linear -> GEGLU -> Linear on the kernels")


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner_dim = int(dim * mult * 2 / 3)  # 1365 for dim 512 (model.py:46)
        self.inner_dim = inner_dim
        self.feedforward = nn.Sequential(nn.Linear(dim, inner_dim * 2, bias=False), GEGLU(),
                                         nn.Linear(inner_dim, dim, bias=False))

    def forward(self, batch):
        from .standalone import feed_forward

        return feed_forward(self, batch)


class Attention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner_dim = dim_head * heads
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)

    def forward(self, x, context=None, attn_mask=None, key_padding_mask=None, return_attn=False):
        """model.py:73-105 as a free-standing call (inside MCA.forward the engine drives the same kernels from the
        model's static schedule): standalone.attention derives the block-sparse schedule from `attn_mask`."""
        from .standalone import attention

        return attention(self, x, context=context, attn_mask=attn_mask, key_padding_mask=key_padding_mask,
                         return_attn=return_attn)


class MCALayer(nn.Module):
    """One shared LayerNorm applied before attention AND before the feed-forward, residual taken from the normed
    tensor (model.py:109-122)."""

    def __init__(self, dim, dim_head, heads, ff_mult):
        super().__init__()
        self.attn = Attention(dim=dim, dim_head=dim_head, heads=heads)
        self.num_heads = heads
        self.ff = FeedForward(dim=dim, mult=ff_mult)
        self.norm = LayerNorm(dim)

    def forward(self, batch, attn_mask=None, padding_mask=None):
        """model.py:117-122 as a free-standing call (quirk Q1: one shared norm, residuals from the normed tensor)."""
        batch = self.norm(batch)
        batch = self.attn(batch, attn_mask=attn_mask, key_padding_mask=padding_mask) + batch
        batch = self.norm(batch)
        batch = self.ff(batch) + batch
        return batch


class MCAPretrainingLoss(nn.Module):
    """Holder of the single shared ContrastiveLossWithTemperature (model.py:126-173); the pair list, names and mask
    rules live in StaticPlan._build_loss_plan and are evaluated by one kernel launch."""

    def __init__(self, modality_names, plan: StaticPlan):
        super().__init__()
        self.modality_names = modality_names
        self.loss_fn = ContrastiveLossWithTemperature()
        self.loss_names = list(plan.loss_names)


class _MCAFunction(torch.autograd.Function):
    """Whole trunk + loss as one autograd node: forward launches the fused forward, backward the fused backward and
    hands per-parameter gradients back to autograd."""

    @staticmethod
    def forward(ctx, model, batch, want_loss, *params):
        eng: Engine = model._engine
        pooled = eng.trunk_forward(batch)
        ctx.model = model
        ctx.want_loss = want_loss
        if want_loss:
            losses, summary = eng.loss_forward(pooled)
            return pooled.clone(), losses.clone(), summary.clone()
        return pooled.clone(), pooled.new_zeros(0), pooled.new_zeros(0)

    @staticmethod
    def backward(ctx, g_pooled, g_losses, g_summary):
        model = ctx.model
        eng: Engine = model._engine
        eng.flat_grad.zero_()
        dp = torch.zeros_like(eng.ws["pooled"]) if g_pooled is None else g_pooled.contiguous().float()
        if ctx.want_loss:
            ws = eng.ws
            # losses[p] feed: 'loss' (summary[0]), 'fcl_loss' ([1]), 'no-fcl_loss' ([2]) and the per-pair outputs
            w = torch.zeros_like(ws["w_default"]) if g_losses is None else g_losses.clone().float()
            if g_summary is not None:
                valid = ~torch.isnan(ws["losses"])
                fcl = torch.from_numpy(eng.plan.loss_is_fcl).to(w.device)
                w = w + g_summary[0] * ws["w_default"]
                nf, nn_ = int(fcl.sum()), int((~fcl).sum())
                if nf:
                    w = w + g_summary[1] * (valid & fcl).float() / nf
                if nn_:
                    w = w + g_summary[2] * (valid & ~fcl).float() / nn_
            w = torch.where(torch.isnan(ws["losses"]), torch.zeros_like(w), w).contiguous()
            dp = dp + eng.loss_backward(w)
        eng.trunk_backward(dp.contiguous())
        flat = eng.flat_grad.clone()
        grads = []
        for name, p in eng._param_list():
            o = eng.offs[name]
            grads.append(flat[o:o + p.numel()].view(p.shape))
        return (None, None, None, *grads)


class _FusedModel(nn.Module):
    """forward() shared by MCA / MMA and EAO: one autograd node around the fused engine, outputs named by the plan."""

    @property
    def engine(self) -> Engine:
        return self._engine

    def _named_outputs(self, pooled):
        return {key: pooled[:, row, :] for key, row in self.plan.output_rows}

    def forward(self, batch, no_loss=False):
        eng = self._engine
        eng.ensure_flat()
        # data parallel without the fused Trainer (the reference loop: model(batch); loss.backward() under DDP): the loss
        # gathers the pooled embeddings across the default process group like gather_tensor does (utils/distributed.py:
        # 23-56), through NCCL — the peer-memory exchange is set up by Trainer / Engine.set_distributed(p2p=True)
        if (eng.world == 1 and torch.distributed.is_available() and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1 and torch.distributed.get_backend() == "nccl"):
            eng.set_distributed(torch.distributed.get_world_size(), torch.distributed.get_rank(), None, p2p=False)
        eng.pack_weights()
        params = [p for _, p in eng._param_list()]
        want_loss = not no_loss
        pooled, losses, summary = _MCAFunction.apply(self, batch, want_loss, *params)
        if self.check_finite:
            flag = int(eng.ws["nonfinite"].item())
            if flag & 2:
                raise IndexError("index out of range in self")  # what nn.Embedding raises for a bad token / index
            if flag & 1:
                raise Exception("Tokens are not finite")  # encoders.py:197-198
        out = self._named_outputs(pooled)
        present = eng.ws["present"].clone().to(torch.bool)
        sample_mask = {name: present[:, i] for i, name in enumerate(self.modality_types)}
        if want_loss:
            out["losses"] = {name: losses[i] for i, name in enumerate(self.plan.loss_names)}
            if self.plan.do_fcl:
                out["fcl_loss"] = summary[1]
                out["no-fcl_loss"] = summary[2]
            out["loss"] = summary[0]
        out["modality_sample_mask"] = sample_mask
        return out


class MCA(_FusedModel):
    def __init__(self, encoder_configs, dim, depth, dim_head=64, heads=8, ff_mult=4, num_fusion_tokens=16,
                 batch_size=8, return_padding=False, return_logits=False, bimodal_contrastive=False,
                 non_fusion_fcl=False, fcl=False, fcl_root=(1, 2, 3, 4, 5), fusion_combos=(4, 5), zorro=False,
                 no_fusion=False, mean_pool=False, **kwargs):
        super().__init__()
        if mean_pool:
            raise NotImplementedError("mean_pool=True is broken for MCA in the reference (model.py:262) and not built")
        if dim != D:
            raise AssertionError("encoders hard-wire embedding_dim=512 (encoders.py:79,104,151,178,230): dim must be 512")
        encoder_configs = {k: dict(v) for k, v in dict(encoder_configs).items()}
        _check_batch_size(batch_size)
        self.batch_size = batch_size
        self.no_fusion = no_fusion
        self.fusion_token, self.global_token = FUSION_TOKEN, GLOBAL_TOKEN
        plan = StaticPlan(encoder_configs, num_fusion_tokens, list(fusion_combos), fcl, zorro, no_fusion,
                          bimodal_contrastive, non_fusion_fcl)
        self.plan = plan
        self.fusion_combos = plan.combos
        self.fcl_root = frozenset(fcl_root) if plan.do_fcl else None
        self.return_token_types = plan.return_token_types
        self.max_return_tokens = plan.R
        self.register_buffer("return_token_types_tensor", torch.tensor(plan.return_token_types), persistent=False)
        self.heads = heads
        self.return_padding, self.return_logits = return_padding, return_logits

        # same construction order as the reference (model.py:337-380) so that a seeded init draws the same stream
        self.encoders = nn.ModuleDict({name: encoders_dict[cfg["type"]](**cfg) for name, cfg in encoder_configs.items()})
        self.modality_types = list(encoder_configs.keys())
        self.encoder_specs = [dict(cfg) for cfg in encoder_configs.values()]
        self.num_fusion_tokens = plan.F
        self.token_dims = plan.lengths
        self.fusion_tokens = nn.Parameter(torch.randn(plan.F, dim))
        self.register_buffer("fusion_mask", torch.zeros(plan.F).to(torch.bool))
        self.layers = nn.ModuleList([MCALayer(dim, dim_head, heads, ff_mult) for _ in range(depth)])
        self.norm = LayerNorm(dim)
        self.register_buffer("token_types", torch.from_numpy(plan.token_types.copy()))
        self.return_tokens = nn.Parameter(torch.randn(plan.R, dim))
        self.attn_pool = Attention(dim=dim, dim_head=dim_head, heads=heads)
        self.register_buffer("attn_mask", torch.from_numpy(plan.attn_mask.copy()))
        self.register_buffer("pool_mask", torch.from_numpy(plan.pool_mask.copy()))
        self.loss = MCAPretrainingLoss(self.modality_types, plan)

        ff_inner = self.layers[0].ff.inner_dim if depth else int(dim * ff_mult * 2 / 3)
        self._engine = Engine(self, plan, depth, heads, ff_inner, batch_size)
        self.check_finite = True


class EAO(_FusedModel):
    """The "everything at once" baseline with the reference's constructor and state_dict (model.py:481-596): the same
    MCALayer stack is run once per modality and once per modality combination, each pass mean-pooled, the pooled tokens
    contrasted pairwise (train_accel_gpu.py:51-52 selects it with `eao: true`; every shipped *_EAO config has
    no_fusion=True, mean_pool=True).  All passes of a sample run as ONE packed sequence with a block-diagonal mask
    (plan.EAOPlan) through the same kernels as MCA; there are no fusion tokens, return tokens or pooling weights."""

    def __init__(self, encoder_configs, dim, depth, dim_head=64, heads=8, ff_mult=4, num_fusion_tokens=16,
                 batch_size=8, return_padding=False, return_logits=False, bimodal_contrastive=False,
                 non_fusion_fcl=False, fcl=False, fcl_root=(1, 2, 3, 4, 5), fusion_combos=(4, 5), zorro=False,
                 no_fusion=True, mean_pool=True, **kwargs):
        super().__init__()
        if not mean_pool:
            raise NotImplementedError("EAO(mean_pool=False) cannot run in the reference either: single_pass reads "
                                      "self.pool_mask, which EAO never defines (model.py:560)")
        if dim != D:
            raise AssertionError("encoders hard-wire embedding_dim=512 (encoders.py:79,104,151,178,230): dim must be 512")
        encoder_configs = {k: dict(v) for k, v in dict(encoder_configs).items()}
        _check_batch_size(batch_size)
        self.batch_size = batch_size
        plan = EAOPlan(encoder_configs, list(fusion_combos), fcl, zorro, no_fusion, bimodal_contrastive, non_fusion_fcl)
        self.plan = plan
        self.fusion_combos = plan.combos
        self.fcl_root = None                                   # model.py:506
        self.fusion_token = FUSION_TOKEN
        self.return_token_types = plan.return_token_types
        self.max_return_tokens = len(plan.return_token_types)
        self.register_buffer("return_token_types_tensor", torch.tensor(plan.return_token_types), persistent=False)
        self.heads = heads
        self.return_padding, self.return_logits = return_padding, return_logits
        # same construction order as the reference (model.py:520-565): a seeded init draws the same stream
        self.encoders = nn.ModuleDict({name: encoders_dict[cfg["type"]](**cfg) for name, cfg in encoder_configs.items()})
        self.modality_types = list(encoder_configs.keys())
        self.encoder_specs = [dict(cfg) for cfg in encoder_configs.values()]
        self.token_dims = plan.lengths
        self.layers = nn.ModuleList([MCALayer(dim, dim_head, heads, ff_mult) for _ in range(depth)])
        self.norm = LayerNorm(dim)
        self.register_buffer("token_types", torch.from_numpy(plan.token_types.copy()))
        self.return_tokens = None
        self.loss = MCAPretrainingLoss(self.modality_types, plan)
        ff_inner = self.layers[0].ff.inner_dim if depth else int(dim * ff_mult * 2 / 3)
        self._engine = Engine(self, plan, depth, heads, ff_inner, batch_size)
        self.check_finite = True
