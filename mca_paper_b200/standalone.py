"""Free-standing calls of the reference's sub-modules: `LayerNorm(x)` model.py:30-31,
`FeedForward(x)` :49-54, `Attention(x, context, attn_mask, key_padding_mask, return_attn)` :73-105 and the encoders'
`forward(batch) -> (tokens, attention_mask)` encoders.py:90-96,114-120,161-166,196-214,268-274.

Inside `MCA.forward` none of this runs (the fused engine sequences the same kernels over the packed token buffer);
these wrappers exist so that every public module of the drop-in surface can be called on its own, with autograd, on the
same C-ABI entry points: LayerNorm -> mca_layernorm512_{fwd,bwd}; FeedForward -> mca_gemm_bf16 with the GEGLU
epilogues; Attention -> QKV / out projections on mca_gemm_bf16 around mca_attn_{fwd,bwd}, whose block-sparse schedule
is derived from the dense boolean mask of the call (plan.MaskPlan: keys with identical mask columns form a key group);
encoders -> an encoders-only Engine (engine.Engine(trunk=False)) that runs Engine.encode / encode_backward for one
modality.  No CPU or eager fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import List, Optional

import numpy as np
import torch

from . import _lib, ops
from .ops import P, S, call

D = 512
DH = 64
ALIGN = 64


def _round_up(x, m):
    return (x + m - 1) // m * m


def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.MCAKernelError(f"{what}: mca_paper_b200 runs on CUDA only (no CPU fallback); got a {t.device} tensor")


def _rows512(x, what):
    if x.shape[-1] != D:
        raise ValueError(f"{what}: the kernels are built for a last dimension of {D} (encoders.py hard-wires "
                         f"embedding_dim=512), got {x.shape[-1]}")
    x2 = x.reshape(-1, D)
    if x2.dtype != torch.float32 or not x2.is_contiguous():
        x2 = x2.to(torch.float32).contiguous()
    return x2


def _bf16(x32):
    """fp32 [rows, cols] -> bf16 GEMM operand (mca_cast_f32_bf16)."""
    rows, cols = x32.shape
    out = torch.empty(rows, cols, device=x32.device, dtype=torch.bfloat16)
    call("mca_cast_f32_bf16", P(x32), cols, P(out), cols, rows, cols, S())
    return out


# ------------------------------------------------------------------------------------------------ weight packs
class _Pack:
    """bf16 kernel-layout copies of a module's weight matrices and the fp32 slabs their gradients are reduced into:
    the small-scale version of Engine._build_pack_descs (same mca_pack_weights / mca_unpack_grads descriptors)."""

    def __init__(self, dev, params: List[torch.Tensor], mats):
        """params: the module's fp32 parameters in a fixed order; mats: [(key, (rows, cols) kernel-layout shape,
        [(param index, prow, pcol, row0, mode, half, scale)])]."""
        self.offs, total = [], 0
        for p in params:
            self.offs.append(total)
            total += _round_up(p.numel(), ALIGN)
        self.n_flat = total
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.w, rows, a_off = {}, [], 0
        for key, (r, c), parts in mats:
            self.w[key] = (a_off, r, c)
            for (pi, prow, pcol, row0, mode, half, scale) in parts:
                rows.append((self.offs[pi], a_off, prow, pcol, c, row0, mode, half, scale, r * c))
            a_off += _round_up(r * c, 512)
        self.arena = torch.zeros(a_off, device=dev, dtype=torch.bfloat16)
        self.garena = torch.zeros(a_off, device=dev, dtype=torch.float32)
        pd = np.zeros(len(rows), dtype=ops.PACK_DESC_DTYPE)
        ud = np.zeros(len(rows), dtype=ops.PACK_DESC_DTYPE)
        for i, (src, aoff, prow, pcol, ld, row0, mode, half, scale, slab) in enumerate(rows):
            pd[i] = (src, aoff, prow, pcol, ld, row0, mode, half, scale, 1, 0)
            ud[i] = (src, aoff, prow, pcol, ld, row0, mode, half, scale, 1, slab)
        self.n_desc = len(rows)
        self.pack_descs = torch.from_numpy(pd.view(np.uint8).copy()).to(dev)
        self.unpack_descs = torch.from_numpy(ud.view(np.uint8).copy()).to(dev)

    def load(self, params):
        for o, p in zip(self.offs, params):
            self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
        call("mca_pack_weights", P(self.flat), P(self.arena), P(self.pack_descs), self.n_desc, S())

    def W(self, key):
        o, r, c = self.w[key]
        return self.arena[o:o + r * c].view(r, c)

    def GW(self, key):
        o, r, c = self.w[key]
        return self.garena[o:o + r * c].view(1, r, c)

    def dw(self, key, dY, X, tokens):
        """slab[key] += dY^T X over `tokens` rows (both operands MN-major, split-K reduce-add)."""
        _, r, c = self.w[key]
        tiles = (r // 128 if r >= 128 else 1) * ((c + 127) // 128)
        splits = ops.effective_splits(tokens, max(1, min(16, 296 // max(1, tiles))))
        ops.gemm(dY, 1, X, 1, r, c, tokens, _lib.EPI_F32_ACC, self.GW(key), ld0=c, k_splits=splits)

    def grads(self, params):
        """state_dict-layout gradients of `params` from the kernel-layout slabs."""
        self.flat_grad.zero_()
        call("mca_unpack_grads", P(self.flat_grad), P(self.garena), P(self.unpack_descs), self.n_desc, S())
        return [self.flat_grad[o:o + p.numel()].view(p.shape).clone() for o, p in zip(self.offs, params)]


_PACKS = weakref.WeakKeyDictionary()   # module -> {device: _Pack}


def _pack_for(module, dev, params, mats_fn):
    per = _PACKS.setdefault(module, {})
    pk = per.get(dev)
    if pk is None:
        pk = per[dev] = _Pack(dev, params, mats_fn())
    return pk


# ------------------------------------------------------------------------------------------------ LayerNorm
class _LayerNorm512(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, gamma, beta):
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        stats = torch.empty(rows, 2, device=x2.device, dtype=torch.float32)
        ops.layernorm512_fwd(x2, gamma, beta, y, None, stats, rows)
        ctx.save_for_backward(x2, gamma, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, gamma, stats = ctx.saved_tensors
        rows = x2.shape[0]
        dy = dy.contiguous().float()
        dx = torch.empty_like(x2)
        dx16 = torch.empty(rows, D, device=x2.device, dtype=torch.bfloat16)
        dgamma = torch.zeros(D, device=x2.device, dtype=torch.float32)
        dbeta = torch.zeros(D, device=x2.device, dtype=torch.float32)
        ops.layernorm512_bwd(dy, x2, stats, gamma, dx, dx16, dgamma, dbeta, rows)
        return dx, dgamma, dbeta


def layer_norm(x, gamma, beta):
    """F.layer_norm(x, x.shape[-1:], gamma, beta) (model.py:30-31) on mca_layernorm512_{fwd,bwd}."""
    _need_cuda(x, "LayerNorm")
    x2 = _rows512(x, "LayerNorm")
    g = gamma if gamma.dtype == torch.float32 else gamma.float()
    b = beta if beta.dtype == torch.float32 else beta.float()
    return _LayerNorm512.apply(x2, g.contiguous(), b.contiguous()).view(x.shape)


# ------------------------------------------------------------------------------------------------ FeedForward
def _ff_mats(I, IP):
    return lambda: [("ff1", (2 * IP, D), [(0, 2 * I, D, 0, 1, I, 1.0)]),     # GEGLU rows interleaved 64 | 64
                    ("ff2", (D, IP), [(1, D, I, 0, 0, 0, 1.0)])]


class _FeedForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x2, w1, w2):
        dev, M = x2.device, x2.shape[0]
        I = module.inner_dim
        IP = _round_up(I, 128)
        pk = _pack_for(module, dev, [w1, w2], _ff_mats(I, IP))
        pk.load([w1, w2])
        x16 = _bf16(x2)
        u = torch.empty(M, 2 * IP, device=dev, dtype=torch.bfloat16)
        h = torch.empty(M, IP, device=dev, dtype=torch.bfloat16)
        y = torch.empty(M, D, device=dev, dtype=torch.float32)
        ops.gemm(x16, 0, pk.W("ff1"), 0, M, 2 * IP, D, _lib.EPI_GEGLU, h, ld0=IP, out1=u, ld1=2 * IP)
        ops.gemm(h, 0, pk.W("ff2"), 0, M, D, IP, _lib.EPI_F32, y)
        ctx.module, ctx.pk, ctx.IP = module, pk, IP
        ctx.save_for_backward(x16, u, h, w1, w2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x16, u, h, w1, w2 = ctx.saved_tensors
        pk, IP = ctx.pk, ctx.IP
        dev, M = x16.device, x16.shape[0]
        pk.load([w1, w2])   # the module may have been called again (other weights) since the forward
        pk.garena.zero_()
        d16 = _bf16(dy.contiguous().float())
        du = torch.empty(M, 2 * IP, device=dev, dtype=torch.bfloat16)
        dx = torch.empty(M, D, device=dev, dtype=torch.float32)
        ops.gemm(d16, 0, pk.W("ff2"), 1, M, IP, D, _lib.EPI_GEGLU_BWD, du, ld0=2 * IP, aux0=u, ldaux=2 * IP)
        pk.dw("ff2", d16, h, M)
        ops.gemm(du, 0, pk.W("ff1"), 1, M, D, 2 * IP, _lib.EPI_F32, dx)
        pk.dw("ff1", du, x16, M)
        g1, g2 = pk.grads([w1, w2])
        return None, dx, g1, g2


def feed_forward(module, x):
    """Linear(512, 2I) -> GEGLU -> Linear(I, 512) (model.py:35-54), both matmuls on tcgen05 with the GEGLU epilogues."""
    _need_cuda(x, "FeedForward")
    x2 = _rows512(x, "FeedForward")
    w1, w2 = module.feedforward[0].weight, module.feedforward[2].weight
    return _FeedForward.apply(module, x2, w1, w2).view(x.shape)


# ------------------------------------------------------------------------------------------------ Attention
_PLANS = {}


def _mask_plan(attn_mask_np, n):
    from .plan import MaskPlan
    key = (n, None if attn_mask_np is None else attn_mask_np.tobytes())
    pl = _PLANS.get(key)
    if pl is None:
        if len(_PLANS) > 16:
            _PLANS.clear()
        pl = _PLANS[key] = MaskPlan(attn_mask_np, n)
    return pl


class _AttnTables:
    """Device copies of a MaskPlan's schedule + the per-call offsets workspace for batch size B."""

    def __init__(self, pl, B, dev):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.pl, self.B, self.N = pl, B, pl.N
        self.keygrp, self.rowbits = t(pl.keygrp), t(pl.rowbits.view(np.int32))
        self.q_tiles, self.tile_grp, self.kt_list = t(pl.q_tiles_sorted), t(pl.tile_grp), t(pl.kt_list)
        self.k_tiles, self.k_tiles_q, self.qt_list = t(pl.tiles), t(pl.k_tiles_q), t(pl.qt_list)
        self.kt_start, self.kt_len = t(pl.tiles[:, 0].copy()), t(pl.tiles[:, 1].copy())
        self.n_kt = int(pl.tiles.shape[0])
        N = pl.N
        u8 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.uint8)
        i32 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.int32)
        self.padding, self.pad_mod, self.present, self.live_count = u8(B, N), u8(B * N), u8(B, 1), i32(B, 1)
        self.live_idx, self.cu_live = i32(B, N), i32(B + 1)
        self.kt_class, self.kt_live, self.any_absent = u8(B, self.n_kt), i32(B, self.n_kt, 4), i32(1)
        self.vmean = torch.zeros(B, D, device=dev, dtype=torch.float32)

    def build_offsets(self, key_padding):
        """key_padding [B, N] uint8 (non-zero = padded key) -> padding / tile classes / live-key words."""
        ptrs = (ctypes.c_void_p * 1)(key_padding.data_ptr())
        es = (ctypes.c_int * 1)(1)
        lens = (ctypes.c_int * 1)(self.N)
        call("mca_build_offsets", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(es, ctypes.c_void_p),
             ctypes.cast(lens, ctypes.c_void_p), 1, self.B, self.N, P(self.kt_start), P(self.kt_len), self.n_kt,
             P(self.padding), P(self.pad_mod), P(self.present), P(self.live_count), P(self.live_idx), P(self.cu_live),
             P(self.kt_class), P(self.kt_live), P(self.any_absent), S())
        # a query row may be fully masked without a whole block being absent (arbitrary masks): always form mean(V)
        self.any_absent.fill_(1)


def _attn_mats():
    return [("qkv", (3 * D, D), [(0, D, D, 0, 0, 0, DH ** -0.5), (1, 2 * D, D, D, 0, 0, 1.0)]),
            ("out", (D, D), [(2, D, D, 0, 0, 0, 1.0)])]


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x2, tab, want_probs, wq, wkv, wo):
        dev, M = x2.device, x2.shape[0]
        B, N, H = tab.B, tab.N, module.heads
        pk = _pack_for(module, dev, [wq, wkv, wo], _attn_mats)
        pk.load([wq, wkv, wo])
        x16 = _bf16(x2)
        qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
        ao = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, N, device=dev, dtype=torch.float32)
        y = torch.empty(M, D, device=dev, dtype=torch.float32)
        ops.gemm(x16, 0, pk.W("qkv"), 0, M, 3 * D, D, _lib.EPI_BF16, qkv)
        call("mca_attn_fwd", P(qkv), P(tab.q_tiles), int(tab.q_tiles.shape[0]), P(tab.kt_list), P(tab.k_tiles), tab.n_kt,
             P(tab.rowbits), P(tab.keygrp), P(tab.tile_grp), P(tab.kt_class), P(tab.kt_live), None, None, P(tab.any_absent),
             P(tab.vmean), P(ao), P(lse), B, N, H, S())
        ops.gemm(ao, 0, pk.W("out"), 0, M, D, D, _lib.EPI_F32, y)
        probs = None
        if want_probs:
            probs = torch.empty(B, H, N, N, device=dev, dtype=torch.float32)
            call("mca_attn_probs", P(qkv), P(lse), P(tab.rowbits), P(tab.keygrp), P(tab.padding), B, N, H, P(probs), S())
            ctx.mark_non_differentiable(probs)
        ctx.module, ctx.pk, ctx.tab = module, pk, tab
        # the offsets workspace of `tab` is shared by later calls with the same mask: keep this call's copies
        ctx.save_for_backward(x16, qkv, ao, lse, tab.padding.clone(), tab.kt_class.clone(), wq, wkv, wo)
        return (y, probs) if want_probs else y

    @staticmethod
    def backward(ctx, dy, *unused):
        x16, qkv, ao, lse, padding, kt_class, wq, wkv, wo = ctx.saved_tensors
        pk, tab, module = ctx.pk, ctx.tab, ctx.module
        dev, M = x16.device, x16.shape[0]
        B, N, H = tab.B, tab.N, module.heads
        pk.load([wq, wkv, wo])
        pk.garena.zero_()
        d16 = _bf16(dy.contiguous().float())
        dattn = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        ops.gemm(d16, 0, pk.W("out"), 1, M, D, D, _lib.EPI_BF16, dattn)
        pk.dw("out", d16, ao, M)
        delta = torch.empty(B, H, N, device=dev, dtype=torch.float32)
        ucorr = torch.empty(B, D, device=dev, dtype=torch.float32)
        dq_acc = torch.empty(M, D, device=dev, dtype=torch.float32)
        dqkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
        call("mca_attn_bwd", P(qkv), P(ao), P(dattn), P(lse), P(tab.k_tiles_q), tab.n_kt, P(tab.qt_list), P(tab.k_tiles),
             int(tab.q_tiles.shape[0]), P(tab.rowbits), P(tab.keygrp), P(tab.tile_grp), P(padding), P(kt_class),
             None, P(delta), P(ucorr), P(dq_acc), P(dqkv), B, N, H, S())
        dx = torch.empty(M, D, device=dev, dtype=torch.float32)
        ops.gemm(dqkv, 0, pk.W("qkv"), 1, M, D, 3 * D, _lib.EPI_F32, dx)
        pk.dw("qkv", dqkv, x16, M)
        gq, gkv, go = pk.grads([wq, wkv, wo])
        return None, dx, None, None, gq, gkv, go


_TABLES = {}


def attention(module, x, context=None, attn_mask=None, key_padding_mask=None, return_attn=False):
    """Attention.forward of model.py:73-105 for an arbitrary boolean `attn_mask` [i, j] / [b?, ...] broadcastable over
    batch and heads (True = may not attend) and `key_padding_mask` [b, j] (True = padded key), dim_head 64, 8 heads.
    Cross-attention (`context`) runs as self-attention over the concatenation [x; context] under a mask that lets the
    rows of x see only the context keys they are allowed (the context rows are computed and dropped)."""
    _need_cuda(x, "Attention")
    if module.heads * DH != D or module.to_q.weight.shape != (D, D):
        raise ValueError("Attention kernels are built for dim = 512 = 8 heads x 64")
    if x.dim() != 3:
        raise ValueError("Attention expects x of shape [batch, tokens, 512]")
    B, Nq = x.shape[0], x.shape[1]
    dev = x.device
    am = None
    if attn_mask is not None:
        am = attn_mask
        while am.dim() > 2:   # the reference broadcasts the mask over batch and heads (model.py:91)
            if am.shape[0] != 1:
                raise ValueError("a per-sample / per-head attn_mask is not supported: pass the [i, j] mask "
                                 "(per-sample key masking goes through key_padding_mask)")
            am = am[0]
        am = am.detach().to(torch.bool).cpu().numpy()
    if context is None:
        seq, Nk, N = x, Nq, Nq
        full = am
        kp = key_padding_mask
    else:
        Nk = context.shape[1]
        N = Nq + Nk
        seq = torch.cat([x, context.to(x.dtype)], dim=1)
        full = np.ones((N, N), dtype=bool)
        full[:Nq, Nq:] = False if am is None else am
        kp = torch.ones(B, N, device=dev, dtype=torch.bool)     # the rows of x are never keys
        kp[:, Nq:] = False if key_padding_mask is None else key_padding_mask.to(torch.bool)
        # A query row without any visible key is uniform over the CONTEXT keys in the reference (-finfo.max fill);
        # under the concatenation it would average over the rows of x as well: refuse instead of returning that.
        vis = ~torch.from_numpy(full[:Nq, Nq:]).to(dev)
        live = ~kp[:, Nq:]
        if bool(((vis[None].float() @ live.float()[:, :, None]) == 0).any()):
            raise NotImplementedError("cross-attention with a query row that sees no live key (reference: uniform over "
                                      "all context keys) is only built on the fused pooling path of MCA.forward")
    if full is not None and full.shape != (N, N):
        raise ValueError(f"attn_mask shape {full.shape} does not match the {N} tokens")
    pl = _mask_plan(full, N)
    tkey = (id(pl), B, str(dev))
    tab = _TABLES.get(tkey)
    if tab is None:
        if len(_TABLES) > 16:
            _TABLES.clear()
        tab = _TABLES[tkey] = _AttnTables(pl, B, dev)
    kp8 = (torch.zeros(B, N, device=dev, dtype=torch.uint8) if kp is None
           else kp.to(device=dev, dtype=torch.uint8).contiguous())
    tab.build_offsets(kp8)
    x2 = _rows512(seq, "Attention")
    res = _Attention.apply(module, x2, tab, bool(return_attn), module.to_q.weight, module.to_kv.weight,
                           module.to_out.weight)
    y, probs = res if return_attn else (res, None)
    y = y.view(B, N, D)
    if context is not None:
        y = y[:, :Nq]
        if probs is not None:
            probs = probs[:, :, :Nq, Nq:]
    y = y.to(x.dtype) if x.dtype != torch.float32 else y
    return (y, probs) if return_attn else y


# ------------------------------------------------------------------------------------------------ encoders
class _EncoderHost:
    """The minimal `model` an encoders-only Engine needs around ONE encoder module (modality name "m")."""

    def __init__(self, encoder):
        self._enc = weakref.ref(encoder)
        self.encoder_specs = [encoder._spec()]

    @property
    def encoders(self):
        return {"m": self._enc()}

    @property
    def training(self):
        return self._enc().training

    def named_parameters(self):
        return [("encoders.m." + n, p) for n, p in self._enc().named_parameters()]


_ENGINES = weakref.WeakKeyDictionary()   # encoder module -> {batch size: Engine}


class _Encode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, host, eng, batch, *params):
        eng.ws["nonfinite"].zero_()
        eng.ws["drop_ctr"].add_(1)
        eng.build_offsets(batch)
        eng.encode(batch)
        ctx.eng = eng
        return eng.ws["xa"][0].view(eng.B, eng.N, D).clone()

    @staticmethod
    def backward(ctx, dx):
        eng = ctx.eng
        eng.flat_grad.zero_()
        eng.garena.zero_()
        eng.encode_backward(dx.contiguous().float().view(eng.M, D))
        if eng.n_desc:
            call("mca_unpack_grads", P(eng.flat_grad), P(eng.garena), P(eng.unpack_descs), eng.n_desc, S())
        flat = eng.flat_grad.clone()
        grads = []
        for name, p in eng._param_list():
            o = eng.offs[name]
            grads.append(flat[o:o + p.numel()].view(p.shape))
        return (None, None, None, *grads)


def encoder_forward(encoder, batch):
    """`encoder(batch) -> (tokens [B, L, 512], attention_mask)` (encoders.py:90-96,114-120,161-166,196-214,268-274)
    through Engine.encode / encode_backward of an encoders-only engine built around this one module.  The engine
    re-homes the encoder's parameters into its flat fp32 buffer (they stay ordinary nn.Parameters), one engine per
    batch size."""
    from .engine import Engine
    from .plan import StaticPlan

    spec = encoder._spec()
    first = next(iter(batch.values()))
    _need_cuda(first, type(encoder).__name__)
    B = int(first.shape[0])
    cache = _ENGINES.setdefault(encoder, {})
    eng = cache.get(B)
    if eng is None:
        plan = StaticPlan({"m": spec}, 0, [2], False, False, True, False, False)
        eng = cache[B] = Engine(_EncoderHost(encoder), plan, depth=0, heads=8, ff_inner=1365, batch_size=B, trunk=False)
    host = eng.model
    eng.ensure_flat()
    if eng.n_desc:   # a SequenceEncoder is a table lookup: no projection matrix to pack
        eng.pack_weights()
    params = [p for _, p in eng._param_list()]
    tokens = _Encode.apply(host, eng, {"m": batch}, *params)
    flag = int(eng.ws["nonfinite"].item())
    if flag & 2:
        raise IndexError("index out of range in self")
    if flag & 1:
        raise Exception("Tokens are not finite")   # encoders.py:197-198
    if spec["type"] == "PatchEncoder":
        mask = eng.ws["enc"]["m"]["mask"].to(torch.long).clone() if getattr(encoder, "attn_mask", True) else None
    else:
        mask = batch["attention_mask"]
    return tokens, mask
