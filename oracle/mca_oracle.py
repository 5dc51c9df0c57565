"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32 ops, state_dict driven) of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file; the
product path (mca_paper_b200) never does and fails loudly without its CUDA extension.

What is restated (reference file:line in each function): the static mask builders, the five-layer masked-attention /
GEGLU trunk, attention pooling, the pair-wise temperature-scaled contrastive loss with its NaN-aware reduction, and
the EmbeddedSequence / Tabular encoders.  The loss arithmetic lives in the un-vendored third-party dependency
`torchmultimodal-nightly` (requirements.txt:5, unpinned); it is restated from the reference's own adapted copy
utils/contrastive_loss_with_temperature.py:40-108,178-195 with the three fixes SURVEY.md §8c lists (rank from
torch.distributed, gathered tuple concatenated, gather skipped when not distributed — utils/distributed.py:42-43).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself), so this file is
pinned against the *live* reference imported in the build container (oracle/ref_shim.py): tests/test_oracle_vs_reference.py
compares every buffer bit-exactly and every output/gradient to 1e-5 when /root/reference is present, and
oracle/make_golden.py freezes live-reference outputs into tests/golden/ so the same check runs where the reference
is absent (the GPU box).
"""
from __future__ import annotations

import math
from itertools import chain, combinations
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

FUSION_TOKEN = -1  # model.py:310
GLOBAL_TOKEN = -2  # model.py:311
MASK_VALUE = -torch.finfo(torch.float32).max  # model.py:91,95


# ----------------------------------------------------------------------------------------------- static tables
def modality_combos(n_modalities: int, powers) -> List[frozenset]:
    """model.py:11-12,312 — combinations of modality indices, one cardinality after another, in `powers` order."""
    return [frozenset(c) for c in chain.from_iterable(combinations(range(n_modalities), r) for r in powers)]


def static_tables(spec: dict) -> dict:
    """token_types / return_token_types / attn_mask / pool_mask exactly as MCA.__init__ builds them
    (model.py:312-329, 355, 362-372, 383-446).  True in a mask = "may NOT attend"."""
    enc = spec["encoder_configs"]
    names = list(enc.keys())
    n_mod = len(names)
    zorro = bool(spec.get("zorro", False))
    fcl = bool(spec.get("fcl", False))
    no_fusion = bool(spec.get("no_fusion", False))
    combos = modality_combos(n_mod, spec.get("fusion_combos", [4, 5]))
    n_fusion = 0 if no_fusion else int(spec.get("num_fusion_tokens", 16))

    if no_fusion:
        rtt = list(range(n_mod)) + [GLOBAL_TOKEN]
    elif (not fcl) or zorro:
        rtt = list(range(n_mod)) + [FUSION_TOKEN, GLOBAL_TOKEN]
    else:
        rtt = list(range(n_mod)) + [FUSION_TOKEN] * len(combos) + [GLOBAL_TOKEN]
    rtt_t = torch.tensor(rtt, dtype=torch.long)

    lengths = [int(enc[k]["max_tokens"]) for k in names]
    tt = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(lengths)]
                   + [torch.full((n_fusion,), FUSION_TOKEN, dtype=torch.long)])

    # model.py:392-398: same type may attend; fusion queries may attend everything
    allowed = tt[:, None] == tt[None, :]
    if not no_fusion:
        allowed = allowed | (tt[:, None] == FUSION_TOKEN)
    attn_mask = ~allowed
    # model.py:400-406
    pool_allowed = (rtt_t[:, None] == tt[None, :]) | (rtt_t[:, None] == GLOBAL_TOKEN)
    pool_mask = ~pool_allowed

    if not zorro:
        # model.py:408-430: fusion sub-block c sees only the modalities of combo c and itself
        is_fusion = tt == FUSION_TOKEN
        assert n_fusion % len(combos) == 0, "fusion tokens must divide evenly into combos"
        nsub = n_fusion // len(combos)
        fusion_pos = is_fusion.nonzero().flatten()
        for c, combo in enumerate(combos):
            row = ~torch.isin(tt, torch.tensor(sorted(combo), dtype=torch.long))
            row[is_fusion] = True
            own = fusion_pos[c * nsub:(c + 1) * nsub]
            row[own] = False
            attn_mask[own] = row
        if fcl:
            # model.py:432-446: pooled fusion row c reads only fusion sub-block c
            f_rows = (rtt_t == FUSION_TOKEN).nonzero().flatten()
            for c, r in enumerate(f_rows.tolist()):
                pool_mask[r, fusion_pos] = True
                pool_mask[r, fusion_pos[c * nsub:(c + 1) * nsub]] = False
    return {
        "names": names, "lengths": lengths, "combos": combos, "return_token_types": rtt,
        "token_types": tt, "attn_mask": attn_mask, "pool_mask": pool_mask, "n_fusion": n_fusion,
    }


# ----------------------------------------------------------------------------------------------- encoders
def sinusoid_table(length: int, d_model: int) -> torch.Tensor:
    """encoders.py:128-135."""
    pos = torch.arange(length).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(length, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def embedded_sequence_encode(sd: dict, prefix: str, data: dict):
    """encoders.py:196-214: zero padded rows, LN(in) -> Linear -> LN(512), zero padded rows again, + sinusoidal PE
    at every position (padded included)."""
    tokens, mask = data["tokens"], data["attention_mask"].to(torch.bool)
    if not torch.isfinite(tokens).all():
        raise Exception("non-finite tokens")
    x = tokens.masked_fill(mask.unsqueeze(-1), 0.0)
    w0, b0 = sd[prefix + "token_encoder.0.weight"], sd[prefix + "token_encoder.0.bias"]
    w1, b1 = sd[prefix + "token_encoder.1.weight"], sd[prefix + "token_encoder.1.bias"]
    w2, b2 = sd[prefix + "token_encoder.2.weight"], sd[prefix + "token_encoder.2.bias"]
    x = F.layer_norm(x, (x.shape[-1],), w0, b0)
    x = F.linear(x, w1, b1)
    x = F.layer_norm(x, (x.shape[-1],), w2, b2)
    x = x.masked_fill(mask.unsqueeze(-1), 0.0)
    pe = sd[prefix + "positional_encoder.pe"][: x.shape[1]]
    return x + pe.unsqueeze(0), data["attention_mask"]


def tabular_encode(sd: dict, prefix: str, data: dict, max_value: float, padding_value: float = -1.0,
                   renorm_in_place: bool = True):
    """encoders.py:90-96 with TokenEncoder (:31-37, max_norm=1.0 renormalises the looked-up rows IN PLACE on every
    forward, padding_idx=-1 -> last row) and ContinuousValueEncoder (:55-72: pad test is `== padding_value`,
    clamp(max=max_value), Linear(1,d) -> ReLU -> Linear(d,d) -> LayerNorm, zero where padded)."""
    values = data["values"]
    emb = sd[prefix + "token_encoder.embedding.weight"]
    with torch.no_grad():
        norms = emb.norm(dim=1, keepdim=True)
        scale = torch.where(norms > 1.0, 1.0 / (norms + 1e-7), torch.ones_like(norms))
        if renorm_in_place:
            emb.mul_(scale)  # nn.Embedding(max_norm=...) mutates the weight
    # index = arange(n): every row, in order; the padding_idx row (-1 -> n-1) receives no gradient
    n_rows = emb.shape[0]
    pi = int(padding_value) % n_rows
    keep = torch.ones(n_rows, 1, dtype=emb.dtype, device=emb.device)
    keep[pi] = 0.0
    x_t = emb * keep + (emb * (1.0 - keep)).detach()
    v = values.unsqueeze(-1)
    pad = v == padding_value
    v = torch.clamp(v, max=max_value)
    h = torch.relu(F.linear(v, sd[prefix + "value_encoder.linear1.weight"], sd[prefix + "value_encoder.linear1.bias"]))
    h = F.linear(h, sd[prefix + "value_encoder.linear2.weight"], sd[prefix + "value_encoder.linear2.bias"])
    h = F.layer_norm(h, (h.shape[-1],), sd[prefix + "value_encoder.norm.weight"], sd[prefix + "value_encoder.norm.bias"])
    h = h.masked_fill(pad, 0.0)
    return x_t.unsqueeze(0) + h, data["attention_mask"]


def _lookup_with_renorm(emb: torch.Tensor, idx: torch.Tensor, padding_idx: int, renorm_in_place: bool = True):
    """nn.Embedding(padding_idx, max_norm=1.0) (encoders.py:31-37): the rows that are looked up are renormalised IN
    PLACE (no gradient through the renorm), the padding_idx row receives no gradient."""
    with torch.no_grad():
        used = torch.unique(idx)
        norms = emb[used].norm(dim=1, keepdim=True)
        scale = torch.where(norms > 1.0, 1.0 / (norms + 1e-7), torch.ones_like(norms))
        if renorm_in_place:
            emb[used] = emb[used] * scale
    n_rows = emb.shape[0]
    keep = torch.ones(n_rows, 1, dtype=emb.dtype, device=emb.device)
    keep[padding_idx % n_rows] = 0.0
    table = emb * keep + (emb * (1.0 - keep)).detach()
    return table[idx]


def sequence_encode(sd: dict, prefix: str, data: dict, padding_idx: int = 0):
    """SequenceEncoder.forward encoders.py:161-166: Embedding(tokens) + sinusoidal PE at every position."""
    tokens = data["tokens"].long()
    x_t = _lookup_with_renorm(sd[prefix + "token_encoder.embedding.weight"], tokens, padding_idx)
    pe = sd[prefix + "positional_encoder.pe"][: tokens.shape[1]]
    return x_t + pe.unsqueeze(0), data["attention_mask"]


def sparse_tabular_encode(sd: dict, prefix: str, data: dict, max_value: float, padding_idx: int = 0):
    """SparseTabularEncoder.forward encoders.py:114-120: Embedding(indices) + ContinuousValueEncoder(data), whose pad
    test is `data == padding_idx` (encoders.py:112 passes padding_idx as padding_value)."""
    x_t = _lookup_with_renorm(sd[prefix + "token_encoder.embedding.weight"], data["indices"].long(), padding_idx)
    v = data["data"].unsqueeze(-1)
    pad = v == float(padding_idx)
    v = torch.clamp(v, max=max_value)
    h = torch.relu(F.linear(v, sd[prefix + "value_encoder.linear1.weight"], sd[prefix + "value_encoder.linear1.bias"]))
    h = F.linear(h, sd[prefix + "value_encoder.linear2.weight"], sd[prefix + "value_encoder.linear2.bias"])
    h = F.layer_norm(h, (h.shape[-1],), sd[prefix + "value_encoder.norm.weight"], sd[prefix + "value_encoder.norm.bias"])
    return x_t + h.masked_fill(pad, 0.0), data["attention_mask"]


def patch_encode(sd: dict, prefix: str, data: dict, patch_size, pad_token: float = -10000.0):
    """PatchEncoder.forward encoders.py:268-274 in "matrix" mode, dropout inactive (eval mode or p = 0):
    'b (h p1) (w p2) -> b (h w) (p1 p2)', LN -> Linear -> LN, + learned position embedding; mask = all(patch == pad)."""
    v = data["values"]
    p1, p2 = patch_size
    Bv, Hh, Ww = v.shape
    patches = v.view(Bv, Hh // p1, p1, Ww // p2, p2).permute(0, 1, 3, 2, 4).reshape(Bv, (Hh // p1) * (Ww // p2), p1 * p2)
    x = F.layer_norm(patches, (p1 * p2,), sd[prefix + "batch_to_tokens.1.weight"], sd[prefix + "batch_to_tokens.1.bias"])
    x = F.linear(x, sd[prefix + "batch_to_tokens.2.weight"], sd[prefix + "batch_to_tokens.2.bias"])
    x = F.layer_norm(x, (x.shape[-1],), sd[prefix + "batch_to_tokens.3.weight"], sd[prefix + "batch_to_tokens.3.bias"])
    x = x + sd[prefix + "embedding.weight"][: x.shape[1]].unsqueeze(0)
    mask = torch.all(patches == pad_token, dim=-1).to(torch.long)
    return x, mask


def encode_modality(sd: dict, name: str, cfg: dict, data: dict):
    kind = cfg["type"]
    prefix = f"encoders.{name}."
    if kind == "EmbeddedSequenceEncoder":
        return embedded_sequence_encode(sd, prefix, data)
    if kind == "TabularEncoder":
        # encoders.py:77-88: padding_idx default -1 is forwarded to the value encoder as its padding_value
        return tabular_encode(sd, prefix, data, float(cfg.get("max_value", 10000)), float(cfg.get("padding_idx", -1)))
    if kind == "SequenceEncoder":
        return sequence_encode(sd, prefix, data, int(cfg.get("padding_idx", 0)))
    if kind == "SparseTabularEncoder":
        return sparse_tabular_encode(sd, prefix, data, float(cfg.get("max_value", 10000)), int(cfg.get("padding_idx", 0)))
    if kind == "PatchEncoder":
        return patch_encode(sd, prefix, data, tuple(cfg.get("patch_size", (16, 16))))
    raise NotImplementedError(kind)


# ----------------------------------------------------------------------------------------------- trunk
def masked_attention(x, context, wq, wkv, wo, heads, attn_mask, key_padding_mask):
    """model.py:73-105: no biases, q scaled before QK^T, masks filled with -finfo.max (so a fully masked row
    becomes uniform over ALL keys), softmax, PV, output projection."""
    b, n, _ = x.shape
    ctx = x if context is None else context
    q = F.linear(x, wq)
    k, v = F.linear(ctx, wkv).chunk(2, dim=-1)
    dh = q.shape[-1] // heads
    split = lambda t: t.view(t.shape[0], t.shape[1], heads, dh).permute(0, 2, 1, 3)
    q, k, v = split(q) * dh ** -0.5, split(k), split(v)
    sim = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        sim = sim.masked_fill(attn_mask, MASK_VALUE)
    if key_padding_mask is not None:
        sim = sim.masked_fill(key_padding_mask[:, None, None, :], MASK_VALUE)
    p = sim.softmax(dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(b, n, heads * dh)
    return F.linear(o, wo)


def geglu_ff(x, w1, w2):
    """model.py:35-54: (value, gate) = chunk(W1 x); W2 (gelu(gate) * value); exact erf GELU."""
    val, gate = F.linear(x, w1).chunk(2, dim=-1)
    return F.linear(F.gelu(gate) * val, w2)


def gamma_norm(x, gamma, beta):
    """model.py:24-31."""
    return F.layer_norm(x, (x.shape[-1],), gamma, beta)


def trunk(sd: dict, spec: dict, batch: dict, tables: Optional[dict] = None):
    """model.py:454-473: encoders -> pack with fusion tokens -> depth x MCALayer (ONE shared norm applied twice,
    residual added to the NORMED tensor, model.py:117-122) -> final norm -> attention pooling + return tokens.
    Returns (pooled [B,R,d], modality_sample_mask, final tokens)."""
    t = tables or static_tables(spec)
    B = int(spec.get("batch_size", 8))
    heads = int(spec.get("heads", 8))
    toks, masks = [], []
    for name in t["names"]:
        x, m = encode_modality(sd, name, spec["encoder_configs"][name], batch[name])
        toks.append(x)
        masks.append(m)
    sample_mask = {k: (m == 0).sum(dim=1) != 0 for k, m in zip(t["names"], masks)}  # model.py:458
    if not spec.get("no_fusion", False):
        toks.append(sd["fusion_tokens"].unsqueeze(0).expand(B, -1, -1))
        masks.append(sd["fusion_mask"].unsqueeze(0).expand(B, -1))
    x = torch.cat(toks, dim=1)
    padding = torch.cat([m.to(torch.bool) for m in masks], dim=1)
    for l in range(int(spec["depth"])):
        p = f"layers.{l}."
        g, bta = sd[p + "norm.gamma"], sd[p + "norm.beta"]
        x = gamma_norm(x, g, bta)
        x = masked_attention(x, None, sd[p + "attn.to_q.weight"], sd[p + "attn.to_kv.weight"],
                             sd[p + "attn.to_out.weight"], heads, t["attn_mask"], padding) + x
        x = gamma_norm(x, g, bta)
        x = geglu_ff(x, sd[p + "ff.feedforward.0.weight"], sd[p + "ff.feedforward.2.weight"]) + x
    x = gamma_norm(x, sd["norm.gamma"], sd["norm.beta"])
    rt = sd["return_tokens"].unsqueeze(0).expand(B, -1, -1)
    pooled = masked_attention(rt, x, sd["attn_pool.to_q.weight"], sd["attn_pool.to_kv.weight"],
                              sd["attn_pool.to_out.weight"], heads, t["pool_mask"], padding) + rt
    return pooled, sample_mask, x


# ----------------------------------------------------------------------------------------------- loss
def info_nce(a, b, logit_scale, mask=None, a_all=None, b_all=None, rank: int = 0):
    """utils/contrastive_loss_with_temperature.py:71-100: T = exp(s); logits_a = a b_all^T T, logits_b = b a_all^T T;
    keep rows where mask; labels = B*rank + arange; (CE_a + CE_b)/2.  An empty selection gives NaN (mean of nothing).
    Embeddings are NOT normalised (no normalise anywhere in model.py:448-478)."""
    T = torch.exp(logit_scale)
    a_all = a if a_all is None else a_all
    b_all = b if b_all is None else b_all
    la = (a @ b_all.t()) * T
    lb = (b @ a_all.t()) * T
    labels = a.shape[0] * rank + torch.arange(a.shape[0], device=a.device)
    if mask is not None:
        la, lb, labels = la[mask], lb[mask], labels[mask]
    return (F.cross_entropy(la, labels) + F.cross_entropy(lb, labels)) / 2


class ContrastiveLossWithTemperature(nn.Module):
    """Stand-in for torchmultimodal's module (constructor/forward per utils/contrastive_loss_with_temperature.py:
    156-195; gather per :21-37 and utils/distributed.py:23-56)."""

    def __init__(self, logit_scale=math.log(1 / 0.07), logit_scale_min=math.log(1), logit_scale_max=math.log(100)):
        super().__init__()
        self.logit_scale_min, self.logit_scale_max = logit_scale_min, logit_scale_max
        self.logit_scale = logit_scale if isinstance(logit_scale, nn.Parameter) else nn.Parameter(
            logit_scale * torch.ones([]))

    def forward(self, embeddings_a, embeddings_b, backprop_type=None, cross_entropy_kwargs=None, mask=None):
        self.logit_scale.data.clamp_(self.logit_scale_min, self.logit_scale_max)
        a_all = b_all = None
        rank = 0
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from torch.distributed.nn.functional import all_gather

            a_all = torch.cat(all_gather(embeddings_a), dim=0)
            b_all = torch.cat(all_gather(embeddings_b), dim=0)
            rank = dist.get_rank()
        return info_nce(embeddings_a, embeddings_b, self.logit_scale, mask, a_all, b_all, rank)


def loss_plan(spec: dict, tables: Optional[dict] = None):
    """Order, names and pooled-row indices of every contrastive pair (model.py:160-168, 198-220).
    Returns list of dicts {name, a, b, all, any, fcl}: rows a/b of the pooled block, `all` = modalities that must all
    be present, `any` = modalities of which at least one must be present (empty = no constraint)."""
    t = tables or static_tables(spec)
    names, combos = t["names"], t["combos"]
    n = len(names)
    no_fusion = bool(spec.get("no_fusion", False))
    do_fcl = bool(spec.get("fcl", False)) and not bool(spec.get("zorro", False))
    row = {name: i for i, name in enumerate(names)}
    if do_fcl:
        for i, c in enumerate(combos):
            row[c] = n + i
        if not no_fusion:
            row["fusion"] = row[combos[0]]  # model.py:151,189: fcl_root is always combos[0]
    elif not no_fusion:
        row["fusion"] = n
    if no_fusion:
        pairs = list(combinations(names, 2))
    elif spec.get("bimodal_contrastive", False):
        pairs = list(combinations(names + ["fusion"], 2))
    else:
        pairs = [(m, "fusion") for m in names]
    plan = []
    for pair in pairs:
        # the reference keys its dict by frozenset(pair) and unpacks it again (model.py:167,199): iteration order of
        # a 2-element frozenset of strings depends on string hashing, but only the (symmetric) loss and the sorted
        # name depend on it, so any orientation is equivalent.
        a, b = pair
        need = [names.index(m) for m in pair if m != "fusion"]
        plan.append({"name": "_".join(sorted(pair)), "a": row[a], "b": row[b], "all": need, "any": [], "fcl": False})
    if do_fcl:
        for c in combos[1:]:
            cname = "_".join(sorted(names[i] for i in c))
            if not no_fusion:
                plan.append({"name": f"fcl_fusion|{cname}", "a": row["fusion"], "b": row[c], "all": [],
                             "any": sorted(c), "fcl": True})
            if spec.get("non_fusion_fcl", False):
                for m in names:
                    plan.append({"name": f"fcl_{m}|{cname}", "a": row[m], "b": row[c], "all": [names.index(m)],
                                 "any": sorted(c), "fcl": True})
    return plan, row


def pretraining_loss(pooled, sample_mask: Dict[str, torch.Tensor], logit_scale, spec: dict, tables=None,
                     no_loss=False, pooled_all=None, rank: int = 0):
    """model.py:175-233.  `pooled_all` ([G*B,R,d], columns of every emulated rank) reproduces the all-gather."""
    t = tables or static_tables(spec)
    names = t["names"]
    plan, row = loss_plan(spec, t)
    out = {}
    for key, r in row.items():
        if key == "fusion":
            continue
        out[key] = pooled[:, r, :]
    if "fusion" in row:
        out["fusion"] = pooled[:, row["fusion"], :]
    if no_loss:
        return out
    with torch.no_grad():
        s = logit_scale.data.clamp_(0.0, math.log(100))  # :187, in place
    present = torch.stack([sample_mask[m].to(torch.bool) for m in names], dim=0)  # [n_mod, B]
    losses = {}
    for p in plan:
        m = torch.ones(pooled.shape[0], dtype=torch.bool, device=pooled.device)
        for i in p["all"]:
            m = m & present[i]
        if p["any"]:
            m = m & present[p["any"]].any(dim=0)
        a_all = None if pooled_all is None else pooled_all[:, p["a"], :]
        b_all = None if pooled_all is None else pooled_all[:, p["b"], :]
        losses[p["name"]] = info_nce(pooled[:, p["a"], :], pooled[:, p["b"], :], logit_scale, m, a_all, b_all, rank)
    out["losses"] = losses
    if any(p["fcl"] for p in plan) or (bool(spec.get("fcl", False)) and not bool(spec.get("zorro", False))):
        out["fcl_loss"] = torch.stack([torch.nan_to_num(v) for k, v in losses.items() if "fcl" in k]).mean()
        out["no-fcl_loss"] = torch.stack([torch.nan_to_num(v) for k, v in losses.items() if "fcl" not in k]).mean()
    vals = list(losses.values())
    n_valid = sum(0 if torch.isnan(v.detach()) else 1 for v in vals)
    total = sum(torch.nan_to_num(v) for v in vals)
    out["loss"] = total if n_valid == 0 else total / float(n_valid)
    return out


def mca_forward(sd: dict, spec: dict, batch: dict, no_loss: bool = False, tables=None):
    """Whole MCA.forward (model.py:448-478) on one rank without a process group."""
    t = tables or static_tables(spec)
    pooled, sample_mask, _ = trunk(sd, spec, batch, t)
    out = pretraining_loss(pooled, sample_mask, sd["loss.loss_fn.logit_scale"], spec, t, no_loss=no_loss)
    out["modality_sample_mask"] = sample_mask
    return out


def mean_token_pool(x, key_padding_mask):
    """MeanTokenProjectionPool(token_types_tensor=None, projection=False) as EAO builds it (model.py:553-556, forward
    :257-280): one output token per sample = mean of the rows that are not padded, zeros when there are none; the
    projection is an Identity.  Returns [B, 1, d]."""
    rows = []
    for i in range(x.shape[0]):
        v = x[i, ~key_padding_mask[i], :]
        rows.append(torch.zeros(x.shape[2], dtype=x.dtype, device=x.device) if v.shape[0] == 0 else v.mean(dim=0))
    out = torch.stack(rows).unsqueeze(1)
    if out.isnan().any():
        raise Exception(f"NaN in output from Mean Pooling {int(out.isnan().sum())}, could be from any place before this")
    return out


def eao_passes(spec: dict, tables: Optional[dict] = None):
    """model.py:583: one pass per modality, then one per fusion combination (adjusted_powerset order)."""
    t = tables or static_tables(spec)
    return [[i] for i in range(len(t["names"]))] + [sorted(c) for c in t["combos"]]


def eao_trunk(sd: dict, spec: dict, batch: dict, tables: Optional[dict] = None):
    """EAO.forward model.py:567-590 ("everything at once" baseline): every modality is encoded once; each pass packs the
    tokens of its modalities, runs the SAME layers with no static mask (key padding only, model.py:544-545), the final
    norm and the mean pooling; the pooled tokens of all passes are concatenated."""
    t = tables or static_tables(spec)
    heads = int(spec.get("heads", 8))
    toks, masks = [], []
    for name in t["names"]:
        x, m = encode_modality(sd, name, spec["encoder_configs"][name], batch[name])
        toks.append(x)
        masks.append(m)
    sample_mask = {k: (m == 0).sum(dim=1) != 0 for k, m in zip(t["names"], masks)}  # model.py:580
    pooled = []
    for members in eao_passes(spec, t):
        x = torch.cat([toks[i] for i in members], dim=1)
        padding = torch.cat([masks[i] for i in members], dim=1).to(torch.bool)
        for l in range(int(spec["depth"])):
            p = f"layers.{l}."
            g, bta = sd[p + "norm.gamma"], sd[p + "norm.beta"]
            x = gamma_norm(x, g, bta)
            x = masked_attention(x, None, sd[p + "attn.to_q.weight"], sd[p + "attn.to_kv.weight"],
                                 sd[p + "attn.to_out.weight"], heads, None, padding) + x
            x = gamma_norm(x, g, bta)
            x = geglu_ff(x, sd[p + "ff.feedforward.0.weight"], sd[p + "ff.feedforward.2.weight"]) + x
        x = gamma_norm(x, sd["norm.gamma"], sd["norm.beta"])
        pooled.append(mean_token_pool(x, padding))
    return torch.cat(pooled, dim=1), sample_mask


def eao_forward(sd: dict, spec: dict, batch: dict, no_loss: bool = False, tables=None):
    """Whole EAO.forward (model.py:567-596): the pooled rows are [modalities..., combinations...], which is the row map
    of MCAPretrainingLoss with no_fusion=True (model.py:181-186), i.e. loss_plan() of the same spec."""
    t = tables or static_tables(spec)
    pooled, sample_mask = eao_trunk(sd, spec, batch, t)
    out = pretraining_loss(pooled, sample_mask, sd["loss.loss_fn.logit_scale"], spec, t, no_loss=no_loss)
    out["modality_sample_mask"] = sample_mask
    return out


def mca_forward_ranks(sd: dict, spec: dict, batches: List[dict], tables=None):
    """Single-process emulation of G data-parallel ranks: rank r scores its rows against the
    concatenated columns of every rank; summing the returned per-rank losses and back-propagating gives the
    embedding gradients the all_gather-with-backprop (reduce-scatter SUM) path produces; DDP then divides
    parameter gradients by G."""
    t = tables or static_tables(spec)
    per_rank = [trunk(sd, spec, b, t) for b in batches]
    pooled_all = torch.cat([p for p, _, _ in per_rank], dim=0)
    outs = []
    for r, (pooled, sample_mask, _) in enumerate(per_rank):
        o = pretraining_loss(pooled, sample_mask, sd["loss.loss_fn.logit_scale"], spec, t, pooled_all=pooled_all, rank=r)
        o["modality_sample_mask"] = sample_mask
        outs.append(o)
    return outs


# ----------------------------------------------------------------------------------------------- optimiser step
def clip_adamw_step(params: List[torch.Tensor], grads: List[torch.Tensor], exp_avg, exp_avg_sq, step: int, lr: float,
                    max_norm: float = 2.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """train_accel_gpu.py:80,116-118: clip_grad_norm_(2.0) then torch.optim.AdamW defaults (decoupled decay on
    every parameter).  Updates in place; returns the pre-clip total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        p.mul_(1 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total
