"""GPU (-m gpu): every kernel through the C ABI against torch fp32 math / the oracle on identical inputs.
Tolerances: bit-exact for integer/byte outputs; bf16-operand kernels 1e-2 relative (bf16 has 8 mantissa bits);
fp32 kernels 1e-5."""
import ctypes
import math

import numpy as np
import pytest
import torch

from mca_paper_b200 import _lib as L, config as C, ops, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.ops import P, call
from oracle import mca_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
dev = "cuda"


def stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(304, 128, 96), (2544, 512, 512), (1000, 1408, 2816)])
def test_gemm_layouts(a_mn, b_mn, M, N, K):
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().t()
    Aop = A.t().contiguous() if a_mn else A
    Bop = B.t().contiguous() if b_mn else B
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(Aop, a_mn, Bop, b_mn, M, N, K, L.EPI_BF16, out)
    assert rel_err(out, ref) < 5e-3
    o32 = torch.zeros(1, M, N, device=dev)
    bias = torch.randn(N, device=dev)
    ops.gemm(Aop, a_mn, Bop, b_mn, M, N, K, L.EPI_F32, o32, bias=bias, alpha=0.5)
    assert rel_err(o32[0], 0.5 * ref + bias) < 1e-5


def test_gemm_split_k_weight_gradient():
    torch.manual_seed(1)
    T, NW, KW = 20304, 512, 1408
    dY = torch.randn(T, NW, device=dev).bfloat16()
    X = torch.randn(T, KW, device=dev).bfloat16()
    eff = ops.effective_splits(T, 6)
    part = torch.zeros(eff, NW, KW, device=dev)
    ops.gemm(dY, 1, X, 1, NW, KW, T, L.EPI_F32, part, ld0=KW, k_splits=6)
    assert rel_err(part.sum(0), dY.float().t() @ X.float()) < 1e-5
    # reduce-add form: every k-split adds into ONE slab (the engine's weight-gradient path)
    acc = torch.zeros(1, NW, KW, device=dev)
    ops.gemm(dY, 1, X, 1, NW, KW, T, L.EPI_F32_ACC, acc, ld0=KW, k_splits=6)
    assert rel_err(acc[0], dY.float().t() @ X.float()) < 1e-5
    ops.gemm(dY, 1, X, 1, NW, KW, T, L.EPI_F32_ACC, acc, ld0=KW, k_splits=3)   # accumulates on top
    assert rel_err(acc[0], 2 * (dY.float().t() @ X.float())) < 1e-5
    small = torch.zeros(1, 128, 64, device=dev)                                  # single-CTA kernel (M < 256)
    ops.gemm(dY[:, :128], 1, X[:, :64], 1, 128, 64, T, L.EPI_F32_ACC, small, ld0=64, k_splits=16, lda=NW, ldb=KW)
    assert rel_err(small[0], dY[:, :128].float().t() @ X[:, :64].float()) < 1e-5


def test_gemm_residual_and_geglu_epilogues():
    torch.manual_seed(2)
    M, I, IP, D = 777, 1365, 1408, 512
    x = torch.randn(M, D, device=dev).bfloat16()
    W = torch.randn(D, D, device=dev).bfloat16()
    res = torch.randn(M, D, device=dev)
    o = torch.zeros(M, D, device=dev)
    ob = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, 0, W, 0, M, D, D, L.EPI_RESID, o, aux0=res, ldaux=D)
    ref = x.float() @ W.float().t() + res
    assert rel_err(o, ref) < 1e-5
    ops.gemm(x, 0, W, 0, M, D, D, L.EPI_BF16, ob)
    assert rel_err(ob, x.float() @ W.float().t()) < 5e-3
    # GEGLU with interleaved W1 (64 value rows | 64 gate rows per 128 block), exact erf GELU (model.py:35-38).
    # The epilogue stores, per (value x, gate g) pair, the two factors the backward needs: a = gelu(g) in the value
    # slot and bv = x * gelu'(g) in the gate slot; h = x * gelu(g) is the forward output.
    W1 = torch.randn(2 * I, D, device=dev) * 0.05
    W1i = torch.zeros(2 * IP, D, device=dev)
    v = torch.arange(IP, device=dev)
    rows_v = (v // 64) * 128 + v % 64
    rows_g = rows_v + 64
    W1i[rows_v[:I]] = W1[:I]
    W1i[rows_g[:I]] = W1[I:]
    W1i = W1i.bfloat16()
    u = torch.zeros(M, 2 * IP, device=dev, dtype=torch.bfloat16)
    h = torch.zeros(M, IP, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, 0, W1i, 0, M, 2 * IP, D, L.EPI_GEGLU, h, ld0=IP, out1=u, ld1=2 * IP)
    uref = x.float() @ W1i.float().t()
    val = uref[:, rows_v].clone().requires_grad_(True)
    gate = uref[:, rows_g].clone().requires_grad_(True)
    href = torch.nn.functional.gelu(gate) * val
    assert rel_err(h, href) < 5e-3
    assert (h[:, I:] == 0).all()                                   # zero padding of the odd inner dim is exact
    assert rel_err(u.float()[:, rows_v], torch.nn.functional.gelu(gate)) < 5e-3
    W2 = (torch.randn(D, IP, device=dev) * 0.05).bfloat16()
    dy = torch.randn(M, D, device=dev).bfloat16()
    du = torch.zeros(M, 2 * IP, device=dev, dtype=torch.bfloat16)
    ops.gemm(dy, 0, W2, 1, M, IP, D, L.EPI_GEGLU_BWD, du, ld0=2 * IP, aux0=u, ldaux=2 * IP)
    href.backward(dy.float() @ W2.float())
    assert rel_err(du.float()[:, rows_v], val.grad) < 1e-2 and rel_err(du.float()[:, rows_g], gate.grad) < 1e-2


def test_gemm_rejects_bad_shapes():
    A = torch.zeros(64, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(AssertionError):
        ops.gemm(A, 0, A, 0, 64, 48, 64, L.EPI_BF16, A)   # N not a multiple of 32 -> MCA_ERR_SHAPE


# ------------------------------------------------------------------------------------------------ LayerNorm
def test_layernorm512_fwd_bwd():
    torch.manual_seed(3)
    rows = 3001
    x = torch.randn(rows, 512, device=dev) * 2 + 0.3
    gamma = torch.rand(512, device=dev) + 0.5
    beta = torch.zeros(512, device=dev)
    y32 = torch.empty_like(x)
    y16 = torch.empty(rows, 512, device=dev, dtype=torch.bfloat16)
    st = torch.empty(rows, 2, device=dev)
    ops.layernorm512_fwd(x, gamma, beta, y32, y16, st, rows)
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (512,), gr, beta)
    assert rel_err(y32, ref) < 1e-5 and rel_err(y16, ref) < 5e-3
    dy = torch.randn_like(x)
    ref.backward(dy)
    dx = torch.empty_like(x)
    dx16 = torch.empty(rows, 512, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(512, device=dev)
    ops.layernorm512_bwd(dy, x, st, gamma, dx, dx16, dg, None, rows)
    assert rel_err(dx, xr.grad) < 1e-4 and rel_err(dg, gr.grad) < 1e-4 and rel_err(dx16, xr.grad) < 5e-3


def test_add_layernorm_fused_residual_and_delta_backward():
    """x_new = LN(x_prev) + y (quirk Q1 residual) fused with the next LayerNorm; backward with a bf16 branch delta."""
    torch.manual_seed(4)
    rows = 2049
    xp = torch.randn(rows, 512, device=dev) * 1.5 - 0.2
    g0, g1 = torch.rand(512, device=dev) + 0.5, torch.rand(512, device=dev) + 0.5
    beta = torch.zeros(512, device=dev)
    st0 = torch.empty(rows, 2, device=dev)
    tmp16 = torch.empty(rows, 512, device=dev, dtype=torch.bfloat16)
    ops.layernorm512_fwd(xp, g0, beta, None, tmp16, st0, rows)
    y = (torch.randn(rows, 512, device=dev) * 0.7).bfloat16()
    xnew = torch.empty_like(xp)
    out16 = torch.empty(rows, 512, device=dev, dtype=torch.bfloat16)
    st1 = torch.empty(rows, 2, device=dev)
    ops.add_layernorm512_fwd(xp, st0, g0, beta, y, xnew, g1, beta, out16, st1, rows)
    ref_new = torch.nn.functional.layer_norm(xp, (512,), g0, beta) + y.float()
    assert rel_err(xnew, ref_new) < 1e-5
    assert rel_err(out16, torch.nn.functional.layer_norm(ref_new, (512,), g1, beta)) < 5e-3
    # backward: dy (fp32 residual path) + delta (bf16 branch path)
    xr = ref_new.clone().requires_grad_(True)
    dy = torch.randn_like(xp)
    dd = (torch.randn_like(xp) * 0.5).bfloat16()
    torch.nn.functional.layer_norm(xr, (512,), g1, beta).backward(dy + dd.float())
    dx = torch.empty_like(xp)
    dg = torch.zeros(512, device=dev)
    ops.layernorm512_bwd(dy, xnew, st1, g1, dx, None, dg, None, rows, dy_delta=dd)
    assert rel_err(dx, xr.grad) < 1e-4


# ------------------------------------------------------------------------------------------------ offsets
@pytest.mark.parametrize("cfg_name,variant", [("CMU_config1_d40", "dropout_ragged"), ("TCGA_config1", "tcga")])
def test_build_offsets_bit_exact(cfg_name, variant):
    cfg = C.named_config(cfg_name)
    model = MCA(**C.get_model_config(cfg)).to(dev)
    eng = model.engine
    eng.ensure_flat()
    batch = S.make_batch(cfg, seed=5, variant=variant)
    eng.build_offsets(S.batch_to(batch, dev))
    torch.cuda.synchronize()
    pl, ws = eng.plan, eng.ws
    masks = [batch[n]["attention_mask"].bool() for n in pl.names]
    padding = torch.cat(masks + [torch.zeros(eng.B, pl.F, dtype=torch.bool)], dim=1)
    assert torch.equal(ws["padding"].cpu().bool(), padding)                                   # == reference `padding`
    present = torch.stack([(m == 0).sum(1) != 0 for m in masks], dim=1)                       # model.py:458
    assert torch.equal(ws["present"].cpu().bool(), present)
    live = torch.stack([(~m).sum(1) for m in masks], dim=1).int()
    assert torch.equal(ws["live_count"].cpu(), live)
    cu = torch.cat([torch.zeros(1, dtype=torch.int64), live.flatten().cumsum(0)]).int()
    assert torch.equal(ws["cu_live"].cpu(), cu)
    idx = ws["live_idx"].cpu()
    for b in range(eng.B):
        for i, (o, n) in enumerate(zip(pl.offsets, pl.lengths)):
            want = (~masks[i][b]).nonzero().flatten().int() + o
            got = idx[b, o:o + n]
            assert torch.equal(got[:len(want)], want) and (got[len(want):] == -1).all()
    cls = ws["kt_class"].cpu()
    for b in range(eng.B):
        for kt, (s, l) in enumerate(pl.tiles):
            npad = int(padding[b, s:s + l].sum())
            assert cls[b, kt] == (0 if npad == 0 else (2 if npad == l else 1))
    assert int(ws["any_absent"].item()) == int((~present).any())


# ------------------------------------------------------------------------------------------------ pooling
@pytest.mark.parametrize("N,R", [(2538, 16), (333, 6), (40, 16)])
def test_pool_attention_cluster_kernels(N, R):
    """Attention pooling core (model.py:472-473) forward + backward against torch fp32, incl. a fully masked row
    (uniform 1/N over ALL keys, no score gradient) and padded keys."""
    torch.manual_seed(N + R)
    B, H = 3, 8
    g = torch.Generator(device=dev).manual_seed(7)
    qp = torch.randn(R, 512, device=dev, generator=g) * 0.2
    kv = torch.randn(B * N, 1024, device=dev, generator=g).bfloat16()
    keygrp = torch.randint(0, 5, (N,), device=dev, generator=g).to(torch.uint8)
    rowbits = torch.randint(1, 32, (R,), device=dev, generator=g).to(torch.int32)
    rowbits[-1] = 31
    padding = (torch.rand(B, N, device=dev, generator=g) < 0.3).to(torch.uint8)
    padding[1, keygrp == 2] = 1                       # sample 1: every key of group 2 padded ...
    rowbits[0] = 1 << 2                               # ... and row 0 may only see group 2 -> fully masked there
    probs = torch.zeros(B, H, R, N, device=dev)
    fm = torch.zeros(B, R, device=dev, dtype=torch.uint8)
    out = torch.zeros(B, R, 512, device=dev)
    call("mca_pool_attn_fwd", P(qp), P(kv), P(padding), P(keygrp), P(rowbits), P(probs), P(fm), P(out), B, H, R, N, stream())
    q = qp.view(R, H, 64).permute(1, 0, 2).clone().requires_grad_(True)                      # [H,R,64]
    kvf = kv.float().view(B, N, 2, H, 64).requires_grad_(True)
    k, v = kvf[:, :, 0].permute(0, 2, 1, 3), kvf[:, :, 1].permute(0, 2, 1, 3)                 # [B,H,N,64]
    allowed = ((rowbits.long()[:, None] >> keygrp.long()[None, :]) & 1).bool()               # [R,N]
    sim = torch.einsum("hrd,bhnd->bhrn", q, k)
    mv = O.MASK_VALUE
    sim = sim.masked_fill(~allowed[None, None], mv).masked_fill(padding.bool()[:, None, None, :], mv)
    pr = sim.softmax(-1)
    o = torch.einsum("bhrn,bhnd->brhd", pr, v).reshape(B, R, 512)
    assert rel_err(probs, pr) < 1e-4 and rel_err(out, o) < 1e-4
    full = (sim.max(-1).values == mv).all(1)                                                  # [B,R]
    assert torch.equal(fm.bool(), full) and bool(full[1, 0])
    do = torch.randn(B, R, 512, device=dev, generator=g)
    o.backward(do)
    dkv = torch.zeros(B * N, 1024, device=dev, dtype=torch.bfloat16)
    dqp = torch.zeros(R, 512, device=dev)
    scratch = torch.zeros(B, H, R, N, device=dev)
    call("mca_pool_attn_bwd", P(do), P(qp), P(kv), P(probs), P(fm), P(scratch), P(dkv), P(dqp), B, H, R, N, stream())
    assert rel_err(dkv.float(), kvf.grad.reshape(B * N, 1024)) < 1e-2                          # bf16 output rounding
    assert rel_err(dqp, q.grad.permute(1, 0, 2).reshape(R, 512)) < 1e-4


@pytest.mark.parametrize("M,N,K,ta,tb", [(16, 512, 512, 0, 0), (128, 512, 512, 0, 1), (512, 512, 128, 1, 1),
                                         (512, 512, 16, 1, 1), (37, 70, 100, 0, 0)])
def test_small_gemm_f32_cluster_split_k(M, N, K, ta, tb):
    torch.manual_seed(M + K)
    A = torch.randn(K, M, device=dev) if ta else torch.randn(M, K, device=dev)
    Bm = torch.randn(K, N, device=dev) if tb else torch.randn(N, K, device=dev)
    ref = (A.t() if ta else A) @ (Bm if tb else Bm.t())
    Cm = torch.randn(M, N, device=dev)
    c0 = Cm.clone()
    add = torch.randn(8, N, device=dev)
    ops.small_gemm(A, 1 if ta else K, M if ta else 1, Bm, 1 if tb else K, N if tb else 1, Cm, N, M, N, K, alpha=0.5,
                   accumulate=True, add=add, ldadd=N, add_rows=8)
    want = 0.5 * ref + c0 + add[torch.arange(M, device=dev) % 8]
    assert rel_err(Cm, want) < 1e-5


def test_column_reductions():
    torch.manual_seed(4)
    rows, width, kp = 1000, 713, 768
    dy = torch.randn(rows, kp, device=dev)
    x = torch.randn(rows, width, device=dev)
    mean, var = x.mean(1), x.var(1, unbiased=False)
    stats = torch.stack([mean, (var + 1e-5).rsqrt()], 1).contiguous()
    pad = (torch.rand(rows, device=dev) < 0.2).to(torch.uint8)
    dw, db = torch.zeros(width, device=dev), torch.zeros(width, device=dev)
    call("mca_layernorm_in_param_bwd", P(dy), kp, P(x), P(stats), P(pad), P(dw), P(db), width, rows, stream())
    keep = (pad == 0).float()[:, None]
    xhat = (x - mean[:, None]) * stats[:, 1:2]
    assert rel_err(dw, (dy[:, :width] * xhat * keep).sum(0)) < 1e-5
    assert rel_err(db, (dy[:, :width] * keep).sum(0)) < 1e-5
    out = torch.zeros(512, device=dev)
    a = torch.randn(rows, 512, device=dev)
    call("mca_colsum", P(a), 512, P(out), 512, rows, stream())
    assert rel_err(out, a.sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ index encoders
def test_embedding_renorm_gather_scatter():
    """nn.Embedding(max_norm=1, padding_idx) pieces of SequenceEncoder / SparseTabularEncoder against torch."""
    torch.manual_seed(5)
    V, B, Lq, d, N, off = 60, 4, 24, 512, 40, 7
    emb = torch.randn(V, d, device=dev) * 0.06          # row norms around 1.36 -> most get renormalised
    emb[3] *= 0.1                                       # one row below max_norm stays untouched
    idx = torch.randint(0, 20, (B, Lq), device=dev)     # rows 20.. are never looked up -> never renormalised
    # expected weights: every looked-up row renormalised ONCE (what the reference's CPU nn.Embedding does; torch's CUDA
    # embedding_renorm_ races on duplicate indices and is itself off by 1-3 % in most runs, scripts/gpu_dbg_renorm.py)
    want = emb.clone()
    u = torch.unique(idx)
    nrm = want[u].norm(dim=1, keepdim=True)
    want[u] = torch.where(nrm > 1.0, want[u] * (1.0 / (nrm + 1e-7)), want[u])
    flags = torch.zeros(V, device=dev, dtype=torch.uint8)
    bad = torch.zeros(1, device=dev, dtype=torch.int32)
    call("mca_embedding_renorm_indexed", P(emb), P(idx), B * Lq, V, d, 1.0, P(flags), P(bad), stream())
    assert rel_err(emb, want) < 1e-6 and int(flags.sum()) == 0 and int(bad.item()) == 0
    assert torch.equal(emb[20:], want[20:])             # untouched rows are bit-identical
    ref_mod = torch.nn.Embedding(V, d, padding_idx=0).to(dev)   # lookups / gradients on the renormalised table
    with torch.no_grad():
        ref_mod.weight.copy_(emb)
    out_ref = ref_mod(idx)
    pe = torch.randn(Lq, d, device=dev)
    dst = torch.zeros(B * N, d, device=dev)
    call("mca_embedding_gather", P(emb), P(idx), V, B, Lq, d, P(pe), P(dst), N, off, 0, stream())
    got = dst.view(B, N, d)[:, off:off + Lq]
    assert rel_err(got, out_ref + pe) < 1e-6
    call("mca_embedding_gather", P(emb), P(idx), V, B, Lq, d, None, P(dst), N, off, 1, stream())      # accumulate form
    assert rel_err(dst.view(B, N, d)[:, off:off + Lq], 2 * out_ref + pe) < 1e-6
    dsrc = torch.randn(B * N, d, device=dev)
    out_ref.backward(dsrc.view(B, N, d)[:, off:off + Lq])
    demb = torch.zeros(V, d, device=dev)
    call("mca_embedding_scatter_add", P(dsrc), P(idx), V, B, Lq, d, N, off, 0, P(demb), stream())
    assert rel_err(demb, ref_mod.weight.grad) < 1e-5 and float(demb[0].abs().max()) == 0.0             # padding row
    idx[1, 2] = V + 5                                                                                   # out of range
    call("mca_embedding_renorm_indexed", P(emb), P(idx), B * Lq, V, d, 1.0, P(flags), P(bad), stream())
    assert int(bad.item()) == 2


def test_patchify_bit_exact():
    torch.manual_seed(6)
    B, Hh, Ww, p1, p2 = 3, 16, 48, 4, 8
    v = torch.randn(B, Hh, Ww, device=dev)
    v[1, :, 24:] = -10000.0
    v[2] = -10000.0
    L = (Hh // p1) * (Ww // p2)
    tok = torch.zeros(B * L, p1 * p2, device=dev)
    mask = torch.zeros(B, L, device=dev, dtype=torch.uint8)
    call("mca_patchify", P(v), B, Hh, Ww, p1, p2, -10000.0, P(tok), P(mask), stream())
    want = v.view(B, Hh // p1, p1, Ww // p2, p2).permute(0, 1, 3, 2, 4).reshape(B, L, p1 * p2)   # 'b (h p1) (w p2) -> b (h w) (p1 p2)'
    assert torch.equal(tok.view(B, L, -1), want)
    assert torch.equal(mask.bool(), (want == -10000.0).all(-1))


# ------------------------------------------------------------------------------------------------ attention
def _attention_case(cfg, variant, scale):
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = MCA(**kw).to(dev)
    eng = model.engine
    eng.ensure_flat()
    eng.set_varlen("off")   # every row is compared, padded query rows included (varlen skipping: tests/test_gpu_varlen.py)
    eng.build_offsets(S.batch_to(S.make_batch(cfg, seed=1, variant=variant), dev))
    B, N, H, M = eng.B, eng.N, eng.H, eng.M
    g = torch.Generator(device=dev).manual_seed(3)
    qkv = (torch.randn(M, 1536, device=dev, generator=g) * scale).bfloat16()
    out = torch.zeros(M, 512, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device=dev)
    eng.attention_fwd(qkv, out, lse)
    x = qkv.float().view(B, N, 3, H, 64).requires_grad_(True)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    mv = O.MASK_VALUE
    sim = (q @ k.transpose(-1, -2)).masked_fill(model.attn_mask, mv).masked_fill(eng.ws["padding"].bool()[:, None, None, :], mv)
    o = (sim.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(M, 512)
    full = sim.max(-1).values == mv
    assert rel_err(out, o) < 5e-3
    assert torch.equal(torch.isinf(lse), full)            # fully masked rows flagged, nothing else
    assert rel_err(lse[~full], torch.logsumexp(sim, -1)[~full]) < 1e-5
    do = torch.randn(M, 512, device=dev, generator=g).bfloat16()
    o.backward(do.float())
    ws = eng.ws
    ws["dattn"].copy_(do)
    call("mca_attn_bwd", P(qkv), P(out), P(ws["dattn"]), P(lse), P(eng.k_tiles_q), eng.n_kt, P(eng.qt_list), P(eng.k_tiles),
         int(eng.q_tiles.shape[0]), P(eng.rowbits), P(eng.keygrp), P(eng.tile_grp), P(ws["padding"]), P(ws["kt_class"]),
         None, P(ws["delta"]), P(ws["ucorr"]), P(ws["dq_acc"]), P(ws["dqkv"]), B, N, H, stream())
    gref = x.grad.view(M, 1536)
    d = ws["dqkv"].float()
    for sl in (slice(0, 512), slice(512, 1024), slice(1024, 1536)):
        assert rel_err(d[:, sl], gref[:, sl]) < 1e-2
    return int(full.sum())


def test_attention_mca_full_length():
    assert _attention_case(C.tiny_config("cmu", fcl=True), "full", 1.0) == 0


def test_attention_ragged_and_absent_modalities():
    # fully masked query rows must become uniform over ALL keys (reference quirk Q4) and feed dV with 1/N
    assert _attention_case(C.tiny_config("cmu", fcl=True), "dropout_ragged", 1.0) > 0


def test_attention_mma_mask():
    assert _attention_case(C.tiny_config("cmu", zorro=True, fcl=False), "dropout_full", 0.5) > 0


def test_attention_scattered_padding_tcga():
    _attention_case(C.tiny_config("tcga", fcl=True), "tcga", 0.7)


def test_attention_full_size_cmu():
    _attention_case(C.named_config("CMU_config1"), "full", 0.5)


# ------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize("world,rank", [(1, 0), (4, 2)])
def test_allpairs_infonce_fwd_bwd(world, rank):
    cfg = C.tiny_config("tcga", fcl=True, bimodal=True, non_fusion_fcl=True)
    kw = C.get_model_config(cfg)
    model = MCA(**kw).to(dev)
    eng = model.engine
    eng.ensure_flat()
    B, R, GB = eng.B, eng.R, world * eng.B
    torch.manual_seed(7)
    pooled_all = (torch.randn(GB, R, 512, device=dev) * 0.2).requires_grad_(True)
    present = torch.rand(B, eng.plan.n_mod, device=dev) > 0.3
    present[1] = False                      # a sample with nothing present
    present_u8 = present.to(torch.uint8).contiguous()
    s = torch.tensor(5.0, device=dev)       # above ln(100): must be clamped in place
    losses = torch.empty(eng.plan.n_pairs, device=dev)
    summary = torch.empty(4, device=dev)
    w = torch.empty(eng.plan.n_pairs, device=dev)
    call("mca_contrastive_allpairs_fwd", P(pooled_all.detach()), P(present_u8), P(eng.loss_plan), eng.plan.n_pairs, P(s), B, GB,
         R, 512, eng.plan.n_mod, rank, 0.0, math.log(100), P(losses), P(summary), P(w), stream())
    assert abs(s.item() - math.log(100)) < 1e-6
    t = O.static_tables(kw)
    sm = {n: present[:, i].cpu() for i, n in enumerate(t["names"])}
    sref = torch.nn.Parameter(torch.tensor(5.0))
    pa = pooled_all.detach().cpu().requires_grad_(True)
    ref = O.pretraining_loss(pa[rank * B:(rank + 1) * B], sm, sref, kw, t, pooled_all=pa, rank=rank)
    for i, name in enumerate(eng.plan.loss_names):
        r = ref["losses"][name]
        assert torch.isnan(r) == torch.isnan(losses[i].cpu()), name
        if not torch.isnan(r):
            assert abs(losses[i].item() - r.item()) < 1e-4 * max(1.0, abs(r.item())), name
    assert abs(summary[0].item() - ref["loss"].item()) < 1e-4 * abs(ref["loss"].item())
    assert abs(summary[1].item() - ref["fcl_loss"].item()) < 1e-4 * abs(ref["fcl_loss"].item())
    assert abs(summary[2].item() - ref["no-fcl_loss"].item()) < 1e-4 * abs(ref["no-fcl_loss"].item())
    ref["loss"].backward()
    dall = torch.zeros(GB, R, 512, device=dev)
    ds = torch.zeros(1, device=dev)
    call("mca_contrastive_allpairs_bwd", P(pooled_all.detach()), P(present_u8), P(eng.loss_plan), eng.plan.n_pairs, P(s), B, GB,
         R, 512, eng.plan.n_mod, rank, P(w), P(dall), P(ds), stream())
    assert rel_err(dall, pa.grad) < 1e-4
    assert abs(ds.item() - sref.grad.item()) < 1e-4 * max(1.0, abs(sref.grad.item()))


def test_standalone_contrastive_loss_module():
    from mca_paper_b200.utils.contrastive_loss_with_temperature import ContrastiveLossWithTemperature

    torch.manual_seed(11)
    a = (torch.randn(8, 512, device=dev) * 0.1).requires_grad_(True)
    b = (torch.randn(8, 512, device=dev) * 0.1).requires_grad_(True)
    mask = torch.tensor([1, 0, 1, 1, 1, 0, 1, 1], dtype=torch.bool, device=dev)
    mod = ContrastiveLossWithTemperature().to(dev)
    loss = mod(a, b, mask=mask)
    loss.backward()
    ar, br = a.detach().cpu().requires_grad_(True), b.detach().cpu().requires_grad_(True)
    oref = O.ContrastiveLossWithTemperature()
    lref = oref(ar, br, mask=mask.cpu())
    lref.backward()
    assert abs(loss.item() - lref.item()) < 1e-5 * max(1, abs(lref.item()))
    assert rel_err(a.grad, ar.grad) < 1e-4 and rel_err(b.grad, br.grad) < 1e-4
    assert abs(mod.logit_scale.grad.item() - oref.logit_scale.grad.item()) < 1e-4 * max(1, abs(oref.logit_scale.grad.item()))
    # empty selection -> NaN (mean over nothing), as F.cross_entropy on an empty batch
    assert torch.isnan(mod(a, b, mask=torch.zeros(8, dtype=torch.bool, device=dev)))


# ------------------------------------------------------------------------------------------------ optimiser
def test_clip_adamw_matches_torch():
    torch.manual_seed(5)
    n = 1_000_003
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev) * 0.01
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([pr], lr=1e-3)
    step = torch.zeros(1, device=dev, dtype=torch.int64)
    sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
    tn = torch.zeros(1, device=dev)
    cfg = ops.AdamWCfg(1e-3, 0.9, 0.999, 1e-8, 0.01, 2.0, 0, 0, 1)
    for it in range(3):
        pr.grad = g.clone() * (it + 1)
        total = torch.nn.utils.clip_grad_norm_([pr], 2.0)
        opt.step()
        call("mca_clip_adamw_step", P(p), P(g * (it + 1)), P(m), P(v), n, P(sumsq), P(step), P(tn), 1.0,
             ctypes.addressof(cfg), stream())
        assert abs(tn.item() - total.item()) < 1e-4 * total.item()
    assert int(step.item()) == 3
    assert rel_err(p, pr.detach()) < 1e-6


def test_pack_unpack_roundtrip():
    cfg = C.tiny_config("cmu", fcl=True)
    model = MCA(**C.get_model_config(cfg)).to(dev)
    eng = model.engine
    eng.ensure_flat()
    eng.pack_weights()
    torch.cuda.synchronize()
    I, IP = eng.I, eng.IP
    w1 = model.layers[0].ff.feedforward[0].weight.detach()
    pk = eng.W("layers.0.ff1").float()
    v = torch.arange(I, device=dev)
    rows_v = (v // 64) * 128 + v % 64
    assert rel_err(pk[rows_v], w1[:I]) < 4e-3 and rel_err(pk[rows_v + 64], w1[I:]) < 4e-3
    qkv = eng.W("layers.0.qkv").float()
    assert rel_err(qkv[:512], model.layers[0].attn.to_q.weight.detach() * 0.125) < 4e-3   # exact 2**-3 scale fold
    assert rel_err(qkv[512:], model.layers[0].attn.to_kv.weight.detach()) < 4e-3
    w2 = eng.W("layers.0.ff2").float()
    assert (w2[:, I:] == 0).all() and rel_err(w2[:, :I], model.layers[0].ff.feedforward[2].weight.detach()) < 4e-3
    # unpack: put a known pattern in the partial arena and read it back in state_dict layout
    part, _ = eng.GW("layers.0.ff1")
    part.zero_()
    part[0] = pk
    call("mca_unpack_grads", P(eng.flat_grad), P(eng.garena), P(eng.unpack_descs), eng.n_desc, stream())
    got = eng.gview("layers.0.ff.feedforward.0.weight")
    assert rel_err(got, w1) < 4e-3


@pytest.mark.parametrize("B,starts", [(3, [0, 5, 5, 300, 1000]), (8, [0, 150, 195, 265, 285, 480, 700])])
def test_mean_pool_fwd_bwd_match_torch(B, starts):
    """mca_mean_pool_fwd / _bwd (MeanTokenProjectionPool as EAO uses it, model.py:257-280,553-563) against torch: per
    pass mean over the live tokens, zeros for a pass without any (and for an empty pass), gradient 1/count on live rows."""
    torch.manual_seed(B)
    N, R = starts[-1], len(starts) - 1
    x = torch.randn(B, N, 512, device=dev).to(torch.bfloat16)
    pad = torch.rand(B, N, device=dev) < 0.3
    pad[0, starts[0]:starts[1]] = True                      # a pass with no live token
    if B > 1:
        pad[1] = False
    pad8 = pad.to(torch.uint8).contiguous()
    ps = torch.tensor(starts, device=dev, dtype=torch.int32)
    tok_pass = torch.zeros(N, device=dev, dtype=torch.int32)
    for r in range(R):
        tok_pass[starts[r]:starts[r + 1]] = r
    pooled, cnt = torch.empty(B, R, 512, device=dev), torch.empty(B, R, device=dev)
    scratch = torch.empty(int(ops.fn("mca_mean_pool_scratch_floats")(B, R)), device=dev)
    call("mca_mean_pool_fwd", P(x), P(pad8), P(ps), B, N, R, 512, P(pooled), P(cnt), P(scratch), stream())
    xr = x.float().requires_grad_(True)
    want = torch.zeros(B, R, 512, device=dev)
    rows = []
    for b in range(B):
        for r in range(R):
            live = ~pad[b, starts[r]:starts[r + 1]]
            seg = xr[b, starts[r]:starts[r + 1]][live]
            rows.append(seg.mean(0) if seg.shape[0] else torch.zeros(512, device=dev))
            assert float(cnt[b, r]) == float(live.sum())
    want = torch.stack(rows).view(B, R, 512)
    assert rel_err(pooled, want) < 1e-5
    assert float(pooled[0, 0].abs().max()) == 0.0
    dp = torch.randn(B, R, 512, device=dev)
    want.backward(dp)
    dx = torch.full((B, N, 512), 7.0, device=dev)
    call("mca_mean_pool_bwd", P(dp), P(pad8), P(tok_pass), P(cnt), B, N, R, 512, P(dx), stream())
    assert rel_err(dx, xr.grad) < 1e-5
    assert float(dx[pad].abs().max()) == 0.0
