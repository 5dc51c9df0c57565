"""GPU (-m gpu): the fp32-parity forward mode (Engine.set_precision("fp32"), csrc/exact.cu) — north_star's "fp32 loss and
embeddings within 1e-3 relative" against the golden vectors frozen from the live reference, against the oracle, and at
full CMU_config1 size; the split-product GEMM, the fp32 attention core and the fp32 GEGLU on their own."""
import math

import pytest
import torch

from mca_paper_b200 import _lib, config as C, ops, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.ops import P, S as STREAM, call
from oracle import mca_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
dev = "cuda"
FP32_TOL = 1e-3   # north_star: fp32 loss and embeddings within 1e-3 relative

MCA_CASES = [c for c in H.CASES if c != "tiny_cmu_eao"]


@pytest.mark.parametrize("case", MCA_CASES)
def test_fp32_mode_matches_reference_golden(case):
    gold = H.load_golden(case)
    cfg, kw, model, sd, batch = H.build_case(case)
    model = model.to(dev)
    model.engine.set_precision("fp32")
    out = model(S.batch_to(batch, dev))
    assert [H.key_to_str(k) for k in out.keys()] == gold["output_keys"]
    worst = 0.0
    for k, v in out.items():
        if k in ("losses", "modality_sample_mask") or (isinstance(k, str) and "loss" in k):
            continue
        worst = max(worst, H.rel_err(v, gold["embeddings"][H.key_to_str(k)]))
    assert worst < FP32_TOL, worst
    assert abs(out["loss"].item() - gold["loss"].item()) < FP32_TOL * abs(gold["loss"].item())
    for k, v in gold["losses"].items():
        if torch.isnan(v):
            assert torch.isnan(out["losses"][k]), k
        else:
            assert abs(out["losses"][k].item() - v.item()) < FP32_TOL * max(1.0, abs(v.item())), (k, out["losses"][k].item(), v.item())
    for k, v in gold["modality_sample_mask"].items():
        assert torch.equal(out["modality_sample_mask"][k].cpu(), v)
    # the regular backward runs on what this forward saved: gradient norms as in the bf16 test
    out["loss"].backward()
    bad = []
    for k, p in model.named_parameters():
        n = gold["grad_norms"][k]
        if n == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        elif abs(p.grad.norm().item() - n) > 0.15 * n:
            bad.append((k, p.grad.norm().item(), n))
    assert not bad, bad


def test_fp32_mode_is_tighter_than_bf16_and_switchable():
    cfg, kw, model, sd, batch = H.build_case("tiny_cmu_fcl_ragged")
    ref = O.mca_forward(sd, kw, batch)
    model = model.to(dev)
    b = S.batch_to(batch, dev)
    errs = {}
    for mode in ("bf16", "fp32", "bf16"):
        model.engine.set_precision(mode)
        out = model(b)
        errs.setdefault(mode, []).append(max(H.rel_err(out[k], v) for k, v in ref.items()
                                             if isinstance(v, torch.Tensor) and v.dim() == 2))
    assert errs["fp32"][0] < FP32_TOL < 2e-2
    assert errs["fp32"][0] < 0.2 * errs["bf16"][0]
    assert errs["bf16"][0] == errs["bf16"][1]          # switching back restores the bf16 path bit for bit
    with pytest.raises(ValueError):
        model.engine.set_precision("fp16")


def test_split_product_gemm_is_fp32_accurate():
    """[A_hi|A_hi|A_lo] x [W_hi|W_lo|W_hi] over K' = 3K on the tcgen05 GEMM against an fp64 product."""
    torch.manual_seed(0)
    M, N, K = 1000, 512, 320
    A = torch.randn(M, K, device=dev) * 3.0
    W = torch.randn(N, K, device=dev) * 0.05
    a3 = torch.empty(M, 3 * K, device=dev, dtype=torch.bfloat16)
    w3 = torch.empty(N, 3 * K, device=dev, dtype=torch.bfloat16)
    call("mca_x_split_f32", P(A), K, P(a3), M, K, K, 0, STREAM())
    call("mca_x_split_f32", P(W), K, P(w3), N, K, K, 1, STREAM())
    # layouts: hi + lo reproduces the source to 2^-16
    hi, hi2, lo = a3[:, :K].float(), a3[:, K:2 * K].float(), a3[:, 2 * K:].float()
    assert torch.equal(hi, hi2) and torch.equal(hi, A.bfloat16().float())
    assert float(((hi + lo) - A).abs().max() / A.abs().max()) < 2 ** -16
    assert torch.equal(w3[:, :K], w3[:, 2 * K:]) and torch.equal(w3[:, :K].float(), W.bfloat16().float())
    out = torch.empty(M, N, device=dev)
    ops.gemm(a3, 0, w3, 0, M, N, 3 * K, _lib.EPI_F32, out)
    ref = A.double() @ W.double().t()
    err = float((out.double() - ref).norm() / ref.norm())
    plain = torch.empty(M, N, device=dev)
    ops.gemm(A.bfloat16(), 0, W.bfloat16(), 0, M, N, K, _lib.EPI_F32, plain)
    err_bf16 = float((plain.double() - ref).norm() / ref.norm())
    assert err < 2e-5 and err < err_bf16 / 50, (err, err_bf16)


@pytest.mark.parametrize("variant", ["full", "dropout_ragged", "dropout_full"])
def test_fp32_attention_core_matches_dense_reference(variant):
    """mca_x_attn_fwd_f32 against the dense fp32 formula with the reference's -finfo.max fills (fully masked rows are
    uniform over all N keys, Q4)."""
    cfg = C.tiny_config("cmu", fcl=True)
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = MCA(**kw).to(dev)
    eng = model.engine
    eng.ensure_flat()
    batch = S.batch_to(S.make_batch(cfg, seed=4, variant=variant), dev)
    eng.build_offsets(batch)
    B, N, Hh = eng.B, eng.N, eng.H
    torch.manual_seed(1)
    qkv = torch.randn(B * N, 3 * 512, device=dev)
    qkv[:, :512] *= 0.35
    out32 = torch.empty(B * N, 512, device=dev)
    out16 = torch.empty(B * N, 512, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, Hh, N, device=dev)
    call("mca_x_attn_fwd_f32", P(qkv), P(eng.rowbits), P(eng.keygrp), P(eng.ws["padding"]), P(eng.ws["vmean"]), P(out32), P(out16),
         P(lse), B, N, Hh, STREAM())
    t = O.static_tables(kw)
    attn_mask = t["attn_mask"].to(dev)          # True = disallowed
    pad = eng.ws["padding"].to(torch.bool)
    q, k, v = (x.view(B, N, Hh, 64).permute(0, 2, 1, 3) for x in qkv.split(512, dim=1))
    sim = q @ k.transpose(-1, -2)
    big = -torch.finfo(sim.dtype).max
    sim = sim.masked_fill(attn_mask[None, None], big).masked_fill(pad[:, None, None, :], big)
    ref = (sim.softmax(dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * N, 512)
    assert H.rel_err(out32, ref) < 1e-5
    assert H.rel_err(out16.float(), ref) < 5e-3
    dead = (attn_mask[None] | pad[:, None, :]).all(dim=-1)           # [B, N] rows with no live allowed key
    lse_ref = torch.logsumexp(sim, dim=-1)
    live = ~dead[:, None, :].expand(B, Hh, N)
    assert torch.allclose(lse[live], lse_ref[live], rtol=1e-5, atol=1e-5)
    assert bool(torch.isinf(lse[~live]).all())


def test_fp32_geglu_matches_torch_and_saves_backward_factors():
    torch.manual_seed(0)
    M, IP = 300, 256
    u = torch.randn(M, 2 * IP, device=dev) * 1.5
    h3 = torch.empty(M, 3 * IP, device=dev, dtype=torch.bfloat16)
    h16 = torch.empty(M, IP, device=dev, dtype=torch.bfloat16)
    u16 = torch.empty(M, 2 * IP, device=dev, dtype=torch.bfloat16)
    call("mca_x_geglu_f32", P(u), P(h3), P(h16), P(u16), M, IP, STREAM())
    ub = u.view(M, IP // 64, 2, 64)
    x, g = ub[:, :, 0, :].reshape(M, IP), ub[:, :, 1, :].reshape(M, IP)
    h = x * torch.nn.functional.gelu(g)
    got = h3[:, :IP].float() + h3[:, 2 * IP:].float()
    assert float((got - h).abs().max()) < 2 ** -15 * float(h.abs().max()) + 1e-7
    assert torch.equal(h3[:, :IP], h3[:, IP:2 * IP]) and torch.equal(h3[:, :IP], h16)
    cdf = 0.5 * (1 + torch.erf(g / math.sqrt(2)))
    pdf = torch.exp(-0.5 * g * g) / math.sqrt(2 * math.pi)
    fac = u16.view(M, IP // 64, 2, 64).float()
    assert H.rel_err(fac[:, :, 0, :].reshape(M, IP), g * cdf) < 4e-3
    assert H.rel_err(fac[:, :, 1, :].reshape(M, IP), x * (g * pdf + cdf)) < 4e-3


def test_fp32_mode_full_size_cmu_config1_forward():
    """BASELINE.json config 2 at full size (B = 8, N = 2538): loss and every embedding within 1e-3 of the oracle run in
    true fp32 on the same GPU."""
    cfg = C.named_config("CMU_config1")
    kw = C.get_model_config(cfg)
    torch.manual_seed(int(cfg["seed"]))
    model = MCA(**kw)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = S.make_batch(cfg, seed=1, variant="full")
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            sdg = {k: v.to(dev) for k, v in sd.items()}
            tables = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in O.static_tables(kw).items()}
            ref = O.mca_forward(sdg, kw, S.batch_to(batch, dev), tables=tables)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    ref = {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in ref.items()}
    torch.cuda.empty_cache()
    model = model.to(dev)
    model.engine.set_precision("fp32")
    out = model(S.batch_to(batch, dev))
    worst = max(H.rel_err(out[k], v) for k, v in ref.items() if isinstance(v, torch.Tensor) and v.dim() == 2)
    assert worst < FP32_TOL, worst
    assert abs(out["loss"].item() - float(ref["loss"])) < FP32_TOL * abs(float(ref["loss"]))
