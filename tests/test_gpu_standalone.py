"""Free-standing calls of the reference's sub-modules through the C ABI, against the oracle:
LayerNorm model.py:30-31, FeedForward :49-54, Attention :73-105 (arbitrary mask / context / return_attn), MCALayer
:117-122, encoder.forward(batch) encoders.py:90-96,114-120,161-166,196-214,268-274, functional
contrastive_loss_with_temperature utils/contrastive_loss_with_temperature.py:40-108 with cross_entropy_kwargs and
BackpropType.  Tolerances: fp32 kernels 1e-5, bf16 tensor-core paths 2e-2 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.encoders import encoders_dict
from mca_paper_b200.model import MCA, Attention, FeedForward, LayerNorm, MCALayer
from mca_paper_b200.utils.contrastive_loss_with_temperature import (ContrastiveLossWithTemperature,
                                                                    contrastive_loss_with_temperature)
from mca_paper_b200.utils.distributed import BackpropType
from oracle import mca_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
dev = "cuda"
BF16_TOL = 2e-2


def test_layernorm_module_forward_backward():
    torch.manual_seed(0)
    ln = LayerNorm(512).to(dev)
    with torch.no_grad():
        ln.gamma.copy_(torch.rand(512) + 0.5)
    x = (torch.randn(3, 37, 512, device=dev) * 2 + 0.3).requires_grad_(True)
    y = ln(x)
    g = torch.randn_like(y)
    y.backward(g)
    xr = x.detach().clone().requires_grad_(True)
    gr = ln.gamma.detach().clone().requires_grad_(True)
    yr = O.gamma_norm(xr, gr, ln.beta)
    yr.backward(g)
    assert y.shape == x.shape and rel_err(y, yr) < 1e-5
    assert rel_err(x.grad, xr.grad) < 1e-4 and rel_err(ln.gamma.grad, gr.grad) < 1e-4
    with pytest.raises(ValueError):
        LayerNorm(256).to(dev)(torch.zeros(2, 256, device=dev))


def test_feedforward_module_forward_backward():
    torch.manual_seed(1)
    ff = FeedForward(512, mult=4).to(dev)
    x = torch.randn(2, 333, 512, device=dev, requires_grad=True)
    y = ff(x)
    g = torch.randn_like(y)
    y.backward(g)
    xr = x.detach().clone().requires_grad_(True)
    w1 = ff.feedforward[0].weight.detach().clone().requires_grad_(True)
    w2 = ff.feedforward[2].weight.detach().clone().requires_grad_(True)
    yr = O.geglu_ff(xr, w1, w2)
    yr.backward(g)
    assert rel_err(y, yr) < BF16_TOL
    assert rel_err(x.grad, xr.grad) < BF16_TOL
    assert rel_err(ff.feedforward[0].weight.grad, w1.grad) < BF16_TOL
    assert rel_err(ff.feedforward[2].weight.grad, w2.grad) < BF16_TOL


def _attn_ref(attn, x, context, attn_mask, kpm):
    ws = [p.detach().clone().requires_grad_(True) for p in (attn.to_q.weight, attn.to_kv.weight, attn.to_out.weight)]
    xr = x.detach().clone().requires_grad_(True)
    cr = None if context is None else context.detach().clone().requires_grad_(True)
    y = O.masked_attention(xr, cr, ws[0], ws[1], ws[2], attn.heads, attn_mask, kpm)
    return y, xr, cr, ws


def test_attention_module_arbitrary_block_mask_self_attention():
    torch.manual_seed(2)
    rng = np.random.default_rng(5)
    B, N = 3, 700
    grp = np.sort(rng.integers(0, 6, size=N))            # 6 key groups of uneven length (some shorter than a tile)
    vis = rng.random((6, 6)) < 0.5
    vis[np.arange(6), np.arange(6)] = True
    vis[4, :] = False                                     # rows of group 4 see nothing at all -> uniform over all keys (Q4)
    attn_mask = torch.from_numpy(~vis[grp][:, grp]).to(dev)
    kpm = torch.zeros(B, N, dtype=torch.bool, device=dev)
    kpm[0, 600:] = True
    kpm[1, torch.from_numpy(grp == 2).to(dev)] = True     # a whole group padded for sample 1
    kpm[2, ::3] = True                                    # scattered pads
    attn = Attention(512).to(dev)
    x = torch.randn(B, N, 512, device=dev, requires_grad=True)
    y, probs = attn(x, attn_mask=attn_mask, key_padding_mask=kpm, return_attn=True)
    g = torch.randn_like(y)
    y.backward(g)
    yr, xr, _, ws = _attn_ref(attn, x, None, attn_mask, kpm)
    yr.backward(g)
    assert rel_err(y, yr) < BF16_TOL
    assert rel_err(x.grad, xr.grad) < BF16_TOL
    for p, w in zip((attn.to_q.weight, attn.to_kv.weight, attn.to_out.weight), ws):
        assert rel_err(p.grad, w.grad) < BF16_TOL
    # return_attn: the probabilities of model.py:96 (fully masked rows uniform over all keys)
    q = F.linear(x.detach(), attn.to_q.weight).view(B, N, 8, 64).permute(0, 2, 1, 3) * 64 ** -0.5
    k = F.linear(x.detach(), attn.to_kv.weight)[..., :512].view(B, N, 8, 64).permute(0, 2, 1, 3)
    sim = (q @ k.transpose(-1, -2)).masked_fill(attn_mask, O.MASK_VALUE).masked_fill(kpm[:, None, None, :], O.MASK_VALUE)
    pr = sim.softmax(-1)
    assert probs.shape == (B, 8, N, N) and rel_err(probs, pr) < BF16_TOL
    assert torch.allclose(probs.sum(-1), torch.ones(B, 8, N, device=dev), atol=2e-2)


def test_attention_module_no_mask_and_cross_attention():
    torch.manual_seed(3)
    attn = Attention(512).to(dev)
    x = torch.randn(2, 200, 512, device=dev, requires_grad=True)
    y = attn(x)
    yr, xr, _, ws = _attn_ref(attn, x, None, None, None)
    assert rel_err(y, yr) < BF16_TOL
    # cross attention like the pooling call of model.py:472-473: few query rows, a [Nq, Nk] mask, key padding
    Nq, Nk = 6, 300
    q_in = torch.randn(2, Nq, 512, device=dev, requires_grad=True)
    ctx = torch.randn(2, Nk, 512, device=dev, requires_grad=True)
    am = torch.zeros(Nq, Nk, dtype=torch.bool, device=dev)
    am[0, 100:] = True
    am[1, :100] = True
    am[2, 50:250] = True
    kpm = torch.zeros(2, Nk, dtype=torch.bool, device=dev)
    kpm[1, 280:] = True
    y = attn(q_in, context=ctx, attn_mask=am, key_padding_mask=kpm)
    g = torch.randn_like(y)
    y.backward(g)
    yr, qr, cr, ws = _attn_ref(attn, q_in, ctx, am, kpm)
    yr.backward(g)
    assert y.shape == (2, Nq, 512) and rel_err(y, yr) < BF16_TOL
    assert rel_err(q_in.grad, qr.grad) < BF16_TOL and rel_err(ctx.grad, cr.grad) < BF16_TOL
    for p, w in zip((attn.to_q.weight, attn.to_kv.weight, attn.to_out.weight), ws):
        assert rel_err(p.grad, w.grad) < BF16_TOL
    # a query row that sees no live key is refused rather than averaged over the wrong set
    am[3, :] = True
    with pytest.raises(NotImplementedError):
        attn(q_in, context=ctx, attn_mask=am, key_padding_mask=kpm)
    # more than 32 distinct mask columns cannot be expressed as key groups
    with pytest.raises(NotImplementedError):
        attn(x, attn_mask=torch.from_numpy(np.random.default_rng(0).random((200, 200)) < 0.5).to(dev))


def test_mca_layer_free_standing_matches_reference_wiring():
    torch.manual_seed(4)
    cfg = C.tiny_config("cmu", fcl=True)
    model = MCA(**C.get_model_config(cfg)).to(dev)
    layer: MCALayer = model.layers[0]
    N = model.plan.N
    x = torch.randn(2, N, 512, device=dev)
    pad = torch.zeros(2, N, dtype=torch.bool, device=dev)
    pad[1, 100:150] = True
    y = layer(x, attn_mask=model.attn_mask, padding_mask=pad)
    sd = {k: v.detach() for k, v in layer.state_dict().items()}
    h = O.gamma_norm(x, sd["norm.gamma"], sd["norm.beta"])
    h = O.masked_attention(h, None, sd["attn.to_q.weight"], sd["attn.to_kv.weight"], sd["attn.to_out.weight"], 8,
                           model.attn_mask, pad) + h
    h = O.gamma_norm(h, sd["norm.gamma"], sd["norm.beta"])
    h = O.geglu_ff(h, sd["ff.feedforward.0.weight"], sd["ff.feedforward.2.weight"]) + h
    assert rel_err(y, h) < BF16_TOL


@pytest.mark.parametrize("kind,variant,seed", [("cmu", "dropout_ragged", 1), ("tcga", "tcga", 1), ("mixed", "dropout_ragged", 3)])
def test_encoder_forward_free_standing(kind, variant, seed):
    cfg = C.tiny_config(kind, fcl=True)
    batch = S.make_batch(cfg, seed=seed, variant=variant)
    for name, ecfg in cfg["encoder_configs"].items():
        torch.manual_seed(7)
        enc = encoders_dict[ecfg["type"]](**ecfg)
        enc.eval()
        sd = {f"encoders.{name}." + k: v.detach().clone() for k, v in enc.state_dict().items()}
        want, want_mask = O.encode_modality(sd, name, dict(ecfg), {k: v.clone() for k, v in batch[name].items()})
        # reference gradients through the oracle (parameters that take part)
        enc = enc.to(dev)
        data = {k: v.to(dev) for k, v in batch[name].items()}
        got, mask = enc(data)
        assert got.shape == want.shape, name
        tol = 1e-5 if ecfg["type"] == "SequenceEncoder" else BF16_TOL
        assert rel_err(got, want) < tol, (name, rel_err(got, want))
        if want_mask is not None:
            assert torch.equal(mask.cpu().to(torch.bool), want_mask.to(torch.bool)), name
        g = torch.randn(got.shape, generator=torch.Generator().manual_seed(1))
        got.backward(g.to(dev))
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.endswith(".pe")}
        sd2 = dict(sd)
        sd2.update(params)
        # the oracle renormalises embedding rows in place: hand it the already-renormalised table like the device has
        w2, _ = O.encode_modality(sd2, name, dict(ecfg), {k: v.clone() for k, v in batch[name].items()})
        w2.backward(g)
        for pname, p in enc.named_parameters():
            ref = params[f"encoders.{name}." + pname].grad
            if ref is None or float(ref.abs().max()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) < 1e-6, (name, pname)
                continue
            if ecfg["type"] == "PatchEncoder" and pname == "batch_to_tokens.1.weight":
                # most patches of this batch are constant padding (rstd = eps^-1/2): the input-LayerNorm gain cancels
                # ~10x and shows the bf16 rounding of its upstream gradient at tens of percent; the kernel itself is
                # checked on the SAME upstream gradient in test_gpu_model.test_patch_encoder_backward_matches_torch
                continue
            assert rel_err(p.grad, ref) < 5e-2, (name, pname, rel_err(p.grad, ref))


def test_functional_contrastive_loss_output_and_options():
    torch.manual_seed(11)
    a = (torch.randn(8, 512, device=dev) * 0.1).requires_grad_(True)
    b = (torch.randn(8, 512, device=dev) * 0.1).requires_grad_(True)
    s = torch.nn.Parameter(torch.tensor(2.0, device=dev))
    mask = torch.tensor([1, 0, 1, 1, 1, 0, 1, 1], dtype=torch.bool, device=dev)
    for kwargs in (None, {"label_smoothing": 0.1}, {"reduction": "sum"}):
        for t in (a, b, s):
            t.grad = None
        out = contrastive_loss_with_temperature(a, b, s, mask=mask, cross_entropy_kwargs=kwargs)
        out.loss.backward()
        ar, br, sr = (t.detach().clone().requires_grad_(True) for t in (a, b, s))
        T = torch.exp(sr)
        la, lb = (ar @ br.t() * T)[mask], (br @ ar.t() * T)[mask]
        labels = torch.arange(8, device=dev)[mask]
        kw = kwargs or {}
        loss_a, loss_b = F.cross_entropy(la, labels, **kw), F.cross_entropy(lb, labels, **kw)
        ((loss_a + loss_b) / 2).backward()
        assert out.logits_a.shape == (6, 8) and rel_err(out.logits_a, la) < 1e-5 and rel_err(out.logits_b, lb) < 1e-5
        assert abs(out.loss_a.item() - loss_a.item()) < 1e-5 * max(1, abs(loss_a.item()))
        assert abs(out.loss_b.item() - loss_b.item()) < 1e-5 * max(1, abs(loss_b.item()))
        assert abs(out.loss.item() - ((loss_a + loss_b) / 2).item()) < 1e-5 * max(1, abs(out.loss.item()))
        assert rel_err(a.grad, ar.grad) < 1e-4 and rel_err(b.grad, br.grad) < 1e-4
        assert abs(s.grad.item() - sr.grad.item()) < 1e-4 * max(1, abs(sr.grad.item()))
    # reduction='none' returns per-row losses
    out = contrastive_loss_with_temperature(a, b, s, cross_entropy_kwargs={"reduction": "none"})
    assert out.loss.shape == (8,)
    with pytest.raises(NotImplementedError):
        contrastive_loss_with_temperature(a, b, s, cross_entropy_kwargs={"weight": torch.ones(8, device=dev)})
    # module: options route through the general form (after the in-place clamp), LOCAL / NONE == GLOBAL at world size 1
    mod = ContrastiveLossWithTemperature().to(dev)
    base = mod(a, b, mask=mask)
    for bt in (BackpropType.LOCAL, BackpropType.NONE):
        assert abs(mod(a, b, backprop_type=bt, mask=mask).item() - base.item()) < 1e-5 * max(1, abs(base.item()))
    ls = mod(a, b, cross_entropy_kwargs={"label_smoothing": 0.2}, mask=mask)
    oref = O.ContrastiveLossWithTemperature()
    T = torch.exp(oref.logit_scale.detach().to(dev))
    la, lb = (a.detach() @ b.detach().t() * T)[mask], (b.detach() @ a.detach().t() * T)[mask]
    want = (F.cross_entropy(la, labels, label_smoothing=0.2) + F.cross_entropy(lb, labels, label_smoothing=0.2)) / 2
    assert abs(ls.item() - want.item()) < 1e-5 * max(1, abs(want.item()))
