"""GPU (-m gpu): varlen query skipping in the attention kernels (mca_query_skip_flags; north_star subsystem 1 — tokens of
absent / padded modalities are never read).

"exact" (default): all-padded query tiles are skipped only in samples where every modality is present; padded rows of
such a sample feed nothing (every consumer masks padded keys), so the results must equal the un-skipped run up to the
run-to-run noise of the step itself: the forward is reproducible to ~1e-5 (fp32 atomics in the V-mean / loss kernels can
flip single bf16 roundings), the backward to ~3e-3 — every bf16 rounding stage turns an fp32-level perturbation eps into
~sqrt(256 eps) * 2^-9 of flipped roundings, which saturates near 1e-3 after a few stages (scripts/gpu_determinism.py,
scripts/gpu_varlen_diag.py: two runs with skipping OFF differ by exactly as much).  A padded row leaking into a live one
would show at the 1e-2 .. 1 level in the embeddings.
"fast": skipped in every sample; only the pooled rows of ABSENT modalities (the reference's uniform-over-all-N rule,
Q4/Q8) may move, everything a present modality returns stays put."""
import pytest
import torch

from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from oracle import mca_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
dev = "cuda"


def _cfg():
    # long enough for several 128-row tiles per modality, so that ragged tails hold whole dead tiles
    cfg = C.tiny_config("cmu", fcl=True)
    enc = cfg["encoder_configs"]
    enc["COVAREP"]["max_tokens"], enc["FACET"]["max_tokens"], enc["OpenFace"]["max_tokens"] = 640, 300, 260
    return cfg


def _run(model, batch, mode):
    eng = model.engine
    eng.set_varlen(mode)
    for p in model.parameters():
        p.grad = None
    out = model(S.batch_to(batch, dev))
    out["loss"].backward()
    torch.cuda.synchronize()
    emb = {k: v.detach().clone() for k, v in out.items() if isinstance(v, torch.Tensor) and v.dim() == 2}
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return float(out["loss"]), emb, grads, eng.ws["skip_ok"].clone().cpu(), {k: v.cpu() for k, v in out["modality_sample_mask"].items()}


@pytest.mark.parametrize("p_absent", [0.0, 0.4])
def test_exact_mode_changes_nothing(p_absent):
    cfg = _cfg()
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = MCA(**kw).to(dev)
    batch = S.make_batch(cfg, seed=5, variant="dropout_ragged", p_absent=p_absent)
    loss0, emb0, g0, _, present = _run(model, batch, "off")
    loss1, emb1, g1, flags, _ = _run(model, batch, "exact")
    all_present = torch.stack([present[m] for m in present]).all(dim=0)
    assert torch.equal(flags.bool(), all_present)                       # only fully present samples may skip
    if p_absent == 0.0:
        assert bool(flags.all())
        dead_tiles = sum(int((batch[m]["attention_mask"].sum(dim=1) >= 128).sum()) for m in batch)
        assert dead_tiles > 0                                          # the batch does exercise the skip
    assert abs(loss0 - loss1) <= 1e-5 * abs(loss0)
    for k in emb0:
        assert H.rel_err(emb1[k], emb0[k]) < 1e-4, k
    for k in g0:
        assert H.rel_err(g1[k], g0[k]) < 1e-2, k   # run-to-run noise of the bf16 backward (see the module docstring)
    # and both agree with the oracle
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ref = O.mca_forward(sd, kw, batch)
    assert abs(loss1 - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"]))


def test_fast_mode_only_moves_absent_modality_rows():
    cfg = _cfg()
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = MCA(**kw).to(dev)
    batch = S.make_batch(cfg, seed=5, variant="dropout_ragged", p_absent=0.4)
    loss0, emb0, g0, _, present = _run(model, batch, "off")
    loss2, emb2, g2, flags, _ = _run(model, batch, "fast")
    assert bool(flags.all())
    names = list(kw["encoder_configs"].keys())
    for k in emb0:
        if isinstance(k, str) and k in names:
            rows = present[k]                                           # samples in which modality k is present
            assert H.rel_err(emb2[k][rows.to(dev)], emb0[k][rows.to(dev)]) < 1e-4, k
        else:                                                           # fusion rows read fusion tokens only: never padded
            assert H.rel_err(emb2[k], emb0[k]) < 1e-4, k
    # the loss keeps absent samples as negatives (Q7/Q8): it may move, a little
    assert abs(loss2 - loss0) < 2e-2 * abs(loss0)
    model.engine.set_varlen("exact")
    with pytest.raises(ValueError):
        model.engine.set_varlen("sometimes")
