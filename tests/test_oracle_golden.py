"""CPU: the oracle restatement (oracle/mca_oracle.py) against golden vectors frozen from the LIVE reference
(oracle/make_golden.py), and against the live reference itself when /root/reference is present."""
import hashlib

import pytest
import torch

from oracle import mca_oracle as O, ref_shim
from tests import helpers as H


@pytest.mark.parametrize("case", list(H.CASES))
def test_oracle_matches_golden(case):
    gold = H.load_golden(case)
    cfg, kw, model, sd, batch = H.build_case(case)
    # same seeded weights / inputs as when the fixture was produced (guards against silent RNG drift)
    assert H.weights_checksum(sd) == gold["weights_sha256"]
    assert H.batch_checksum(batch) == gold["batch_sha256"]
    names = [k for k, _ in model.named_parameters()]
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    sd2 = dict(sd)
    sd2.update(params)
    out = H.oracle_forward(kw)(sd2, kw, batch)
    assert [H.key_to_str(k) for k in out.keys()] == gold["output_keys"]          # Q17: keys and order
    assert list(out["losses"].keys()) == list(gold["losses"].keys())
    for k, v in out.items():
        if k in ("losses", "modality_sample_mask") or (isinstance(k, str) and "loss" in k):
            continue
        torch.testing.assert_close(v, gold["embeddings"][H.key_to_str(k)], rtol=1e-4, atol=1e-4)
    for k, v in gold["losses"].items():
        if torch.isnan(v):
            assert torch.isnan(out["losses"][k]), k
        else:
            torch.testing.assert_close(out["losses"][k], v, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["loss"], gold["loss"], rtol=1e-5, atol=1e-5)
    if gold["fcl_loss"] is not None:
        torch.testing.assert_close(out["fcl_loss"], gold["fcl_loss"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out["no-fcl_loss"], gold["no-fcl_loss"], rtol=1e-5, atol=1e-5)
    for k, v in gold["modality_sample_mask"].items():
        assert torch.equal(out["modality_sample_mask"][k], v)                      # bit-exact
    out["loss"].backward()
    for k, n in gold["grad_norms"].items():
        g = params[k].grad
        got = 0.0 if g is None else g.norm().item()
        assert abs(got - n) <= 2e-4 * max(n, 1e-6) + 1e-7, (k, got, n)
    for k, g in gold["grads"].items():
        torch.testing.assert_close(params[k].grad, g, rtol=2e-4, atol=1e-6 + 1e-5 * float(g.abs().max()))


def test_static_tables_match_reference_hashes():
    static = torch.load(H.GOLDEN_DIR + "/static_tables.pt", weights_only=False)
    from mca_paper_b200 import config as C

    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    for name, g in static.items():
        kw = C.get_model_config(C.named_config(name))
        t = O.static_tables(kw)
        assert sha(t["token_types"]) == g["token_types_sha256"]
        assert sha(t["attn_mask"]) == g["attn_mask_sha256"]
        assert sha(t["pool_mask"]) == g["pool_mask_sha256"]
        assert int((~t["attn_mask"]).sum()) == g["attn_allowed_pairs"]
        assert t["return_token_types"] == g["return_token_types"]
        assert [sorted(c) for c in t["combos"]] == g["fusion_combos"]
        plan, _ = O.loss_plan(kw, t)
        assert [p["name"] for p in plan] == g["loss_names"]


@pytest.mark.skipif(not ref_shim.available(), reason="live reference only exists in the build container")
@pytest.mark.parametrize("case", ["tiny_cmu_fcl_ragged", "tiny_tcga_all_losses", "tiny_cmu_eao"])
def test_oracle_matches_live_reference(case):
    cfg, kw, model, sd, batch = H.build_case(case)
    ref = ref_shim.build_reference_model(kw, state_dict={k: v.clone() for k, v in sd.items()})   # strict: same schema
    t = O.static_tables(kw)
    for name in (("token_types",) if kw.get("eao") else ("token_types", "attn_mask", "pool_mask")):
        assert torch.equal(t[name], getattr(ref, name)), name
    out_ref = ref_shim.reference_forward(ref, batch)
    out = H.oracle_forward(kw)({k: v.clone() for k, v in sd.items()}, kw, batch)
    assert list(out.keys()) == list(out_ref.keys())
    torch.testing.assert_close(out["loss"], out_ref["loss"].detach(), rtol=1e-5, atol=1e-5)


def test_multi_rank_emulation_rows_and_labels():
    """G-rank emulation: with identical batches on both ranks, rank r's positives sit at columns B*r + i."""
    cfg, kw, model, sd, batch = H.build_case("tiny_cmu_fcl_ragged")
    outs = O.mca_forward_ranks(sd, kw, [batch, batch])
    single = O.mca_forward(sd, kw, batch)
    # duplicated columns double every softmax denominator's positive term: loss differs from single-rank,
    # but both ranks (identical data, symmetric labels) must agree with each other
    for k in outs[0]["losses"]:
        a, b = outs[0]["losses"][k], outs[1]["losses"][k]
        assert (torch.isnan(a) and torch.isnan(b)) or torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert not torch.allclose(outs[0]["loss"], single["loss"])


@pytest.mark.skipif(not ref_shim.available(), reason="live reference only exists in the build container")
@pytest.mark.parametrize("kind,variant", [("tcga", "tcga"), ("mixed", "dropout_ragged")])
def test_eao_oracle_matches_live_reference_other_encoders(kind, variant):
    """EAO restatement (oracle eao_forward, model.py:567-596) against the live reference EAO on the TabularEncoder and
    the Sequence / SparseTabular / Patch encoder geometries (the committed fixture covers EmbeddedSequenceEncoder)."""
    from mca_paper_b200 import config as C, synthetic as S
    from mca_paper_b200.model import EAO
    cfg = C.tiny_config(kind, fcl=True, bimodal=True, non_fusion_fcl=True, eao=True)
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in EAO(**kw).state_dict().items()}
    batch = S.make_batch(cfg, seed=2, variant=variant)
    ref = ref_shim.build_reference_model(kw, state_dict={k: v.clone() for k, v in sd.items()})
    ref.eval()  # PatchEncoder's nn.Dropout off (the oracle restates the deterministic part)
    out_ref = ref_shim.reference_forward(ref, batch)
    out = O.eao_forward({k: v.clone() for k, v in sd.items()}, kw, batch)
    assert list(out.keys()) == list(out_ref.keys()) and list(out["losses"]) == list(out_ref["losses"])
    for k, v in out_ref["losses"].items():
        if torch.isnan(v):
            assert torch.isnan(out["losses"][k]), k
        else:
            torch.testing.assert_close(out["losses"][k], v.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["loss"], out_ref["loss"].detach(), rtol=1e-5, atol=1e-5)
