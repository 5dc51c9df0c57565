"""Batch producers in front of the hot path: the collate oracle against outputs of the live
reference collators (tests/golden/collate.pt, made by oracle/make_golden.py), and — on a GPU — the device collator against
the oracle, bit for bit."""
import pytest
import torch

from oracle import collate_oracle as CO
from oracle.make_golden import COLLATE_CONFIG, collate_samples
from tests import helpers as H


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].keys() == b[k].keys(), k
        for kk in a[k]:
            x, y = a[k][kk], b[k][kk]
            assert x.shape == y.shape and x.dtype == y.dtype, (k, kk, x.dtype, y.dtype)
            assert torch.equal(x.cpu(), y.cpu()), (k, kk)


@pytest.mark.parametrize("seed", [11, 12])
def test_collate_oracle_matches_reference_golden(seed):
    gold = H.load_golden("collate")
    out = CO.multimodal_collate(COLLATE_CONFIG, collate_samples(seed))
    _same(out, gold[f"collate_{seed}"])


@pytest.mark.parametrize("seed", [11, 12])
def test_predrop_matches_reference_rng_order(seed):
    gold = H.load_golden("collate")
    torch.manual_seed(100 + seed)
    dropped = CO.predrop(collate_samples(seed), COLLATE_CONFIG)
    got = [{k: all(x is None for x in s[k].values()) for k, c in COLLATE_CONFIG.items() if c.get("dropout")} for s in dropped]
    assert got == gold[f"dropped_{seed}"]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_device_collator_bit_exact(seed):
    from mca_paper_b200.collate import DeviceCollator
    samples = collate_samples(seed)
    want = CO.multimodal_collate(COLLATE_CONFIG, samples)
    col = DeviceCollator(COLLATE_CONFIG, batch_size=len(samples), device="cuda")
    got = col(samples)
    torch.cuda.synchronize()
    _same(got, want)
    _same(got, H.load_golden("collate")[f"collate_{seed}"])      # = the live reference's output
    # varlen staging: only live rows crossed PCIe
    live = sum(min(s["speech"]["data"].shape[0], 12) * 5 * 4 for s in samples if s["speech"]["data"] is not None)
    assert col.h2d_bytes < sum(v.numel() * v.element_size() for d in want.values() for v in d.values())
    assert col.h2d_bytes >= live
    # dropout: same decisions as the reference's BatchPreDropout under the same seed, dropped modalities never staged
    torch.manual_seed(100 + seed)
    col2 = DeviceCollator(COLLATE_CONFIG, batch_size=len(samples), device="cuda", apply_dropout=True)
    got2 = col2(samples)
    torch.manual_seed(100 + seed)
    want2 = CO.multimodal_collate(COLLATE_CONFIG, CO.predrop(samples, COLLATE_CONFIG))
    torch.cuda.synchronize()
    _same(got2, want2)


@pytest.mark.gpu
def test_device_collator_feeds_the_model():
    """CMU-shaped ragged samples -> DeviceCollator -> MCA.forward equals the dense synthetic batch path."""
    from mca_paper_b200 import config as C, synthetic as S
    from mca_paper_b200.collate import DeviceCollator
    from mca_paper_b200.model import MCA
    cfg = C.tiny_config("cmu", fcl=True)
    dense = S.make_batch(cfg, seed=1, variant="dropout_ragged")
    mc = {n: {"type": "embedded_sequence", "pad_len": e["max_tokens"], "data_col_name": "data", "pad_token": -10000,
              "embedding_size": e["input_size"]} for n, e in cfg["encoder_configs"].items()}
    B = cfg["batch_size"]
    samples = []
    for b in range(B):
        s = {}
        for n in mc:
            live = int((~dense[n]["attention_mask"][b]).sum())
            s[n] = {"data": dense[n]["tokens"][b, :live].clone() if live else None}
        samples.append(s)
    batch = DeviceCollator(mc, B, "cuda")(samples)
    for n in mc:
        assert torch.equal(batch[n]["tokens"].cpu(), dense[n]["tokens"]) and torch.equal(batch[n]["attention_mask"].cpu(), dense[n]["attention_mask"])
    torch.manual_seed(0)
    model = MCA(**C.get_model_config(cfg)).to("cuda").eval()
    with torch.no_grad():
        a = model(batch)["loss"].item()
        b_ = model(S.batch_to(dense, "cuda"))["loss"].item()
    assert a == b_
