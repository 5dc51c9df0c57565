"""Embedding-space evaluation metrics (utils/metrics.py:20-99).  CPU: the numpy oracle against the
golden outputs of the live reference functions (tests/golden/metrics.pt, oracle/make_golden.py make_metrics_golden) and,
in the build container, against the live functions.  GPU: the CUDA kernels through the C ABI against the golden
outputs, the oracle on other shapes, and size-independent properties."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO, ref_shim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.pt")
ATOL, RTOL = 2e-6, 2e-5   # fp32 reference vs float64 oracle / fp32 kernels with double accumulation


def close(a, b, atol=ATOL, rtol=RTOL):
    a, b = float(a), float(b)
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= atol + rtol * abs(b)


def cases(inp):
    """(golden key, function name, args) of every scalar in the fixture."""
    x, y = inp["x"], inp["y"]
    out = []
    for norm in (True, False):
        for alpha in (2, 1, 3.5):
            out.append((f"lalign_a{alpha}_n{int(norm)}", "lalign", (x, y, alpha, norm)))
        for t in (2, 0.5):
            out.append((f"lunif_t{t}_n{int(norm)}", "lunif", (x, t, norm)))
            out.append((f"lunif_collapsed_t{t}_n{int(norm)}", "lunif", (inp["collapsed"], t, norm)))
            out.append((f"lunif_small_t{t}_n{int(norm)}", "lunif", (inp["small"], t, norm)))
    out.append(("wang", "wang_loss", (x, y)))
    return out


def test_oracle_matches_reference_golden():
    g = torch.load(GOLD)
    inp, want = g["inputs"], g["outputs"]
    for key, fn, args in cases(inp):
        assert close(getattr(MO, fn)(*args), want[key]), key
    for name, tg in (("easy", inp["y"]), ("hard", inp["yr"])):
        ranks, med, r1, r5, r10 = MO.get_rank_metrics(inp["x"], inp["mask"], tg)
        slack = MO.near_ties(inp["x"], inp["mask"], tg)
        assert (np.abs(ranks - want[f"ranks_{name}"].numpy()) <= slack).all()
        if slack.sum() == 0:
            assert np.allclose([med, r1, r5, r10], want[f"rank_metrics_{name}"].numpy(), atol=1e-7)


@pytest.mark.skipif(not ref_shim.available(), reason="live reference only exists in the build container")
def test_oracle_matches_live_reference():
    R = ref_shim.load_reference_metrics()
    g = torch.Generator().manual_seed(3)
    x, y = torch.randn(37, 96, generator=g), torch.randn(37, 96, generator=g)
    mask = torch.rand(37, generator=g) > 0.5
    for norm in (True, False):
        assert close(MO.lalign(x, y, 2, norm), R.lalign(x, y, 2, norm))
        assert close(MO.lunif(0.1 * x, 2, norm), R.lunif(0.1 * x, 2, norm))
    med, r1, r5, r10 = R.get_rank_metrics(x, mask, y, device="cpu")
    ranks, omed, o1, o5, o10 = MO.get_rank_metrics(x, mask, y)
    if MO.near_ties(x, mask, y).sum() == 0:
        assert omed == int(med) and np.allclose([o1, o5, o10], [float(r1), float(r5), float(r10)], atol=1e-6)


def test_host_wrappers_refuse_cpu_tensors():
    from mca_paper_b200 import _lib
    from mca_paper_b200.utils import metrics as M
    with pytest.raises(_lib.MCAKernelError):
        M.lalign(torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(_lib.MCAKernelError):
        M.lunif(torch.zeros(4, 8))


# ----------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_kernels_match_reference_golden():
    from mca_paper_b200.utils import metrics as M
    dev = torch.device("cuda:0")
    g = torch.load(GOLD)
    inp = {k: v.to(dev) for k, v in g["inputs"].items()}
    want = g["outputs"]
    for key, fn, args in cases(inp):
        got = getattr(M, fn)(*args)
        assert got.device.type == "cuda" and close(got, want[key]), (key, float(got), float(want[key]))
    # accumulators: three updates == one evaluation over the concatenation; calling the object scores one batch alone
    al, un = M.Alignment(), M.Uniformity()
    for a, b in zip(inp["x"].chunk(3), inp["y"].chunk(3)):
        al.update(a, b)
        un.update(a)
    assert close(al.compute(), want["Alignment"]) and close(un.compute(), want["Uniformity"])
    assert close(al.compute(norm=True), want["Alignment_norm"]) and close(un.compute(norm=True), want["Uniformity_norm"])
    one = M.Alignment()(inp["x"][:10], inp["y"][:10])
    assert close(one, MO.lalign(inp["x"][:10], inp["y"][:10], 2, False))
    al.reset()
    assert al.preds == [] and al.target == []
    with pytest.raises(ValueError):
        al.update(inp["x"][:4], inp["y"][:5])
    for name, tg in (("easy", inp["y"]), ("hard", inp["yr"])):
        med, r1, r5, r10 = M.get_rank_metrics(inp["x"], inp["mask"], tg)
        idx = torch.nonzero(inp["mask"]).reshape(-1)
        ranks = M.retrieval_ranks(inp["x"][idx], tg, idx).cpu().numpy()
        slack = MO.near_ties(g["inputs"]["x"], g["inputs"]["mask"], g["inputs"]["y" if name == "easy" else "yr"])
        assert (np.abs(ranks - want[f"ranks_{name}"].numpy()) <= slack).all()
        if slack.sum() == 0:
            got = [float(med), float(r1), float(r5), float(r10)]
            assert np.allclose(got, want[f"rank_metrics_{name}"].numpy(), atol=1e-6), (name, got)


@pytest.mark.gpu
@pytest.mark.parametrize("M_,D", [(1, 512), (2, 7), (63, 33), (64, 512), (65, 100), (200, 512), (777, 64)])
def test_kernels_match_oracle_ragged_shapes(M_, D):
    """Row counts around the 64-row tile and widths that are not multiples of the 32-column chunk."""
    from mca_paper_b200.utils import metrics as M
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M_ * 1000 + D)
    x, y = torch.randn(M_, D, generator=g), torch.randn(M_, D, generator=g)
    x[0] = 0                                   # a zero row: F.normalize / cosine eps paths
    s = 0.5 / math.sqrt(D)                     # keep exp(-t d^2) away from underflow
    for norm in (True, False):
        assert close(M.lalign(x.to(dev), y.to(dev), 2, norm), MO.lalign(x, y, 2, norm))
        assert close(M.lalign(x.to(dev), y.to(dev), 1.5, norm), MO.lalign(x, y, 1.5, norm))
        assert close(M.lunif((s * x).to(dev), 2, norm), MO.lunif(s * x, 2, norm)), (M_, D, norm)
    mask = torch.rand(M_, generator=g) > 0.4
    mask[0] = True
    if M_ >= 2:
        mask[1] = True
    tg = 0.2 * x + y
    idx = torch.nonzero(mask).reshape(-1)
    ranks = M.retrieval_ranks(x[idx].to(dev), tg.to(dev), idx.to(dev)).cpu().numpy()
    want, *_ = MO.get_rank_metrics(x, mask, tg)
    assert (np.abs(ranks - want) <= MO.near_ties(x, mask, tg)).all()


@pytest.mark.gpu
def test_metric_properties_at_eval_scale():
    """Size-independent properties on an eval-set-sized input (M = 5000 pooled embeddings of width 512): identical
    sets align perfectly, uniformity is invariant under row permutation and bit-reproducible, a set against itself
    retrieves every sample at rank 0, and out-of-range indices raise like the reference's get_rank."""
    from mca_paper_b200.utils import metrics as M
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(5000, 512, generator=g).to(dev)
    assert float(M.lalign(x, x.clone())) == 0.0
    u1, u2 = M.lunif(x), M.lunif(x)
    assert float(u1) == float(u2) and -4.2 < float(u1) < -3.8      # random unit vectors: E||xi-xj||^2 = 2 -> about -4
    perm = torch.randperm(5000, generator=g).to(dev)
    assert abs(float(M.lunif(x[perm])) - float(u1)) < 1e-5
    med, r1, r5, r10 = M.get_rank_metrics(x, torch.ones(5000, dtype=torch.bool), x)
    assert int(med) == 0 and float(r1) == 1.0 and float(r5) == 1.0 and float(r10) == 1.0
    assert math.isnan(float(M.lunif(x[:1]))) and math.isnan(float(M.lalign(x[:0], x[:0])))
    with pytest.raises(IndexError):
        M.retrieval_ranks(x[:4], x[:3], torch.tensor([0, 1, 2, 3]))
