"""Eval metrics at eval-set scale: device time of lalign / lunif / retrieval ranks through the C
ABI against the reference formulas run by stock PyTorch on the same GPU (F.normalize + torch.pdist; a dense cosine
matrix instead of the reference's per-row Python loop, which is far slower), and fp32 pair-FLOP rates of the tile kernel."""
import json, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from mca_paper_b200.utils import metrics as M

dev = "cuda"


def dtime(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def torch_lunif(x, t=2):
    return torch.pdist(F.normalize(x), p=2).pow(2).mul(-t).exp().mean().log()


def torch_lalign(x, y, alpha=2):
    return (F.normalize(x) - F.normalize(y)).norm(dim=1).pow(alpha).mean()


def torch_ranks(x, y):
    c = F.normalize(x, eps=1e-8) @ F.normalize(y, eps=1e-8).T
    return (c > c.diagonal()[:, None]).sum(1)


res = {}
for Mrows in (4096, 16384):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(Mrows, 512, generator=g).to(dev)
    y = (x + 0.5 * torch.randn(Mrows, 512, generator=g).to(dev))
    idx = torch.arange(Mrows, device=dev)
    r = {}
    r["lunif_ms"], r["torch_pdist_lunif_ms"] = dtime(lambda: M.lunif(x)), dtime(lambda: torch_lunif(x))
    r["lalign_ms"], r["torch_lalign_ms"] = dtime(lambda: M.lalign(x, y)), dtime(lambda: torch_lalign(x, y))
    r["ranks_ms"], r["torch_dense_cosine_ranks_ms"] = dtime(lambda: M.retrieval_ranks(x, y, idx)), dtime(lambda: torch_ranks(x, y))
    pairs = Mrows * (Mrows - 1) / 2
    r["lunif_fp32_tflops"] = 3 * pairs * 512 / (r["lunif_ms"] * 1e-3) / 1e12      # sub + fma per element
    r["ranks_fp32_tflops"] = 2 * Mrows * Mrows * 512 / (r["ranks_ms"] * 1e-3) / 1e12
    r["lunif_diff_vs_torch"] = abs(float(M.lunif(x)) - float(torch_lunif(x)))
    r["ranks_equal_torch_frac"] = float((M.retrieval_ranks(x, y, idx) == torch_ranks(x, y)).float().mean())
    res[f"M{Mrows}"] = r
print(json.dumps(res))
