"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float64) of the reference's embedding-space evaluation metrics
(SURVEY.md §8f rank 4): lalign / lunif / wang_loss (utils/metrics.py:20-33) and the retrieval ranks
compute_cosines / get_rank / get_rank_metrics (utils/metrics.py:73-99).

Only tests/ may import this module; the product path (mca_paper_b200/utils/metrics.py -> csrc/metrics.cu) must never do
so.  Pinned against outputs of the live reference functions frozen in tests/golden/metrics.pt (oracle/make_golden.py,
make_metrics_golden) and, in the build container, against the live functions themselves (tests/test_metrics.py).
"""
from __future__ import annotations

import numpy as np


def _f64(x):
    return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)


def normalize(x, eps=1e-12):
    """torch.nn.functional.normalize(x) (p=2, dim=1): x / max(||x||, eps)."""
    n = np.sqrt((x * x).sum(1, keepdims=True))
    return x / np.maximum(n, eps)


def lalign(x, y, alpha=2, norm=True):
    """utils/metrics.py:20-23."""
    x, y = _f64(x), _f64(y)
    if norm:
        x, y = normalize(x), normalize(y)
    if x.shape[0] == 0:
        return float("nan")
    return float((np.sqrt(((x - y) ** 2).sum(1)) ** alpha).mean())


def lunif(x, t=2, norm=True):
    """utils/metrics.py:26-29: torch.pdist = the M(M-1)/2 distances of rows i < j."""
    x = _f64(x)
    if norm:
        x = normalize(x)
    M = x.shape[0]
    if M < 2:
        return float("nan")
    tot = 0.0
    for i in range(M - 1):
        d2 = ((x[i + 1:] - x[i]) ** 2).sum(1)
        # the reference evaluates exp in fp32: terms below ~exp(-103.97) underflow to 0 (-> log 0 = -inf for spread-out,
        # unnormalised embeddings); keep that semantic, sum in float64
        tot += np.exp((-t * d2).astype(np.float32)).astype(np.float64).sum()
    with np.errstate(divide="ignore"):
        return float(np.log(tot / (M * (M - 1) / 2)))


def wang_loss(x, y, lam=1.0, alpha=2, t=2):
    """utils/metrics.py:32-33."""
    return lalign(x, y, alpha) + lam * (lunif(x, t) + lunif(y, t)) / 2


def compute_cosines(embedding, embeddings, eps=1e-8):
    """utils/metrics.py:73-76: nn.CosineSimilarity(dim=1) of one row against every row."""
    a = embedding / max(np.sqrt((embedding * embedding).sum()), eps)
    b = embeddings / np.maximum(np.sqrt((embeddings * embeddings).sum(1, keepdims=True)), eps)
    return b @ a


def get_rank(x, indices):
    """utils/metrics.py:78-80."""
    vals = x[np.arange(len(x)), indices]
    return (x > vals[:, None]).sum(1)


def cosine_matrix(embeddings, mask, targets):
    e, y = _f64(embeddings), _f64(targets)
    idx = [i for i in range(e.shape[0]) if bool(mask[i])]
    return np.stack([compute_cosines(e[i], y) for i in idx]), np.asarray(idx)


def get_rank_metrics(embeddings, mask, targets):
    """utils/metrics.py:82-99 -> (ranks, median_rank, r1, r5, r10); the median is torch's lower median."""
    c, idx = cosine_matrix(embeddings, mask, targets)
    ranks = get_rank(c, idx)
    med = int(np.sort(ranks)[(len(ranks) - 1) // 2])
    n = len(ranks)
    return ranks, med, (ranks == 0).sum() / n, (ranks < 5).sum() / n, (ranks < 10).sum() / n


def near_ties(embeddings, mask, targets, tol=2e-6):
    """Per masked row: how many other targets score within `tol` of the row's own cosine — the fp32 kernels (and the
    fp32 reference itself) may fall either side of those, so a parity test allows that many rank steps."""
    c, idx = cosine_matrix(embeddings, mask, targets)
    own = c[np.arange(len(c)), idx]
    close = np.abs(c - own[:, None]) <= tol
    close[np.arange(len(c)), idx] = False
    return close.sum(1)
