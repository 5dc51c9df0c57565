"""Embedding-space evaluation metrics with the reference's names and argument meaning (utils/metrics.py): `lalign`,
`lunif`, `wang_loss` (:20-33), the `Alignment` / `Uniformity` accumulators (:37-70; torchmetrics is not required — the
classes keep its update / compute / reset / __call__ protocol) and `get_rank_metrics` (:73-99).  The arithmetic runs in
the CUDA kernels of csrc/metrics.cu through the C ABI; inputs must live on the GPU (no CPU fallback — the product
fails loudly instead).  Results stay on the device; nothing here synchronises except the index check of
`get_rank_metrics` (the reference raises IndexError at the same place).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _lib
from ..ops import P, S, call


def _dev2d(x: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or x.dim() != 2:
        raise ValueError(f"{what}: expected a [rows, dim] tensor")
    if x.device.type != "cuda":
        raise _lib.MCAKernelError(f"{what}: mca_paper_b200 metrics run on CUDA only (no CPU fallback)")
    return x.detach().to(torch.float32).contiguous()


def _scratch(M: int, dev) -> torch.Tensor:
    tiles = (M + 63) // 64
    return torch.empty(max(tiles * tiles, (M + 7) // 8, 1), device=dev, dtype=torch.float64)


def lalign(x, y, alpha=2, norm=True):
    """mean_i ||x_i - y_i||^alpha over positive pairs (utils/metrics.py:20-23)."""
    x, y = _dev2d(x, "lalign x"), _dev2d(y, "lalign y")
    if x.shape != y.shape:
        raise ValueError("preds and target must have the same shape")
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    scratch = _scratch(x.shape[0], x.device)
    call("mca_alignment", P(x), P(y), x.shape[0], x.shape[1], float(alpha), int(bool(norm)), P(scratch), P(out), S())
    return out[0]


def lunif(x, t=2, norm=True):
    """log mean_{i<j} exp(-t ||x_i - x_j||^2) (utils/metrics.py:26-29)."""
    x = _dev2d(x, "lunif x")
    M = x.shape[0]
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    inv = torch.empty(max(M, 1), device=x.device, dtype=torch.float32)
    scratch = _scratch(M, x.device)
    call("mca_uniformity", P(x), M, x.shape[1], float(t), int(bool(norm)), P(inv), P(scratch), P(out), S())
    return out[0]


def wang_loss(x, y, lam=1.0, alpha=2, t=2):
    """utils/metrics.py:32-33."""
    return lalign(x, y, alpha) + lam * (lunif(x, t) + lunif(y, t)) / 2


def gather_cat(chunks, dim0_width=None):
    """torchmetrics' `dist_reduce_fx="cat"` (utils/metrics.py:44-45,64): the state a metric computes on is the
    concatenation, in rank order, of every rank's accumulated rows.  Outside a process group (or with one rank) it is the
    local concatenation.  Ranks may hold different numbers of rows: the counts are exchanged first and the rows padded to
    the longest for one all_gather."""
    local = torch.cat(list(chunks), 0) if len(chunks) else None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        if local is None:
            raise ValueError("metric has no accumulated state")
        return local
    world = dist.get_world_size()
    if local is None:
        raise ValueError("metric has no accumulated state on this rank")
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    padded = local.new_zeros((mx,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


class _Accumulator:
    """The slice of the torchmetrics.Metric protocol the reference scripts use: update(), compute(), reset() and
    calling the object on one batch (= update + compute on that batch alone, state kept).  compute() evaluates on the
    rows of EVERY rank when a process group is initialised (the reference's states are `dist_reduce_fx="cat"`); the
    per-batch value of a call is local, as torchmetrics' forward is with dist_sync_on_step=False."""

    _states = ()

    def __init__(self):
        self._sync = True
        self.reset()

    def _state(self, name):
        chunks = getattr(self, name)
        return gather_cat(chunks) if self._sync else torch.cat(chunks, 0)

    def reset(self):
        for s in self._states:
            setattr(self, s, [])

    def to(self, *a, **k):
        return self

    def __call__(self, *args):
        keep = {s: getattr(self, s) for s in self._states}
        for s in self._states:
            setattr(self, s, [])
        self.update(*args)
        self._sync = False
        try:
            val = self.compute()
        finally:
            self._sync = True
        for s in self._states:
            setattr(self, s, keep[s] + getattr(self, s))
        return val


class Alignment(_Accumulator):
    """utils/metrics.py:37-55: concatenates every update and evaluates lalign on the whole set (`norm` defaults to
    False in compute(), as in the reference)."""
    _states = ("preds", "target")

    def __init__(self, alpha=2, **kwargs):
        self.alpha = alpha
        super().__init__()

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        self.preds.append(preds)
        self.target.append(target)
        if preds.shape != target.shape:
            raise ValueError("preds and target must have the same shape")

    def compute(self, norm=False):
        return lalign(self._state("preds"), self._state("target"), self.alpha, norm)


class Uniformity(_Accumulator):
    """utils/metrics.py:58-70."""
    _states = ("preds",)

    def __init__(self, t=2, **kwargs):
        self.t = t
        super().__init__()

    def update(self, preds: torch.Tensor) -> None:
        self.preds.append(preds)

    def compute(self, norm=False):
        return lunif(self._state("preds"), self.t, norm)


def retrieval_ranks(embeddings, targets, indices):
    """ranks[i] = number of targets whose cosine with embeddings[i] exceeds that of targets[indices[i]]
    (compute_cosines + get_rank, utils/metrics.py:73-80)."""
    e, t = _dev2d(embeddings, "retrieval_ranks embeddings"), _dev2d(targets, "retrieval_ranks targets")
    if e.shape[1] != t.shape[1]:
        raise ValueError("embeddings and targets must have the same width")
    idx = torch.as_tensor(indices, device=e.device, dtype=torch.int64).contiguous()
    M, T = e.shape[0], t.shape[0]
    if idx.numel() != M:
        raise ValueError("one target index per embedding row")
    if M and (int(idx.max()) >= T or int(idx.min()) < -T):
        raise IndexError(f"index {int(idx.max())} is out of bounds for dimension 1 with size {T}")
    idx = torch.where(idx < 0, idx + T, idx)
    ranks = torch.zeros(M, device=e.device, dtype=torch.int64)
    if M == 0:
        return ranks
    inv_e, inv_t, own = (torch.empty(n, device=e.device, dtype=torch.float32) for n in (M, T, M))  # distinct live buffers
    call("mca_retrieval_ranks", P(e), P(t), P(idx), M, T, e.shape[1], P(inv_e), P(inv_t), P(own), P(ranks), S())
    return ranks


def get_rank_metrics(embeddings, mask, targets, fusion="fusion", device="cuda"):
    """(median_rank, r1, r5, r10) over the rows of `embeddings` whose `mask` is set; row i's own target is
    targets[i] (utils/metrics.py:82-99; `fusion` is unused there too)."""
    embeddings = embeddings.to(device)
    targets = targets.to(device)
    mask = torch.as_tensor(mask, device=embeddings.device).to(torch.bool).reshape(-1)
    idx = torch.nonzero(mask[:embeddings.shape[0]], as_tuple=False).reshape(-1)
    ranks = retrieval_ranks(embeddings[idx], targets, idx)
    n = ranks.numel()
    median_rank = ranks.median()          # lower median of an int64 tensor, like the reference
    r1 = (ranks == 0).sum() / n
    r5 = (ranks < 5).sum() / n
    r10 = (ranks < 10).sum() / n
    return median_rank, r1, r5, r10
