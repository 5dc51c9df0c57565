"""CPU, world_size 2 over gloo: the contrastive loss with its autograd-aware all-gather (the one exchange step of
the path) — validates the single-process multi-rank emulation the GPU parity tests use."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mca_oracle as O


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    a = torch.randn(8, 512, requires_grad=True)
    b = torch.randn(8, 512, requires_grad=True)
    mask = torch.tensor([1, 1, 0, 1, 1, 1, 0, 1], dtype=torch.bool) if rank == 0 else torch.ones(8, dtype=torch.bool)
    loss_fn = O.ContrastiveLossWithTemperature()
    loss = loss_fn(a * 0.05, b * 0.05, mask=mask)
    loss.backward()
    q.put((rank, loss.item(), a.grad.numpy().copy(), b.grad.numpy().copy(), float(loss_fn.logit_scale.grad)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process_emulation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    # emulation: rank r scores its rows against the concatenated columns; embedding grads SUM over ranks
    emb = []
    for rank in range(2):
        torch.manual_seed(100 + rank)
        emb.append((torch.randn(8, 512, requires_grad=True), torch.randn(8, 512, requires_grad=True)))
    masks = [torch.tensor([1, 1, 0, 1, 1, 1, 0, 1], dtype=torch.bool), torch.ones(8, dtype=torch.bool)]
    s = torch.nn.Parameter(torch.tensor(O.math.log(1 / 0.07)))
    a_all = torch.cat([e[0] * 0.05 for e in emb])
    b_all = torch.cat([e[1] * 0.05 for e in emb])
    losses = [O.info_nce(emb[r][0] * 0.05, emb[r][1] * 0.05, s, masks[r], a_all, b_all, rank=r) for r in range(2)]
    sum(losses).backward()
    for r in range(2):
        assert abs(got[r][1] - losses[r].item()) < 1e-5
        torch.testing.assert_close(torch.from_numpy(got[r][2]), emb[r][0].grad, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(torch.from_numpy(got[r][3]), emb[r][1].grad, rtol=1e-4, atol=1e-6)
    # each rank's logit_scale grad is its own loss's; DDP would average them
    assert abs(got[0][4] + got[1][4] - float(s.grad)) < 1e-4 * max(1.0, abs(float(s.grad)))


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mca_paper_b200.utils.metrics import gather_cat
    torch.manual_seed(7 + rank)
    chunks = [torch.randn(3 + rank, 4), torch.randn(2 * rank + 1, 4)]   # ranks hold different numbers of rows
    q.put((rank, gather_cat(chunks).numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_metric_state_is_concatenated_across_ranks_in_rank_order():
    """utils/metrics.py:44-45,64: Alignment / Uniformity states are dist_reduce_fx="cat"."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    want = []
    for rank in range(2):
        torch.manual_seed(7 + rank)
        want += [torch.randn(3 + rank, 4), torch.randn(2 * rank + 1, 4)]
    want = torch.cat(want, 0)
    for r in range(2):
        assert torch.equal(torch.from_numpy(got[r][1]), want)
