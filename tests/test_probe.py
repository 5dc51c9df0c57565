"""Linear probe (lp_accel_gpu.py): host logic on CPU (head initialisation and visiting order under the reference's RNG
consumption) and, -m gpu, the whole training against the oracle restatement."""
import pytest
import torch

from mca_paper_b200.linear_probe import FineTuneDataset, LinearProbe
from oracle import probe_oracle as PO


def _data(n_train=700, n_eval=180, n_labels=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(512, n_labels, generator=g) * 0.05
    e_tr = {"fusion": torch.randn(n_train, 512, generator=g)}
    e_ev = {"fusion": torch.randn(n_eval, 512, generator=g)}
    s_tr = e_tr["fusion"] @ w + 0.3 * torch.randn(n_train, n_labels, generator=g)
    s_ev = e_ev["fusion"] @ w + 0.3 * torch.randn(n_eval, n_labels, generator=g)
    return e_tr, s_tr, e_ev, s_ev


CFG = dict(task=0, loss_type="L1", lr=2e-3, lr_scheduler_type="cosine", num_warmup_steps=5, epochs=3, clip=2.0,
           batch_size=64, seed=42)


def test_probe_host_logic_consumes_rng_like_the_reference():
    e_tr, s_tr, e_ev, s_ev = _data()
    torch.manual_seed(CFG["seed"])
    probe = LinearProbe(FineTuneDataset(e_tr, s_tr, index=0), FineTuneDataset(e_ev, s_ev, index=0), device="cpu", **CFG)
    order0 = torch.cat([b for b in probe.train_dl])
    for _ in probe.eval_dl:
        pass
    order1 = torch.cat([b for b in probe.train_dl])
    # the reference's classes, iterated the way lp_accel_gpu.py iterates them
    from torch import nn
    from torch.utils.data import DataLoader
    torch.manual_seed(CFG["seed"])
    tr = PO.FineTuneDataset(e_tr, s_tr, "fusion", 0)
    train_dl = DataLoader(tr, batch_size=CFG["batch_size"], shuffle=True)
    eval_dl = DataLoader(PO.FineTuneDataset(e_ev, s_ev, "fusion", 0), batch_size=CFG["batch_size"])
    e, l = next(iter(train_dl))
    head = nn.Linear(512, 1)
    assert torch.equal(probe.weight, head.weight.detach()) and torch.equal(probe.bias, head.bias.detach())
    seen = torch.cat([lab for _, lab in train_dl])
    assert torch.equal(seen, tr.labels[order0])          # epoch 0 visits the rows in the same order
    for _ in eval_dl:
        pass
    seen = torch.cat([lab for _, lab in train_dl])
    assert torch.equal(seen, tr.labels[order1])
    assert probe.n_out == 1 and len(probe.train_dl) == 11
    with pytest.raises(NotImplementedError):
        LinearProbe(FineTuneDataset(e_tr, s_tr), FineTuneDataset(e_ev, s_ev), device="cpu", model_type="mlp")
    with pytest.raises(Exception):
        LinearProbe(FineTuneDataset(e_tr, s_tr), FineTuneDataset(e_ev, s_ev), device="cpu", loss_type="huber")


@pytest.mark.parametrize("sched", ["cosine", "linear", "constant_with_warmup", "constant"])
def test_probe_lr_schedule_matches_transformers(sched):
    """lr_at(k) (host mirror of the schedule inside mca_probe_epoch) == the rate transformers.get_scheduler gives the k-th
    optimiser step (lp_accel_gpu.py:161-167: scheduler.step() after every optimizer.step())."""
    from transformers import get_scheduler
    e_tr, s_tr, e_ev, s_ev = _data(n_train=200, n_eval=50)
    cfg = dict(CFG, lr_scheduler_type=sched, epochs=5, num_warmup_steps=4, lr=3e-3)
    probe = LinearProbe(FineTuneDataset(e_tr, s_tr, index=0), FineTuneDataset(e_ev, s_ev, index=0), device="cpu", **cfg)
    total = cfg["epochs"] * len(probe.train_dl)
    opt = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=cfg["lr"])
    sch = get_scheduler(name=sched, optimizer=opt, num_warmup_steps=4, num_training_steps=total)
    for k in range(1, total + 1):
        assert abs(probe.lr_at(k) - opt.param_groups[0]["lr"]) <= 1e-9 + 1e-6 * cfg["lr"], (sched, k)
        opt.step()
        sch.step()


@pytest.mark.gpu
@pytest.mark.parametrize("loss_type,task,sched", [("L1", 0, "cosine"), ("MSE", 2, "linear"), ("MSE", -1, "constant_with_warmup"),
                                                  ("BCE", 1, "cosine"), ("CE", -1, "constant")])
def test_probe_training_matches_oracle(loss_type, task, sched):
    e_tr, s_tr, e_ev, s_ev = _data()
    if loss_type == "BCE":
        s_tr, s_ev = (s_tr > 0).float(), (s_ev > 0).float()
    if loss_type == "CE":
        s_tr, s_ev = s_tr.softmax(dim=1), s_ev.softmax(dim=1)
    cfg = dict(CFG, loss_type=loss_type, task=task, lr_scheduler_type=sched)
    torch.manual_seed(cfg["seed"])
    logs_ref, w_ref, b_ref = PO.probe_fit(e_tr, s_tr, e_ev, s_ev, cfg)
    torch.manual_seed(cfg["seed"])
    probe = LinearProbe(FineTuneDataset(e_tr, s_tr, index=task), FineTuneDataset(e_ev, s_ev, index=task), device="cuda", **cfg)
    logs = probe.fit()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.cpu().double() - b.double()).norm() / b.double().norm())
    assert rel(probe.weight, w_ref) < 2e-4 and rel(probe.bias, b_ref) < 2e-3, (rel(probe.weight, w_ref), rel(probe.bias, b_ref))
    for got, want in zip(logs, logs_ref):
        for k in ("train_loss", "eval_loss", "param_norm", "lr"):
            assert abs(float(got[k]) - want[k]) <= 2e-4 * abs(want[k]) + 1e-9, (k, float(got[k]), want[k])
        if "train_PCC" in want:
            assert abs(float(got["train_PCC"]) - want["train_PCC"]) < 2e-4
            assert abs(float(got["eval_PCC"]) - want["eval_PCC"]) < 2e-4
    # predict() == the trained head
    x = e_ev["fusion"][:50]
    want = (x @ w_ref.t() + b_ref).squeeze()
    assert rel(probe.predict(x), want) < 2e-4
