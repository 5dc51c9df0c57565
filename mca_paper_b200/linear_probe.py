"""Linear probe on frozen embeddings — the host mirror of the reference's lp_accel_gpu.py on the device kernels.

Reference: lp_accel_gpu.py:22-35 (`FineTuneDataset`), :96-117 (data loaders, `nn.Linear(num_emb, num_labels)` head),
:118-157 (loss / metric selection), :160-167 (AdamW + `get_scheduler`), :182-231 (epoch loop: forward, loss, backward,
`clip_grad_norm_`, `optimizer.step()`, `lr_scheduler.step()` per batch; evaluation pass; the logged dict);
defaults utils/config.py:129-153.

The embeddings and labels stay resident on the device; ONE kernel launch (`mca_probe_epoch`, csrc/probe.cu: a thread-block
cluster holding parameters and AdamW moments in shared memory) runs a whole epoch of mini-batch steps, a second launch the
evaluation pass, a third the Pearson correlation.  The host only supplies the visiting order — from a real
`torch.utils.data.DataLoader(shuffle=True)` over the row indices, iterated exactly like the reference iterates its loaders,
so that the global RNG is consumed identically (head initialisation and every epoch's permutation match the reference
under the same `torch.manual_seed`).

Not covered: `model_type: mlp` (its nn.Dropout stream cannot be reproduced) and the torchmetrics classification suite of
the BCE / CE branches (precision, recall, AUROC ...): only the losses and, for single-output regression, PearsonCorrCoef.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch
from torch import nn
from torch.utils.data import DataLoader, Dataset

from . import ops
from .engine import LR_MODES
from .ops import P, S, call

LOSS_KINDS = {"L1": 0, "MSE": 1, "BCE": 2, "CE": 3}

# utils/config.py:129-153
PROBE_DEFAULTS = {"task": 0, "loss_type": "L1", "model_type": "linear", "hidden_size": 256, "dropout": 0.1, "lr": 1e-5,
                  "lr_scheduler_type": "cosine", "num_warmup_steps": 1000, "rank_metrics": True, "epochs": 1024, "clip": 2.0,
                  "metric": "PCC", "seed": 42, "batch_size": 1024, "threshold": 0.0}


class FineTuneDataset(Dataset):
    """lp_accel_gpu.py:22-35: (embeddings[key][i], labels[i, index]) pairs; index == -1 keeps every label column."""

    def __init__(self, embeddings, labels, key="fusion", index=0, transform=None, target_transform=None):
        self.embeddings = embeddings[key]
        self.labels = labels if index == -1 else labels[:, index]
        self.transform = transform
        self.target_transform = target_transform

    def __len__(self):
        return self.labels.shape[0]

    def __getitem__(self, idx):
        return self.embeddings[idx], self.labels[idx]


class _Rows(Dataset):
    """Row indices only: the loaders below yield the visiting ORDER; the rows themselves never leave the device."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class LinearProbe:
    """probe = LinearProbe(train_ds, eval_ds, loss_type="L1", lr=..., ...); logs = probe.fit()  (one dict per epoch, the
    keys the reference logs at lp_accel_gpu.py:221-226)."""

    def __init__(self, train_ds: FineTuneDataset, eval_ds: FineTuneDataset, device="cuda", **config):
        cfg = dict(PROBE_DEFAULTS)
        cfg.update(config)
        self.cfg = cfg
        if str(cfg["model_type"]).lower() != "linear":
            raise NotImplementedError("model_type 'mlp' (nn.Dropout between the layers) is not built; use 'linear'")
        if cfg["loss_type"] not in LOSS_KINDS:
            raise Exception("Didn't recognize config.metric")                       # lp_accel_gpu.py:150
        if cfg["lr_scheduler_type"] not in LR_MODES:
            raise ValueError(f"unknown lr schedule {cfg['lr_scheduler_type']!r}")
        self.device = torch.device(device)
        B = int(cfg["batch_size"])
        # lp_accel_gpu.py:96-97: the loaders (here over row indices; same sampler classes, same RNG consumption)
        self.train_dl = DataLoader(_Rows(len(train_ds)), batch_size=B, shuffle=True)
        self.eval_dl = DataLoader(_Rows(len(eval_ds)), batch_size=B)
        # lp_accel_gpu.py:99-104: one batch is drawn to read the shapes, THEN the head is initialised
        first = next(iter(self.train_dl))
        l0 = train_ds.labels[first]
        self.n_out = int(l0.shape[1]) if l0.dim() > 1 else 1
        self.n_emb = int(train_ds.embeddings.shape[1])
        if self.n_emb != 512 or self.n_out > 8:
            raise AssertionError("the probe kernel is built for 512-wide embeddings and at most 8 outputs")
        if cfg["loss_type"] == "CE" and self.n_out < 2:
            raise Exception("CrossEntropyLoss needs class-probability targets [n, C] (task: -1)")
        head = nn.Linear(self.n_emb, self.n_out)
        Pn = self.n_out * 512 + self.n_out
        dev = self.device
        self.state = torch.zeros(3, Pn, device=dev, dtype=torch.float32)
        self.state[0, :self.n_out * 512] = head.weight.detach().reshape(-1).to(dev)
        self.state[0, self.n_out * 512:] = head.bias.detach().to(dev)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.x_train, self.x_eval = f32(train_ds.embeddings), f32(eval_ds.embeddings)
        self.y_train = f32(train_ds.labels).view(len(train_ds), self.n_out)
        self.y_eval = f32(eval_ds.labels).view(len(eval_ds), self.n_out)
        self.pred_train, self.pred_eval = torch.zeros_like(self.y_train), torch.zeros_like(self.y_eval)
        self.loss_sum = torch.zeros(2, device=dev, dtype=torch.float64)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.pcc = torch.zeros(2, device=dev, dtype=torch.float32)
        self.eval_order = torch.arange(len(eval_ds), device=dev, dtype=torch.int32)
        steps = int(cfg["epochs"]) * len(self.train_dl)                                       # lp_accel_gpu.py:160
        clip = float(cfg["clip"]) if cfg["clip"] else 0.0                                      # `if config.clip:` :199
        self.adamw = ops.AdamWCfg(float(cfg["lr"]), 0.9, 0.999, 1e-8, 0.01, clip, LR_MODES[cfg["lr_scheduler_type"]],
                                  int(cfg["num_warmup_steps"]), steps, 1)
        self.loss_kind = LOSS_KINDS[cfg["loss_type"]]
        self.epoch = 0

    # ---- parameters in nn.Linear layout
    @property
    def weight(self):
        return self.state[0, :self.n_out * 512].view(self.n_out, 512)

    @property
    def bias(self):
        return self.state[0, self.n_out * 512:]

    def lr_at(self, step: int) -> float:
        """transformers.get_scheduler value the `step`-th optimiser step (1-based) uses."""
        c = self.adamw
        if c.lr_mode == 0:
            return float(c.lr)
        cur = float(step - 1)
        if cur < c.warmup_steps:
            return float(c.lr) * cur / max(1.0, float(c.warmup_steps))
        if c.lr_mode == 2:
            return float(c.lr)
        if c.lr_mode == 3:
            return float(c.lr) * max(0.0, (c.total_steps - cur) / max(1.0, float(c.total_steps - c.warmup_steps)))
        prog = (cur - c.warmup_steps) / max(1.0, float(c.total_steps - c.warmup_steps))
        return float(c.lr) * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))

    def _launch(self, x, y, order, n, train, pred, loss_slot):
        call("mca_probe_epoch", P(x), P(y), P(order), n, int(self.cfg["batch_size"]), self.n_out, self.loss_kind, int(train),
             P(self.state), P(self.step_dev), ctypes.addressof(self.adamw), P(pred), P(self.loss_sum[loss_slot:]),
             P(self.grad_norm), S())

    def train_epoch(self) -> Dict[str, torch.Tensor]:
        """One pass of lp_accel_gpu.py:183-231; returns the logged dict (device tensors: nothing is synchronised)."""
        order = torch.cat([b for b in self.train_dl]).to(torch.int32)          # this epoch's permutation (host RNG)
        order_dev = order.to(self.device, non_blocking=False)
        self.loss_sum.zero_()
        n_tr, n_ev = order.numel(), self.eval_order.numel()
        self._launch(self.x_train, self.y_train, order_dev, n_tr, True, self.pred_train, 0)
        for _ in self.eval_dl:                                                   # consumes the loader's base seed like :208
            pass
        self._launch(self.x_eval, self.y_eval, self.eval_order, n_ev, False, self.pred_eval, 1)
        out = {"train_loss": (self.loss_sum[0] / len(self.train_dl)).float(),   # epoch_loss / len(dl), :221-222
               "eval_loss": (self.loss_sum[1] / len(self.eval_dl)).float(),
               # optimizer.param_groups[0]["lr"] after the epoch's last scheduler.step(): the rate the NEXT step would use
               "lr": self.lr_at(int(self.epoch + 1) * len(self.train_dl) + 1),
               "param_norm": self.state[0].double().norm(), "grad_norm": self.grad_norm[0].clone()}
        if self.cfg["loss_type"] in ("L1", "MSE") and self.n_out == 1:
            call("mca_probe_pcc", P(self.pred_train), P(self.y_train), n_tr, P(self.pcc), S())
            call("mca_probe_pcc", P(self.pred_eval), P(self.y_eval), n_ev, P(self.pcc[1:]), S())
            out["train_PCC"], out["eval_PCC"] = self.pcc[0].clone(), self.pcc[1].clone()
        self.epoch += 1
        return out

    def fit(self, epochs: Optional[int] = None):
        return [self.train_epoch() for _ in range(int(self.cfg["epochs"] if epochs is None else epochs))]

    def predict(self, embeddings: torch.Tensor) -> torch.Tensor:
        """model(embedding).squeeze() of the trained head (lp_accel_gpu.py:191) for [n, 512] embeddings on the device."""
        x = embeddings.detach().to(self.device, torch.float32).contiguous()
        n = x.shape[0]
        pred = torch.empty(n, self.n_out, device=self.device)
        scratch = torch.zeros(1, device=self.device, dtype=torch.float64)
        zeros = torch.zeros(n, self.n_out, device=self.device)
        order = torch.arange(n, device=self.device, dtype=torch.int32)
        call("mca_probe_epoch", P(x), P(zeros), P(order), n,
             int(self.cfg["batch_size"]), self.n_out, self.loss_kind, 0, P(self.state), P(self.step_dev),
             ctypes.addressof(self.adamw), P(pred), P(scratch), None, S())
        torch.cuda.current_stream().synchronize()   # the temporaries above must outlive the launch
        return pred.squeeze()
