"""TEST INFRASTRUCTURE (checker only; the product never imports this): CPU restatement of the reference's linear probe,
lp_accel_gpu.py:22-35 (FineTuneDataset), :96-117 (loaders, nn.Linear head), :118-167 (loss, AdamW, get_scheduler),
:182-231 (epoch loop).  The script itself cannot be imported (it runs at import and needs accelerate / wandb / torchmetrics),
so its loop is restated here over the SAME stock classes it instantiates — torch.utils.data.DataLoader, nn.Linear,
torch.optim.AdamW, transformers.get_scheduler, the torch.nn loss modules — in the same order, so that the global RNG is
consumed identically.  torchmetrics.PearsonCorrCoef (absent here) is restated as the textbook formula over the epoch's
(prediction, label) stream in float64.  Parity pin: those stock classes ARE the reference's arithmetic for this path."""
import numpy as np
import torch
from torch import nn
from torch.optim import AdamW
from torch.utils.data import DataLoader, Dataset
from transformers import get_scheduler


class FineTuneDataset(Dataset):
    def __init__(self, embeddings, labels, key="fusion", index=0):
        self.embeddings = embeddings[key]
        self.labels = labels if index == -1 else labels[:, index]

    def __len__(self):
        return self.labels.shape[0]

    def __getitem__(self, idx):
        return self.embeddings[idx], self.labels[idx]


def pearson(p, y):
    p, y = np.asarray(p, dtype=np.float64).ravel(), np.asarray(y, dtype=np.float64).ravel()
    pc, yc = p - p.mean(), y - y.mean()
    return float((pc * yc).sum() / np.sqrt((pc * pc).sum() * (yc * yc).sum()))


def probe_fit(e_train, s_train, e_test, s_test, cfg, key="fusion"):
    """cfg keys as utils/config.py:129-153.  Call torch.manual_seed(cfg['seed']) first (lp_accel_gpu.py:52).
    Returns (logs per epoch, weight [L,512], bias [L], visiting order of every epoch)."""
    train_dl = DataLoader(FineTuneDataset(e_train, s_train, key, cfg["task"]), batch_size=cfg["batch_size"], shuffle=True)
    eval_dl = DataLoader(FineTuneDataset(e_test, s_test, key, cfg["task"]), batch_size=cfg["batch_size"])
    e, l = next(iter(train_dl))
    num_labels = l.shape[1] if l.dim() > 1 else 1
    model = nn.Linear(e.shape[1], num_labels)
    loss_fn = {"L1": nn.L1Loss(), "MSE": nn.MSELoss(), "BCE": nn.BCEWithLogitsLoss(), "CE": nn.CrossEntropyLoss()}[cfg["loss_type"]]
    optimizer = AdamW(model.parameters(), lr=cfg["lr"])
    sched = get_scheduler(name=cfg["lr_scheduler_type"], optimizer=optimizer, num_warmup_steps=cfg["num_warmup_steps"],
                          num_training_steps=cfg["epochs"] * len(train_dl))
    logs = []
    for epoch in range(cfg["epochs"]):
        tl, el = 0.0, 0.0
        preds, labs = [], []
        model.train()
        for emb, label in train_dl:
            pred = model(emb).squeeze()
            loss = loss_fn(pred, label)
            optimizer.zero_grad()
            loss.backward()
            preds.append(pred.detach().reshape(-1)), labs.append(label.reshape(-1))
            tl += float(loss.detach())
            if cfg["clip"]:
                gn = torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["clip"])
            optimizer.step()
            sched.step()
        log = {"train_loss": tl / len(train_dl), "lr": optimizer.param_groups[0]["lr"]}
        if cfg["loss_type"] in ("L1", "MSE") and num_labels == 1:
            log["train_PCC"] = pearson(torch.cat(preds).numpy(), torch.cat(labs).numpy())
        preds, labs = [], []
        model.eval()
        with torch.no_grad():
            for emb, label in eval_dl:
                pred = model(emb).squeeze()
                el += float(loss_fn(pred, label))
                preds.append(pred.reshape(-1)), labs.append(label.reshape(-1))
        log["eval_loss"] = el / len(eval_dl)
        if cfg["loss_type"] in ("L1", "MSE") and num_labels == 1:
            log["eval_PCC"] = pearson(torch.cat(preds).numpy(), torch.cat(labs).numpy())
        log["param_norm"] = float(torch.sqrt(sum((p.detach().double() ** 2).sum() for p in model.parameters())))
        logs.append(log)
    return logs, model.weight.detach().clone(), model.bias.detach().clone()
