"""Synthetic batches in the collators' layout (there is no dataset access): shapes and value conventions follow
SURVEY.md §8(d) and the reference collators encoders.py:286-343 / dropout utils/dataset.py:29-57.

  embedded_sequence (CMU):  {'tokens': f32 [B,L,in], 'attention_mask': bool [B,L]}   True = padded (encoders.py:339-343)
  sequence/tabular (TCGA):  {'values': f32 [B,L],    'attention_mask': int64 [B,L]}  1 = padded  (encoders.py:303-311)
"""
from __future__ import annotations

from typing import Dict

import torch

PAD_VALUE = -10000.0


def make_batch(config, seed: int = 1, variant: str = "full", batch_size: int | None = None,
               p_absent: float | None = None) -> Dict[str, Dict[str, torch.Tensor]]:
    """variant: 'full' (no padding), 'dropout_full' (absent modalities, present ones full length),
    'dropout_ragged' (absent modalities + random suffix lengths), 'tcga' (scattered protein pads + absent)."""
    g = torch.Generator().manual_seed(seed)
    B = int(batch_size or config["batch_size"])
    batch = {}
    for name, enc in config["encoder_configs"].items():
        L = int(enc["max_tokens"])
        kind = enc["type"]
        if kind == "EmbeddedSequenceEncoder":
            d_in = int(enc["input_size"])
            tokens = torch.randn(B, L, d_in, generator=g)
            mask = torch.zeros(B, L, dtype=torch.bool)
            if variant in ("dropout_full", "dropout_ragged"):
                p = 0.4 if p_absent is None else p_absent
                for b in range(B):
                    if torch.rand(1, generator=g).item() < p:  # utils/dataset.py:41-42 -> None -> all pad
                        mask[b] = True
                    elif variant == "dropout_ragged":
                        n = int(torch.randint(1, L + 1, (1,), generator=g).item())
                        mask[b, n:] = True
                tokens = tokens.masked_fill(mask.unsqueeze(-1), 0.0)  # fill_value 0.0, encoders.py:341
            batch[name] = {"tokens": tokens, "attention_mask": mask}
        elif kind == "TabularEncoder":
            values = torch.randn(B, L, generator=g)
            if variant != "full":
                if name == "protein":  # NaN -> -10000 anywhere (data/process_tcga.ipynb cell 43)
                    values = torch.where(torch.rand(B, L, generator=g) < 0.1, torch.full_like(values, PAD_VALUE), values)
                p = 0.25 if p_absent is None else p_absent
                for b in range(B):
                    if torch.rand(1, generator=g).item() < p:
                        values[b] = PAD_VALUE
            batch[name] = {"values": values, "attention_mask": (values == PAD_VALUE).to(torch.long)}
        elif kind in ("SequenceEncoder", "SparseTabularEncoder"):
            # SequenceCollator encoders.py:286-311: suffix padding with pad_token 0, attention_mask = (index == 0)
            V = int(enc["num_embeddings"])
            idx = torch.randint(1, V, (B, L), generator=g)
            n_live = torch.full((B,), L, dtype=torch.long)
            if variant != "full":
                p = 0.3 if p_absent is None else p_absent
                for b in range(B):
                    n_live[b] = 0 if torch.rand(1, generator=g).item() < p else int(torch.randint(1, L + 1, (1,), generator=g))
            live = torch.arange(L).unsqueeze(0) < n_live.unsqueeze(1)
            idx = idx * live
            mask = (idx == 0).to(torch.long)
            if kind == "SequenceEncoder":
                batch[name] = {"tokens": idx, "attention_mask": mask}
            else:
                data = torch.randn(B, L, generator=g) * live  # padded with 0.0 (encoders.py:308-309)
                batch[name] = {"indices": idx, "data": data, "attention_mask": mask}
        elif kind == "PatchEncoder":
            # MatrixCollator layout: {'values': f32 [B,H,W]}; absent / padded regions hold pad_token -10000
            Hh, Ww = int(enc["height"]), int(enc["width"])
            values = torch.randn(B, Hh, Ww, generator=g)
            if variant != "full":
                p = 0.3 if p_absent is None else p_absent
                for b in range(B):
                    if torch.rand(1, generator=g).item() < p:
                        values[b] = PAD_VALUE
                    else:
                        w_live = int(torch.randint(1, Ww + 1, (1,), generator=g))
                        values[b, :, w_live:] = PAD_VALUE
            batch[name] = {"values": values}
        else:
            raise NotImplementedError(kind)
    return batch


def batch_to(batch, device, non_blocking: bool = False):
    return {m: {k: v.to(device, non_blocking=non_blocking) for k, v in d.items()} for m, d in batch.items()}


def live_token_fraction(batch, n_fusion: int) -> float:
    live = total = 0
    for d in batch.values():
        if "attention_mask" not in d:
            continue
        m = d["attention_mask"].to(torch.bool)
        live += int((~m).sum())
        total += m.numel()
    B = next(iter(batch.values()))["attention_mask"].shape[0]
    return (live + B * n_fusion) / (total + B * n_fusion)
