"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's batch producers, the step in front of the hot path
(SURVEY.md §8f rank 1): MultimodalCollator / EmbeddedSequenceCollator / SequenceCollator / MatrixCollator
(encoders.py:286-403) and the dataset-time modality dropout BatchPreDropout (utils/dataset.py:29-57).

Only tests/ may import this module; the product path (mca_paper_b200/collate.py) must never do so.  Pinned against the
live reference collators in tests/test_collate.py (build container) and against tests/golden/collate_*.pt elsewhere.
"""
from __future__ import annotations

from collections import defaultdict

import torch
from torch.nn.functional import pad


def embedded_sequence_collate(items, pad_len, embedding_size, fill_value=0.0, truncate=True, clean=True):
    """EmbeddedSequenceCollator.__call__ encoders.py:330-343: None -> empty [0, E]; truncate to pad_len; nan_to_num;
    attention_mask True on the padded suffix; tokens padded with fill_value."""
    items = [x if x is not None else torch.empty([0, embedding_size]) for x in items]
    if truncate:
        items = [x[:pad_len] for x in items]
    if clean:
        items = [x.nan_to_num() for x in items]
    mask = [pad(torch.zeros(x.shape[0]), (0, pad_len - x.shape[0]), mode="constant", value=1).to(torch.bool) for x in items]
    tokens = [pad(x, (0, 0, 0, pad_len - x.shape[-2]), mode="constant", value=fill_value) for x in items]
    return {"attention_mask": torch.stack(mask), "tokens": torch.stack(tokens)}


def sequence_collate(data: dict, pad_len, pad_token=0, data_col_name="indices", other_col="data"):
    """SequenceCollator.__call__ encoders.py:300-311: None -> empty; pad with pad_token; attention_mask =
    (padded == pad_token) as int64 — pads INSIDE the data count too (TCGA protein NaN fills).  Reference quirk: line 301
    rebuilds `data` with ONLY the data_col_name key, so the `other_col` branch (:308-310) can never fire and a 'data'
    column of the samples is silently dropped (verified against the live collator, tests/golden/collate.pt)."""
    idx = [x if x is not None else torch.empty([0]) for x in data[data_col_name]]
    out = {data_col_name: [pad(x, (0, pad_len - x.shape[-1]), mode="constant", value=pad_token) for x in idx]}
    out["attention_mask"] = [(x == pad_token).to(torch.long) for x in out[data_col_name]]
    return {k: torch.stack(v) for k, v in out.items()}


def matrix_collate(items, pad_len, pad_token=-10000, max_channels=0):
    """MatrixCollator.__call__ encoders.py:355-364."""
    items = [torch.full((max_channels, pad_len), pad_token, dtype=torch.float) if x is None else x for x in items]
    vals = [pad(x, (0, 0, 0, pad_len - x.shape[0]), mode="constant", value=pad_token) for x in items]
    if max_channels:
        vals = [x[:, :max_channels] for x in vals]
    return {"values": torch.stack(vals)}


def multimodal_collate(modality_config: dict, batch: list):
    """MultimodalCollator.__call__ encoders.py:386-403 (labels omitted: they do not enter the model)."""
    d = defaultdict(lambda: defaultdict(list))
    for b in batch:
        for k in modality_config:
            for k2, v2 in b[k].items():
                d[k][k2].append(v2)
    out = {}
    for k, cfg in modality_config.items():
        t = cfg["type"]
        if t == "embedded_sequence":
            col = cfg.get("data_col_name", "values")
            out[k] = embedded_sequence_collate(d[k][col], cfg.get("pad_len", 2048), cfg.get("embedding_size", 512),
                                               cfg.get("fill_value", 0.0), cfg.get("truncate", True), cfg.get("clean", True))
        elif t == "sequence":
            out[k] = sequence_collate(d[k], cfg.get("pad_len", 2048), cfg.get("pad_token", 0),
                                      cfg.get("data_col_name", "indices"), cfg.get("other_col", "data"))
        elif t == "matrix":
            out[k] = matrix_collate(d[k]["values"], cfg.get("pad_len", 2048), cfg.get("pad_token", -10000),
                                    cfg.get("max_channels", 0))
        else:
            raise KeyError(t)
    return out


def predrop(samples: list, modality_config: dict):
    """batch_predrop / BatchPreDropout utils/dataset.py:29-69 in "delete" mode: for every sample, for every modality with
    a dropout rate, `torch.rand(1) < dropout` sets all of the modality's fields to None (global torch RNG, this order)."""
    out = []
    for s in samples:
        s2 = {}
        for k, v in s.items():
            p = modality_config.get(k, {}).get("dropout") if k in modality_config else None
            if p and bool(torch.rand(1) < p):
                s2[k] = {kk: None for kk in v}
            else:
                s2[k] = dict(v)
        out.append(s2)
    return out
