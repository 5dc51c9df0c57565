"""Device-side collate + modality dropout: the producer of the hot path's batch.

Mirrors the reference's `MultimodalCollator(modality_config)(list_of_samples)` (encoders.py:374-403) and the dataset-time
`BatchPreDropout` (utils/dataset.py:29-69), but builds the batch ON THE GPU: the host only concatenates the live rows of the
present samples into one pinned staging buffer per modality (varlen: no padding, nothing for absent modalities — the
north-star's "tokens of absent or padded modalities are never read"), one H2D copy per modality moves them, and
mca_collate_rows / mca_collate_values_* expand them into the collators' dense layouts and masks.  At 1500 samples/s per
GPU the reference's 8 CPU collator workers + a dense 14.8 MB/batch H2D (train_accel_gpu.py:70,111) are the next bottleneck.

    collate = DeviceCollator(cfg["modality_config"], batch_size=8, device="cuda")
    batch = collate(samples)                 # samples: list of {modality: {data_col_name: Tensor | None, ...}}
    out = model(batch)

Dropout: `dropout` entries of modality_config are applied per (sample, modality) with `torch.rand(1) < p` in the
reference's order (sample-major, modality order of the sample dict); dropped modalities are never staged or copied.
Semantics differ from the reference in ONE documented way: `batch_predrop` (utils/dataset.py:59-69) draws the drops once,
at dataset-construction time, so a sample keeps the same dropped modalities in every epoch; `apply_dropout=True` here
redraws them at every call (fresh drops per epoch).  To reproduce the reference exactly, call `predrop(samples)` ONCE
over the whole dataset under the run's seed (same RNG calls, same order) and collate the result with
`apply_dropout=False`.  The device buffers handed out are reused every second call of the same collator.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .ops import P, S, call


class DeviceCollator:
    def __init__(self, modality_config: Dict[str, dict], batch_size: int, device="cuda", apply_dropout: bool = False):
        self.cfg = {k: dict(v) for k, v in modality_config.items()}
        self.B = int(batch_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceCollator builds batches with CUDA kernels: it needs a CUDA device (no CPU fallback)")
        self.apply_dropout = apply_dropout
        self._buf = {}
        self.h2d_bytes = 0

    # ------------------------------------------------------------------------------------------------ dropout
    def predrop(self, samples: List[dict]) -> List[dict]:
        """BatchPreDropout in "delete" mode (utils/dataset.py:41-52): same RNG calls in the same order."""
        out = []
        for s in samples:
            s2 = {}
            for k, v in s.items():
                p = self.cfg[k].get("dropout") if k in self.cfg else None
                s2[k] = {kk: None for kk in v} if (p and bool(torch.rand(1) < p)) else v
            out.append(s2)
        return out

    # ------------------------------------------------------------------------------------------------ staging
    def _staging(self, key, shape, dtype):
        """(pinned, device, event) for `key`.  Two pinned/device pairs per key are used alternately and each carries the
        event recorded after its last H2D copy: the host only rewrites a pinned buffer whose copy has drained (the
        collator never synchronises the device, so it can run several batches ahead of it)."""
        entry = self._buf.get(key)
        if entry is None or entry["bufs"][0][0].shape != torch.Size(shape) or entry["bufs"][0][0].dtype != dtype:
            entry = {"bufs": [(torch.empty(shape, dtype=dtype, pin_memory=True),
                               torch.empty(shape, dtype=dtype, device=self.device)) for _ in range(2)],
                     "events": [None, None], "next": 0}
            self._buf[key] = entry
        i = entry["next"]
        entry["next"] ^= 1
        if entry["events"][i] is not None:
            entry["events"][i].synchronize()
        else:
            entry["events"][i] = torch.cuda.Event()
        pin, dev = entry["bufs"][i]
        return pin, dev, entry["events"][i]

    def _stage_varlen(self, key, items: List[Optional[torch.Tensor]], max_rows: int, width: int, dtype):
        """Concatenate the items' rows (each truncated to max_rows) into pinned memory; returns (device rows, device
        offsets [B+1])."""
        pin, dev, ev_rows = self._staging(key + ".rows", (self.B * max_rows, width) if width else (self.B * max_rows,), dtype)
        opin, odev, ev_off = self._staging(key + ".off", (self.B + 1,), torch.int32)
        pos = 0
        opin[0] = 0
        for b, x in enumerate(items):
            n = 0 if x is None else min(int(x.shape[0]), max_rows)
            if n:
                pin[pos:pos + n].copy_(x[:n].to(dtype).reshape(pin[pos:pos + n].shape))
            pos += n
            opin[b + 1] = pos
        if pos:
            dev[:pos].copy_(pin[:pos], non_blocking=True)
            self.h2d_bytes += pos * max(width, 1) * pin.element_size()
        odev.copy_(opin, non_blocking=True)
        ev_rows.record()
        ev_off.record()
        self.h2d_bytes += opin.numel() * 4
        return dev, odev

    def _out(self, key, shape, dtype):
        t = self._buf.get(key)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._buf[key] = t
        return t

    # ------------------------------------------------------------------------------------------------ collate
    def __call__(self, samples: List[dict]) -> Dict[str, Dict[str, torch.Tensor]]:
        if len(samples) != self.B:
            raise AssertionError(f"batch of {len(samples)} samples, collator built for {self.B}")
        assert self.cfg.keys() <= samples[0].keys(), f"{self.cfg.keys()} - {samples[0].keys()}"  # encoders.py:387
        if self.apply_dropout:
            samples = self.predrop(samples)
        self.h2d_bytes = 0
        B, out = self.B, {}
        for name, c in self.cfg.items():
            kind, L = c["type"], int(c.get("pad_len", 2048))
            if kind == "embedded_sequence":
                col, E = c.get("data_col_name", "values"), int(c.get("embedding_size", 512))
                rows, off = self._stage_varlen(name, [s[name][col] for s in samples], L if c.get("truncate", True) else 1 << 30,
                                               E, torch.float32)
                tokens = self._out(name + ".tokens", (B, L, E), torch.float32)
                mask = self._out(name + ".mask", (B, L), torch.bool)
                call("mca_collate_rows", P(rows), P(off), B, L, E, float(c.get("fill_value", 0.0)),
                     int(bool(c.get("clean", True))), P(tokens), P(mask), S())
                out[name] = {"attention_mask": mask, "tokens": tokens}
            elif kind == "sequence":
                col, other = c.get("data_col_name", "indices"), c.get("other_col", "data")
                items = [s[name][col] for s in samples]
                first = next((x for x in items if x is not None), None)
                is_int = first is not None and not first.dtype.is_floating_point
                for x in items:
                    if x is not None and x.shape[-1] > L:
                        raise RuntimeError(f"sequence of {x.shape[-1]} > pad_len {L} (the reference collator cannot truncate)")
                dt = torch.int64 if is_int else torch.float32
                vals, off = self._stage_varlen(name, items, L, 0, dt)
                padded = self._out(name + "." + col, (B, L), dt)
                mask = self._out(name + ".mask", (B, L), torch.int64)
                if is_int:
                    call("mca_collate_values_i64", P(vals), P(off), B, L, int(c.get("pad_token", 0)), P(padded), P(mask), S())
                else:
                    call("mca_collate_values_f32", P(vals), P(off), B, L, float(c.get("pad_token", 0)), P(padded), P(mask), S())
                # (the reference collator drops the samples' `other_col` column: encoders.py:301 rebuilds its input dict
                #  with the data_col_name key only, so :308-310 never run — reproduced; `keep_other_col=True` restores it)
                out[name] = {col: padded, "attention_mask": mask}
                if c.get("keep_other_col") and other in samples[0][name]:
                    ovals, ooff = self._stage_varlen(name + ".other", [s[name][other] for s in samples], L, 0, torch.float32)
                    opad = self._out(name + "." + other, (B, L), torch.float32)
                    call("mca_collate_values_f32", P(ovals), P(ooff), B, L, 0.0, P(opad), None, S())
                    out[name][other] = opad
            elif kind == "matrix":
                C_ = int(c.get("max_channels", 0))
                items = [s[name]["values"] for s in samples]
                width = C_ or int(next(x for x in items if x is not None).shape[1])
                rows, off = self._stage_varlen(name, [None if x is None else x[:, :width] for x in items], L, width, torch.float32)
                vals = self._out(name + ".values", (B, L, width), torch.float32)
                call("mca_collate_rows", P(rows), P(off), B, L, width, float(c.get("pad_token", -10000)), 0, P(vals), None, S())
                out[name] = {"values": vals}
            else:
                raise KeyError(kind)
        return out
