"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import hashlib
import os

import torch

from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import EAO, MCA

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror oracle/make_golden.py CASES
CASES = {
    "tiny_cmu_fcl_full": dict(cfg=("cmu", dict(fcl=True)), variant="full", seed=1),
    "tiny_cmu_fcl_ragged": dict(cfg=("cmu", dict(fcl=True)), variant="dropout_ragged", seed=1),
    "tiny_cmu_mma_absent": dict(cfg=("cmu", dict(zorro=True, fcl=False)), variant="dropout_full", seed=2),
    "tiny_tcga_all_losses": dict(cfg=("tcga", dict(fcl=True, bimodal=True, non_fusion_fcl=True)), variant="tcga", seed=1),
    # SequenceEncoder + SparseTabularEncoder + PatchEncoder + EmbeddedSequenceEncoder, ragged / absent modalities
    "tiny_mixed_encoders": dict(cfg=("mixed", dict(fcl=True)), variant="dropout_ragged", seed=3),
    # EAO baseline (model.py:481-596): 4 single + 6 pair passes, mean pooling, 26 losses
    "tiny_cmu_eao": dict(cfg=("cmu", dict(fcl=True, bimodal=True, non_fusion_fcl=True, eao=True)), variant="dropout_ragged", seed=1),
}


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def weights_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def batch_checksum(batch) -> str:
    h = hashlib.sha256()
    for m in batch:
        for k in sorted(batch[m].keys()):
            h.update(batch[m][k].contiguous().numpy().tobytes())
    return h.hexdigest()


def build_case(name):
    """(config, model kwargs, CPU model with the seeded init the fixtures were made from, CPU state_dict, batch)."""
    spec = CASES[name]
    kind, kwargs = spec["cfg"]
    cfg = C.tiny_config(kind, **kwargs)
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = (EAO if kw.get("eao") else MCA)(**kw)   # train_accel_gpu.py:47-52
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = S.make_batch(cfg, seed=spec["seed"], variant=spec["variant"])
    return cfg, kw, model, sd, batch


def oracle_forward(kw):
    """The oracle entry point of the model family the kwargs select."""
    from oracle import mca_oracle as O
    return O.eao_forward if kw.get("eao") else O.mca_forward


def key_to_str(k):
    return k if isinstance(k, str) else "combo:" + ",".join(str(i) for i in sorted(k))


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
