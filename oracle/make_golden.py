"""TEST INFRASTRUCTURE ONLY — freezes outputs of the LIVE reference (/root/reference, build container only) into
tests/golden/ so that the oracle and the CUDA path can be checked where the reference is absent (the GPU box).

    python -m oracle.make_golden            # rewrites tests/golden/*.pt

For every case: weights come from mca_paper_b200.model.MCA's own seeded CPU init (the reference model loads that
state_dict with strict=True, which also pins the state_dict schema), the batch from mca_paper_b200.synthetic, the
reference runs in fp32 on the CPU with its debug torch.save neutralised (model.py:94) and the torchmultimodal loss
replaced by the restatement in oracle/mca_oracle.py (declared in oracle/ref_shim.py).  Stored: every returned
embedding, every named loss, parameter-gradient norms and a handful of full gradient tensors, a checksum of the
weights and inputs, and for the three full-size configs the state_dict schema and SHA-256 of the static buffers.
"""
from __future__ import annotations

import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mca_paper_b200 import config as C, synthetic as S  # noqa: E402
from mca_paper_b200.model import MCA  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    "tiny_cmu_fcl_full": dict(cfg=("cmu", dict(fcl=True)), variant="full", seed=1),
    "tiny_cmu_fcl_ragged": dict(cfg=("cmu", dict(fcl=True)), variant="dropout_ragged", seed=1),
    "tiny_cmu_mma_absent": dict(cfg=("cmu", dict(zorro=True, fcl=False)), variant="dropout_full", seed=2),
    "tiny_tcga_all_losses": dict(cfg=("tcga", dict(fcl=True, bimodal=True, non_fusion_fcl=True)), variant="tcga", seed=1),
    # SequenceEncoder + SparseTabularEncoder + PatchEncoder + EmbeddedSequenceEncoder, ragged / absent modalities
    "tiny_mixed_encoders": dict(cfg=("mixed", dict(fcl=True)), variant="dropout_ragged", seed=3),
}
FULL_GRADS = ["return_tokens", "fusion_tokens", "norm.gamma", "layers.0.norm.gamma", "loss.loss_fn.logit_scale",
              "layers.1.norm.gamma"]


def tensor_checksum(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


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


def key_to_str(k):
    return k if isinstance(k, str) else "combo:" + ",".join(str(i) for i in sorted(k))


def make_case(name, spec):
    kind, kwargs = spec["cfg"]
    cfg = C.tiny_config(kind, **kwargs)
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    mine = MCA(**kw)
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    ref = ref_shim.build_reference_model(kw, state_dict=sd)  # strict load: schema parity
    batch = S.make_batch(cfg, seed=spec["seed"], variant=spec["variant"])
    out = ref_shim.reference_forward(ref, batch)
    out["loss"].backward()
    gold = {
        "case": name, "weights_sha256": weights_checksum(sd), "batch_sha256": batch_checksum(batch),
        "embeddings": {key_to_str(k): v.detach().clone() for k, v in out.items()
                       if k not in ("losses", "modality_sample_mask") and not (isinstance(k, str) and "loss" in k)},
        "output_keys": [key_to_str(k) for k in out.keys()],
        "losses": {k: v.detach().clone() for k, v in out["losses"].items()},
        "loss": out["loss"].detach().clone(),
        "fcl_loss": out["fcl_loss"].detach().clone() if "fcl_loss" in out else None,
        "no-fcl_loss": out["no-fcl_loss"].detach().clone() if "no-fcl_loss" in out else None,
        "modality_sample_mask": {k: v.clone() for k, v in out["modality_sample_mask"].items()},
        "grad_norms": {k: (p.grad.norm().item() if p.grad is not None else 0.0) for k, p in ref.named_parameters()},
        "grads": {k: p.grad.detach().clone() for k, p in ref.named_parameters() if k in FULL_GRADS and p.grad is not None},
        "torch_version": torch.__version__,
    }
    torch.save(gold, os.path.join(GOLDEN_DIR, name + ".pt"))
    print(f"{name}: loss {gold['loss'].item():.6f}, {len(gold['losses'])} losses, {len(gold['embeddings'])} embeddings")


def make_static():
    """Schema + buffer hashes of the three full-size configs (constructed by the live reference)."""
    static = {}
    for name in ("CMU_config1", "CMU_config1_z", "TCGA_config1"):
        kw = C.get_model_config(C.named_config(name))
        torch.manual_seed(0)
        ref = ref_shim.build_reference_model(kw)
        sd = ref.state_dict()
        static[name] = {
            "state_dict_schema": {k: tuple(v.shape) for k, v in sd.items()},
            "param_names": [k for k, _ in ref.named_parameters()],
            "n_params": sum(p.numel() for p in ref.parameters()),
            "token_types_sha256": tensor_checksum(ref.token_types),
            "attn_mask_sha256": tensor_checksum(ref.attn_mask),
            "pool_mask_sha256": tensor_checksum(ref.pool_mask),
            "attn_allowed_pairs": int((~ref.attn_mask).sum()),
            "return_token_types": list(ref.return_token_types),
            "fusion_combos": [sorted(c) for c in ref.fusion_combos],
            "loss_names": None,
        }
        # loss names need a forward: run the tiny sibling config (same flags) through the reference
    for name, (kind, kwargs) in {"CMU_config1": ("cmu", dict(fcl=True)), "CMU_config1_z": ("cmu", dict(zorro=True, fcl=False)),
                                 "TCGA_config1": ("tcga", dict(fcl=True, bimodal=True, non_fusion_fcl=True))}.items():
        cfg = C.tiny_config(kind, **kwargs)
        kw = C.get_model_config(cfg)
        torch.manual_seed(0)
        ref = ref_shim.build_reference_model(kw)
        out = ref_shim.reference_forward(ref, S.make_batch(cfg, seed=1, variant="full"))
        static[name]["loss_names"] = list(out["losses"].keys())
    torch.save(static, os.path.join(GOLDEN_DIR, "static_tables.pt"))
    for k, v in static.items():
        print(k, v["n_params"], v["attn_allowed_pairs"], len(v["loss_names"]))


def main():
    if not ref_shim.available():
        raise SystemExit("the live reference is not present; golden fixtures can only be regenerated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, spec in CASES.items():
        make_case(name, spec)
    make_static()


if __name__ == "__main__":
    main()
