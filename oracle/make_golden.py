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
from mca_paper_b200.model import EAO, MCA  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    "tiny_cmu_fcl_full": dict(cfg=("cmu", dict(fcl=True)), variant="full", seed=1),
    "tiny_cmu_fcl_ragged": dict(cfg=("cmu", dict(fcl=True)), variant="dropout_ragged", seed=1),
    "tiny_cmu_mma_absent": dict(cfg=("cmu", dict(zorro=True, fcl=False)), variant="dropout_full", seed=2),
    "tiny_tcga_all_losses": dict(cfg=("tcga", dict(fcl=True, bimodal=True, non_fusion_fcl=True)), variant="tcga", seed=1),
    # SequenceEncoder + SparseTabularEncoder + PatchEncoder + EmbeddedSequenceEncoder, ragged / absent modalities
    "tiny_mixed_encoders": dict(cfg=("mixed", dict(fcl=True)), variant="dropout_ragged", seed=3),
    # EAO baseline (model.py:481-596) as the shipped *_EAO configs run it: 4 single + 6 pair passes, mean pooling,
    # 6 modality pairs + 20 fcl_<modality>|<pair> losses; ragged lengths and absent modalities
    "tiny_cmu_eao": dict(cfg=("cmu", dict(fcl=True, bimodal=True, non_fusion_fcl=True, eao=True)), variant="dropout_ragged", seed=1),
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
    mine = (EAO if kw.get("eao") else MCA)(**kw)
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


COLLATE_CONFIG = {
    "speech": {"type": "embedded_sequence", "pad_len": 12, "data_col_name": "data", "pad_token": -10000, "embedding_size": 5,
               "dropout": 0.3},
    "expr": {"type": "sequence", "pad_len": 9, "data_col_name": "values", "pad_token": -10000, "dropout": 0.3},
    "text": {"type": "sequence", "pad_len": 7, "data_col_name": "indices", "pad_token": 0},
    "spec": {"type": "matrix", "pad_len": 6, "pad_token": -10000, "max_channels": 4},
}


def collate_samples(seed: int, batch_size: int = 8):
    """Ragged synthetic samples in the reference's per-sample layout: absent modalities (None), over-long sequences
    (truncation), NaN / inf entries (clean), pad values inside the data (TCGA protein convention)."""
    g = torch.Generator().manual_seed(seed)
    samples = []
    for b in range(batch_size):
        n = int(torch.randint(0, 17, (1,), generator=g))
        sp = torch.randn(n, 5, generator=g)
        if n > 2:
            sp[1, 2], sp[2, 0], sp[0, 4] = float("nan"), float("inf"), float("-inf")
        ex = torch.randn(int(torch.randint(1, 10, (1,), generator=g)), generator=g)
        if ex.numel() > 3:
            ex[2] = -10000.0
        tx = torch.randint(1, 50, (int(torch.randint(1, 8, (1,), generator=g)),), generator=g)
        mt = torch.randn(int(torch.randint(1, 7, (1,), generator=g)), 4, generator=g)
        samples.append({
            "speech": {"data": None if b == 3 else sp},
            "expr": {"values": None if b == 5 else ex},
            "text": {"indices": tx, "data": torch.randn(tx.numel(), generator=g)},
            "spec": {"values": None if b == 6 else mt},
        })
    return samples


def make_collate_golden():
    """Outputs of the LIVE reference collators (encoders.py:374-403) and of its dataset-time modality dropout
    (utils/dataset.py:29-57) on collate_samples(seed)."""
    ref_shim.load_reference()
    import encoders as ref_enc  # the reference's module (sys.path set up by ref_shim)
    from utils import dataset as ref_ds
    gold = {}
    for seed in (11, 12):
        samples = collate_samples(seed)
        out = ref_enc.MultimodalCollator(COLLATE_CONFIG)([{k: dict(v) for k, v in s.items()} for s in samples])
        gold[f"collate_{seed}"] = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in out.items()}
        # dropout decisions of BatchPreDropout, one instance per modality as batch_predrop builds them
        torch.manual_seed(100 + seed)
        drops = {k: ref_ds.BatchPreDropout(kvs={"attention_mask": c["pad_token"], "data": 0.0}, dropout=c["dropout"])
                 for k, c in COLLATE_CONFIG.items() if c.get("dropout")}
        dropped = []
        for s in samples:
            row = {}
            for k, v in s.items():
                if k in drops:
                    res = drops[k]({kk: vv for kk, vv in v.items()})
                    row[k] = all(x is None for x in res.values())
            dropped.append(row)
        gold[f"dropped_{seed}"] = dropped
    torch.save(gold, os.path.join(GOLDEN_DIR, "collate.pt"))
    print("collate golden:", {k: (list(v.keys()) if isinstance(v, dict) else len(v)) for k, v in gold.items()})


def metrics_inputs(seed: int = 11):
    """Embeddings shaped like the eval loop's (pooled [M, 512] rows): two correlated sets (positive pairs), a mask with
    holes, and a nearly collapsed set (uniformity close to 0, all cancellation)."""
    g = torch.Generator().manual_seed(seed)
    M, D = 150, 512
    x = torch.randn(M, D, generator=g)
    y = x + 0.5 * torch.randn(M, D, generator=g)
    mask = torch.rand(M, generator=g) > 0.3
    collapsed = torch.randn(1, D, generator=g) + 1e-2 * torch.randn(70, D, generator=g)
    small = torch.randn(5, 48, generator=g)
    yr = 0.07 * x + torch.randn(M, D, generator=g)   # weakly aligned targets: retrieval ranks spread over 0..M
    return {"x": x, "y": y, "yr": yr, "mask": mask, "collapsed": collapsed, "small": small}


def make_metrics_golden():
    """Outputs of the LIVE reference functions utils/metrics.py:20-33,73-99 on metrics_inputs()."""
    R = ref_shim.load_reference_metrics()
    inp = metrics_inputs()
    x, y, mask = inp["x"], inp["y"], inp["mask"]
    out = {}
    for norm in (True, False):
        for alpha in (2, 1, 3.5):
            out[f"lalign_a{alpha}_n{int(norm)}"] = R.lalign(x, y, alpha, norm)
        for t in (2, 0.5):
            out[f"lunif_t{t}_n{int(norm)}"] = R.lunif(x, t, norm)
            out[f"lunif_collapsed_t{t}_n{int(norm)}"] = R.lunif(inp["collapsed"], t, norm)
            out[f"lunif_small_t{t}_n{int(norm)}"] = R.lunif(inp["small"], t, norm)
    out["wang"] = R.wang_loss(x, y)
    # accumulators: three updates, compute over the concatenation
    al, un = R.Alignment(), R.Uniformity()
    for a, b in zip(x.chunk(3), y.chunk(3)):
        al.update(a, b)
        un.update(a)
    out["Alignment"], out["Uniformity"] = al.compute(), un.compute()
    out["Alignment_norm"], out["Uniformity_norm"] = al.compute(norm=True), un.compute(norm=True)
    for name, tg in (("easy", y), ("hard", inp["yr"])):
        med, r1, r5, r10 = R.get_rank_metrics(x, mask, tg, device="cpu")
        out[f"rank_metrics_{name}"] = torch.stack([med.double(), r1.double(), r5.double(), r10.double()])
        c = torch.stack([R.compute_cosines(x[i], tg) for i in range(x.shape[0]) if mask[i]])
        out[f"ranks_{name}"] = R.get_rank(c, torch.nonzero(mask).reshape(-1))
    gold = {"inputs": inp, "outputs": {k: v.detach().clone() for k, v in out.items()}}
    torch.save(gold, os.path.join(GOLDEN_DIR, "metrics.pt"))
    print("metrics golden:", {k: (v.tolist() if v.numel() < 5 else tuple(v.shape)) for k, v in gold["outputs"].items()})


def main():
    if not ref_shim.available():
        raise SystemExit("the live reference is not present; golden fixtures can only be regenerated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in ("", "cases") or only.startswith("case:"):
        for name, spec in CASES.items():
            if only.startswith("case:") and name != only[5:]:
                continue
            make_case(name, spec)
    if only in ("", "static"):
        make_static()
    if only in ("", "collate"):
        make_collate_golden()
    if only in ("", "metrics"):
        make_metrics_golden()


if __name__ == "__main__":
    main()
