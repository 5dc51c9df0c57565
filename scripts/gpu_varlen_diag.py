"""Diagnostic: run-to-run reproducibility of the fused step and the effect of the varlen modes (off / exact / fast) on loss,
embeddings and every parameter gradient.  usage: python scripts/gpu_varlen_diag.py [p_absent]"""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA

p_absent = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
cfg = C.tiny_config("cmu", fcl=True)
enc = cfg["encoder_configs"]
enc["COVAREP"]["max_tokens"], enc["FACET"]["max_tokens"], enc["OpenFace"]["max_tokens"] = 640, 300, 260
kw = C.get_model_config(cfg)
torch.manual_seed(0)
model = MCA(**kw).to("cuda")
batch = S.batch_to(S.make_batch(cfg, seed=5, variant="dropout_ragged", p_absent=p_absent), "cuda")


def run(mode):
    model.engine.set_varlen(mode)
    for p in model.parameters():
        p.grad = None
    out = model(batch)
    out["loss"].backward()
    torch.cuda.synchronize()
    return (float(out["loss"].detach()), {k: v.detach().clone() for k, v in out.items() if isinstance(v, torch.Tensor) and v.dim() == 2},
            {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


runs = [(m, run(m)) for m in ("off", "off", "exact", "exact", "fast")]
base = runs[0][1]
for i, (m, r) in enumerate(runs[1:], 1):
    e = max(rel(r[1][k], base[1][k]) for k in base[1])
    g = sorted(((rel(r[2][k], base[2][k]), k) for k in base[2]), reverse=True)
    print(f"run {i} [{m}] vs run 0 [off]: dloss {abs(r[0] - base[0]) / abs(base[0]):.2e}  worst emb {e:.2e}  worst grads {[(f'{x:.1e}', k) for x, k in g[:3]]}  median grad {g[len(g) // 2][0]:.1e}", flush=True)
