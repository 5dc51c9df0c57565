"""Run-to-run reproducibility of the backward: one forward, then the loss + trunk backward twice from the same saved state;
every workspace buffer of the (single-layer) backward is compared between the two runs in dependency order.  Bitwise
equality is expected everywhere except behind fp32 atomics / TMA reduce-adds (dQ, weight gradients, column sums), where
differences must stay at the fp32 rounding level (~1e-7).  usage: python scripts/gpu_determinism.py [layers] [variant]"""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.trainer import Trainer

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 1
variant = sys.argv[2] if len(sys.argv) > 2 else "dropout_ragged"
cfg = C.named_config("CMU_config1")
cfg["layers"] = layers
kw = C.get_model_config(cfg)
torch.manual_seed(0)
model = MCA(**kw).to("cuda")
tr = Trainer(model, use_graphs=False)
eng = tr.eng
batch = S.make_batch(cfg, seed=5, variant=variant, p_absent=0.0 if variant != "full" else None)
tr.stage(batch)
tr._seg_forward()
torch.cuda.synchronize()
NAMES = ["dpo", "dkvp", "dqp", "dxf", "dx_a", "dx_b", "d16", "du", "dy16", "dattn", "delta", "ucorr", "dq_acc", "dqkv"]


def once():
    tr._seg_loss()
    tr._seg_backward()
    torch.cuda.synchronize()
    snap = {k: eng.ws[k].detach().clone() for k in NAMES if k in eng.ws}
    snap["dpooled"] = tr._dpooled.detach().clone()
    snap["garena"] = eng.garena.detach().clone()
    snap["flat_grad"] = eng.flat_grad.detach().clone()
    for name, e in eng.ws["enc"].items():
        for k in ("dz32", "dz16", "dy"):
            if k in e:
                snap[f"enc.{name}.{k}"] = e[k].detach().clone()
    return snap


a, b, c = once(), once(), once()
for k in a:
    x, y, z = a[k].double(), b[k].double(), c[k].double()
    n = x.norm().clamp_min(1e-30)
    print(f"{k:28s} bitwise {bool(torch.equal(a[k], b[k]))!s:5s} {bool(torch.equal(b[k], c[k]))!s:5s}  rel {float((x - y).norm() / n):.2e} {float((y - z).norm() / n):.2e}"
          f"  max|d| {float((x - y).abs().max()):.2e}  finite {bool(torch.isfinite(x).all())}", flush=True)
