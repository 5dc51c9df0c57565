"""Per-tensor gradient error of the fused path against the oracle (fp32 on the same GPU) at full size, well-conditioned
variant (what tests/test_gpu_fullsize.py asserts on).  usage: python scripts/gpu_grad_fullsize.py [config] [variant]"""
import sys
sys.path.insert(0, ".")
import torch
from mca_paper_b200 import synthetic as S
from tests import helpers as H
from tests import test_gpu_fullsize as T

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "CMU_config1"
variant = sys.argv[2] if len(sys.argv) > 2 else "full"
scales = dict(to_out=float(sys.argv[3]), return_tokens=float(sys.argv[4]), logit_scale=float(sys.argv[5])) if len(sys.argv) > 5 else None
cfg, kw, model, sd, names = T._build(cfg_name, well_conditioned=True, scales=scales)
batch = S.make_batch(cfg, seed=1, variant=variant)
ref, ref_grads = T._oracle_on_gpu(kw, sd, batch, names)
model = model.to("cuda")
out = model(S.batch_to(batch, "cuda"))
out["loss"].backward()
torch.cuda.synchronize()
print("loss", out["loss"].item(), float(ref["loss"]))
errs = sorted(((H.rel_err(p.grad, ref_grads[k]), k, float(ref_grads[k].norm())) for k, p in model.named_parameters()
               if float(ref_grads[k].abs().max()) > 0), reverse=True)
for e in errs[:12]:
    print("%.4f  %-50s |g|=%.3e" % e)
print("median %.4f" % errs[len(errs) // 2][0])
