"""Device time of the all-pairs InfoNCE kernels for the gathered batch of G ranks (GB = 8 G columns), CMU_config1 pairs."""
import sys, torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C
from mca_paper_b200.model import MCA
from mca_paper_b200.ops import P, S, call
dev = "cuda"
cfg = C.named_config(sys.argv[1] if len(sys.argv) > 1 else "CMU_config1")
model = MCA(**C.get_model_config(cfg)).to(dev)
eng = model.engine
eng.ensure_flat()
ws = eng.ws
ws["present"].fill_(1)
B, R, D = eng.B, eng.R, 512
for G in (1, 2, 4, 8):
    GB = G * B
    pooled = torch.randn(GB, R, D, device=dev) * 0.3
    dall = torch.zeros(GB, R, D, device=dev)
    dscale = torch.zeros(1, device=dev)
    s = eng.pview("loss.loss_fn.logit_scale")
    def fwd():
        call("mca_contrastive_allpairs_fwd", P(pooled), P(ws["present"]), P(eng.loss_plan), eng.plan.n_pairs, P(s), B, GB, R, D,
             eng.plan.n_mod, G - 1, 0.0, 4.6052, P(ws["losses"]), P(ws["summary"]), P(ws["w_default"]), S())
    def bwd():
        call("mca_contrastive_allpairs_bwd", P(pooled), P(ws["present"]), P(eng.loss_plan), eng.plan.n_pairs, P(s), B, GB, R, D,
             eng.plan.n_mod, G - 1, P(ws["w_default"]), P(dall), P(dscale), S())
    for f, name in ((fwd, "fwd"), (bwd, "bwd")):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        print(f"G={G} GB={GB} pairs={eng.plan.n_pairs} loss {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
