"""Debug timeline of one attention-backward CTA (library built with MCA_NVCC_EXTRA=-DMCA_TRACE)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S, _lib
from mca_paper_b200.model import MCA

cfg = C.named_config("CMU_config1")
model = MCA(**C.get_model_config(cfg)).to("cuda")
eng = model.engine
eng.ensure_flat()
eng.build_offsets(S.batch_to(S.make_batch(cfg, seed=1, variant="full"), "cuda"))
ws = eng.ws
ws["qkv"][0].copy_((torch.randn(eng.M, 1536, device="cuda") * 0.5).bfloat16())
ws["dattn"].copy_(torch.randn(eng.M, 512, device="cuda").bfloat16())
eng.attention_fwd(ws["qkv"][0], ws["ao"][0], ws["lse"][0])
for _ in range(3):
    eng.attention_bwd(0)
torch.cuda.synchronize()
n = 4 * 16 * 16 + 8
buf = (ctypes.c_longlong * n)()
rc = _lib.lib().mca_debug_read_trace(buf, n)
a = np.array(buf[:], dtype=np.int64)
g = a[4 * 16 * 16:]
t0 = g[0]
print("global: setup_done=0 kv_full=%d q0_full=%d epi_wg0=%d epi_wg1=%d end_wg0=%d end_wg1=%d cta_end=%d" % tuple(int(x - t0) for x in g[1:8]))
ev = a[:4 * 16 * 16].reshape(4, 16, 16)[:, :, :16]
names = {0: "WG0", 1: "WG1", 2: "MMA"}
for t in range(13):
    for role in (2, 0, 1):
        row = ev[role, t]
        print(f"t={t:2d} {names[role]}: " + " ".join(f"{int(x - t0):7d}" if x else "      -" for x in row))
