"""Device timing of the attention kernels in isolation at CMU_config1 shape (CUDA events around back-to-back launches)."""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA

dev = "cuda"
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "CMU_config1"
variant = sys.argv[2] if len(sys.argv) > 2 else "full"
cfg = C.named_config(cfg_name)
kw = C.get_model_config(cfg)
torch.manual_seed(0)
model = MCA(**kw).to(dev)
eng = model.engine
eng.ensure_flat()
eng.build_offsets(S.batch_to(S.make_batch(cfg, seed=1, variant=variant), dev))
B, N, H, M = eng.B, eng.N, eng.H, eng.M
ws = eng.ws
qkv = (torch.randn(M, 1536, device=dev) * 0.5).bfloat16()
ws["qkv"][0].copy_(qkv)
ws["dattn"].copy_(torch.randn(M, 512, device=dev).bfloat16())


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


fl = eng.plan.allowed_pairs * B * H * 4 * 64
t = timeit(lambda: eng.attention_fwd(ws["qkv"][0], ws["ao"][0], ws["lse"][0]))
print(f"attn_fwd {cfg_name}/{variant}: {t:.1f} us  ({fl / t / 1e6:.0f} TFLOP/s algorithmic, tile pairs {eng.plan.n_tile_pairs})")
t = timeit(lambda: eng.attention_bwd(0))
print(f"attn_bwd {cfg_name}/{variant}: {t:.1f} us  ({2 * fl / t / 1e6:.0f} TFLOP/s algorithmic)")
