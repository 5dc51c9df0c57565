"""Where does the end-to-end step (bench.py `e2e`) lose time against the device-resident one?  Times the fused step's loop
with the host-side pieces added one by one.  usage: python scripts/gpu_e2e_probe.py [steps]"""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg = C.named_config("CMU_config1")
torch.manual_seed(43)
model = MCA(**C.get_model_config(cfg)).to("cuda")
tr = Trainer(model, lr=1e-4, clip=2.0, schedule="cosine", warmup_steps=3000, total_steps=100000)
batch = S.make_batch(cfg, seed=1)
tr.stage(batch)
for _ in range(6):
    tr.step_staged()
tr._ensure_pipeline()
torch.cuda.synchronize()
pin = [torch.empty(4, pin_memory=True) for _ in range(2)]
evt = [torch.cuda.Event() for _ in range(2)]


def timed(name, body, pre=None):
    torch.cuda.synchronize()
    if pre:
        pre()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        body(i)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:58s} {e0.elapsed_time(e1) / steps:7.3f} ms/step", flush=True)


def handover(slot):
    cur = torch.cuda.current_stream()
    cur.wait_event(tr._ready[slot])
    for m, d in tr._slots[slot].items():
        for k, v in d.items():
            tr._dev_batch[m][k].copy_(v, non_blocking=True)
    tr._consumed[slot].record()


def full(i, d2h=True, sync=True, h2d=True, d2d=True):
    if h2d and i + 1 < steps:
        tr.prefetch((i + 1) & 1)
    if d2d:
        handover(i & 1)
    s = tr.step_staged()
    if d2h:
        pin[i & 1].copy_(s, non_blocking=True)
        evt[i & 1].record()
        if sync and i > 0:
            evt[(i - 1) & 1].synchronize()


for rep in range(2):
    timed("graph replay only", lambda i: tr.step_staged())
    timed("+ device-to-device hand-over of the batch", lambda i: full(i, d2h=False, h2d=False), pre=lambda: (tr.prefetch(0), tr.prefetch(1)))
    timed("+ H2D prefetch on the copy stream", lambda i: full(i, d2h=False), pre=lambda: tr.prefetch(0))
    timed("+ D2H of the loss summary (no host wait)", lambda i: full(i, sync=False), pre=lambda: tr.prefetch(0))
    timed("+ host waits for the previous step's loss (= bench e2e)", lambda i: full(i), pre=lambda: tr.prefetch(0))
    timed("H2D prefetch only (no hand-over)", lambda i: full(i, d2h=False, d2d=False), pre=lambda: tr.prefetch(0))
