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
# (the peer-memory optimiser keeps AdamW moments only for the rank's own shard: compare parameters there)
chk = torch.stack([eng.flat.double().sum(), eng.flat.double().abs().sum(),
                   eng.exp_avg.double().abs().sum() if eng._p2p is None else eng.flat.double().pow(2).sum()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
eng.check_p2p()
# same weights, same batch: the peer-memory loss exchange against the NCCL all_gather / reduce_scatter form
if eng._p2p is not None:
    pooled = eng.trunk_forward(tr._dev_batch)
    l_a = eng.loss_forward(pooled)[0].clone()
    d_a = eng.loss_backward(eng.ws["w_default"]).clone()
    saved, eng._p2p = eng._p2p, None
    l_b = eng.loss_forward(pooled)[0].clone()
    d_b = eng.loss_backward(eng.ws["w_default"]).clone()
    eng._p2p = saved
    torch.cuda.synchronize()
    dl = (l_a - l_b).abs().nan_to_num().max().item()
    dd = ((d_a - d_b).norm() / d_b.norm()).item()
    print(f"rank {rank}: p2p vs nccl  max |dloss| {dl:.3e}  rel |d dpooled| {dd:.3e}  losses[:4] {l_a[:4].tolist()} / {l_b[:4].tolist()}", flush=True)
    assert dl < 1e-4 and dd < 1e-5, "peer-memory loss path disagrees with the NCCL path"
mode = "p2p (symmetric memory)" if eng._p2p is not None else "nccl all_gather/reduce_scatter"
if rank == 0:
    print(f"world {world} [{mode}]: losses rank0 {[f'{l:.6f}' for l in losses]}; replicas identical: {same}; "
          f"param checksum {allc[0][1].item():.6f}", flush=True)
    assert same, "replicas diverged"
    assert all(l == l for l in losses), "NaN loss"
# checkpoint under the sharded optimiser: save_state gathers the moment shards (collective), rank 0 writes accelerate's
# directory layout; a perturbed trainer that loads it must reproduce the uninterrupted run's next step on every rank
from mca_paper_b200 import checkpoint as K
ckpt = "/tmp/mca_dp_ckpt"
K.save_state(tr, ckpt)
dist.barrier()
s_next = float(tr.step(batch)[0])
want = eng.flat.clone()
tr.step(batch)                                   # move on, so that load_state has something to undo
assert K.load_state(tr, ckpt) == 3
assert abs(eng.lr_at(4) - 1e-4 * (3 * world) / 10.0) < 1e-10 or 3 * world >= 10   # scheduler walked `world` per step
s_again = float(tr.step(batch)[0])
torch.cuda.synchronize()
rel = ((eng.flat - want).norm() / want.norm()).item()
print(f"rank {rank}: resume loss {s_again:.6f} vs {s_next:.6f}, rel |dparams| {rel:.3e}", flush=True)
assert abs(s_again - s_next) <= 2e-3 * abs(s_next) and rel < 1e-5, "resume from checkpoint differs"
eng.check_p2p()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DP CHECK OK", flush=True)
