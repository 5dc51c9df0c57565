"""Isolated check of the tcgen05 attention fwd/bwd kernels against torch fp32 math on identical bf16 inputs."""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.ops import P, call

dev = "cuda"
def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()

def run(cfg, variant, tag, scale=1.0):
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = MCA(**kw).to(dev)
    eng = model.engine
    eng.ensure_flat()
    batch = S.batch_to(S.make_batch(cfg, seed=1, variant=variant), dev)
    eng.build_offsets(batch)
    B, N, H, M = eng.B, eng.N, eng.H, eng.M
    g = torch.Generator(device=dev).manual_seed(3)
    qkv = (torch.randn(M, 1536, device=dev, generator=g) * scale).bfloat16()
    out = torch.zeros(M, 512, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device=dev)
    eng.attention_fwd(qkv, out, lse)
    torch.cuda.synchronize()
    # reference
    x = qkv.float().view(B, N, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, 0].permute(0, 2, 1, 3), x[:, :, 1].permute(0, 2, 1, 3), x[:, :, 2].permute(0, 2, 1, 3)
    sim = q @ k.transpose(-1, -2)
    mv = -torch.finfo(torch.float32).max
    sim = sim.masked_fill(model.attn_mask, mv).masked_fill(eng.ws["padding"].bool()[:, None, None, :], mv)
    p = sim.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(M, 512)
    print(f"== {tag}: fwd out rel {rel(out, o):.3e}")
    full = (sim.max(-1).values == mv)
    lse_ref = torch.logsumexp(sim, -1)
    ok = ~full
    print(f"   lse rel (non-masked rows) {rel(lse[ok], lse_ref[ok]):.3e}; fully-masked rows {int(full.sum())}, flagged inf {int(torch.isinf(lse).sum())}")
    do = torch.randn(M, 512, device=dev, generator=g).bfloat16()
    o.backward(do.float())
    gref = x.grad.view(M, 1536)
    ws = eng.ws
    ws["dattn"].copy_(do)
    call("mca_attn_bwd", P(qkv), P(out), P(ws["dattn"]), P(lse), P(eng.k_tiles_q), eng.n_kt, P(eng.qt_list), P(eng.k_tiles),
         int(eng.q_tiles.shape[0]), P(eng.rowbits), P(eng.keygrp), P(eng.tile_grp), P(ws["padding"]), P(ws["kt_class"]), None, P(ws["delta"]),
         P(ws["ucorr"]), P(ws["dq_acc"]), P(ws["dqkv"]), B, N, H, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d = ws["dqkv"].float()
    print(f"   bwd dQ rel {rel(d[:, :512], gref[:, :512]):.3e}  dK rel {rel(d[:, 512:1024], gref[:, 512:1024]):.3e}  dV rel {rel(d[:, 1024:], gref[:, 1024:]):.3e}")
    # per-modality breakdown of dQ error
    off = 0
    for n_, L in zip(eng.plan.names + ["fusion"], eng.plan.lengths + [eng.plan.F]):
        idx = torch.arange(off, off + L, device=dev)
        rows = (torch.arange(B, device=dev)[:, None] * N + idx[None]).flatten()
        print(f"     {n_:14s} dQ {rel(d[rows, :512], gref[rows, :512]):.3e} dK {rel(d[rows, 512:1024], gref[rows, 512:1024]):.3e} dV {rel(d[rows, 1024:], gref[rows, 1024:]):.3e}")
        off += L

run(C.tiny_config("cmu", fcl=True), "full", "tiny full", 1.0)
run(C.tiny_config("cmu", fcl=True), "dropout_ragged", "tiny ragged", 1.0)
run(C.tiny_config("cmu", zorro=True, fcl=False), "dropout_full", "tiny mma absent", 0.5)
run(C.named_config("CMU_config1"), "full", "CMU full size", 0.5)
print("DONE")
