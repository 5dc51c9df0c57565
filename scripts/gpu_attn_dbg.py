"""Diagnostic for the attention backward: which precision of (lse, delta) does the kernel effectively apply?"""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.ops import P, call
from oracle import mca_oracle as O

dev = "cuda"
cfg = C.tiny_config("cmu", fcl=True)
torch.manual_seed(0)
model = MCA(**C.get_model_config(cfg)).to(dev)
eng = model.engine
eng.ensure_flat()
eng.build_offsets(S.batch_to(S.make_batch(cfg, seed=1, variant="full"), dev))
B, N, H, M = eng.B, eng.N, eng.H, eng.M
g = torch.Generator(device=dev).manual_seed(3)
qkv = torch.randn(M, 1536, device=dev, generator=g).bfloat16()
out = torch.zeros(M, 512, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, N, device=dev)
eng.attention_fwd(qkv, out, lse)
do = torch.randn(M, 512, device=dev, generator=g).bfloat16()
ws = eng.ws
ws["dattn"].copy_(do)
call("mca_attn_bwd", P(qkv), P(out), P(ws["dattn"]), P(lse), P(eng.k_tiles_q), eng.n_kt, P(eng.qt_list), P(eng.k_tiles),
     int(eng.q_tiles.shape[0]), P(eng.rowbits), P(eng.keygrp), P(eng.tile_grp), P(ws["padding"]), P(ws["kt_class"]),
     None, P(ws["delta"]), P(ws["ucorr"]), P(ws["dq_acc"]), P(ws["dqkv"]), B, N, H, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
d = ws["dqkv"].float()
x = qkv.float().view(B, N, 3, H, 64)
q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
dO = do.float().view(B, N, H, 64).permute(0, 2, 1, 3)
Of = out.float().view(B, N, H, 64).permute(0, 2, 1, 3)
sim = q @ k.transpose(-1, -2)
masked = model.attn_mask[None, None] | eng.ws["padding"].bool()[:, None, None, :]
delta = (dO * Of).sum(-1)


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def bf(x, terms):
    acc = torch.zeros_like(x)
    r = x.clone()
    for _ in range(terms):
        h = r.bfloat16().float()
        acc += h
        r = r - h
    return acc


for name, lt, dt in [("exact", 9, 9), ("lse hi only", 1, 9), ("lse hi+mid", 2, 9), ("delta hi only", 9, 1), ("both hi only", 1, 1)]:
    l = lse if lt == 9 else bf(lse, lt)
    dl = delta if dt == 9 else bf(delta, dt)
    Pm = torch.exp(sim - l[..., None]).masked_fill(masked, 0.0)
    dP = dO @ v.transpose(-1, -2)
    dS = Pm * (dP - dl[..., None])
    dQ = (dS @ k).permute(0, 2, 1, 3).reshape(M, 512)
    dK = (dS.transpose(-1, -2) @ q).permute(0, 2, 1, 3).reshape(M, 512)
    dV = (Pm.transpose(-1, -2) @ dO).permute(0, 2, 1, 3).reshape(M, 512)
    print(f"{name:14s} dQ {rel(d[:, :512], dQ):.4f}  dK {rel(d[:, 512:1024], dK):.4f}  dV {rel(d[:, 1024:], dV):.4f}")
print("lse range", lse.min().item(), lse.max().item(), " delta range", delta.min().item(), delta.max().item())
