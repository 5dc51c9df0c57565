import sys, torch
sys.path.insert(0, ".")
from mca_paper_b200.ops import P, call
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
V, B, Lq, d = 60, 4, 24, 512
bad_mine = bad_torch = 0
for trial in range(300):
    torch.manual_seed(trial)
    emb0 = torch.randn(V, d, device=dev) * 0.06
    idx = torch.randint(0, 20, (B, Lq), device=dev)
    # manual deterministic reference
    want = emb0.clone()
    u = torch.unique(idx)
    n = want[u].norm(dim=1, keepdim=True)
    want[u] = torch.where(n > 1.0, want[u] * (1.0 / (n + 1e-7)), want[u])
    ref_mod = torch.nn.Embedding(V, d, padding_idx=0, max_norm=1.0).to(dev)
    with torch.no_grad():
        ref_mod.weight.copy_(emb0)
    ref_mod(idx)
    emb = emb0.clone()
    flags = torch.zeros(V, device=dev, dtype=torch.uint8)
    bad = torch.zeros(1, device=dev, dtype=torch.int32)
    call("mca_embedding_renorm_indexed", P(emb), P(idx), B * Lq, V, d, 1.0, P(flags), P(bad), st())
    e_m = ((emb - want).norm() / want.norm()).item()
    e_t = ((ref_mod.weight.detach() - want).norm() / want.norm()).item()
    bad_mine += e_m > 1e-6
    bad_torch += e_t > 1e-6
    if e_m > 1e-6 or e_t > 1e-6:
        print(trial, "mine", e_m, "torch", e_t)
print("trials 300: mine off", bad_mine, "torch off", bad_torch)
