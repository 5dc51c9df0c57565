"""One launch each of the uniformity / retrieval pair-tile kernels at M = 8192, D = 512 for an ncu capture."""
import sys
import torch
sys.path.insert(0, ".")
from mca_paper_b200.utils import metrics as M
g = torch.Generator().manual_seed(0)
x = torch.randn(8192, 512, generator=g).cuda()
y = x + 0.5 * torch.randn(8192, 512, generator=g).cuda()
print(float(M.lunif(x)), M.retrieval_ranks(x, y, torch.arange(8192, device="cuda")).sum().item())
