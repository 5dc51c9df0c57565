"""Standalone GPU check of the tcgen05 GEMM (all operand layouts / epilogues) against torch fp32."""
import ctypes, sys, time
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import _lib as L

lib = L.lib()
dev = "cuda"
torch.manual_seed(0)

def gemm(A, a_mn, B, b_mn, M, N, K, mode, out0, out1=None, aux0=None, bias=None, alpha=1.0, splits=1, sync=True):
    rc = lib.mca_gemm_bf16(L.ptr(A), a_mn, ctypes.c_longlong(A.stride(0)), L.ptr(B), b_mn, ctypes.c_longlong(B.stride(0)),
                           M, N, K, splits, mode, L.ptr(out0), ctypes.c_longlong(out0.stride(-2)),
                           L.ptr(out1), ctypes.c_longlong(out1.stride(0) if out1 is not None else 0),
                           L.ptr(aux0), ctypes.c_longlong(aux0.stride(0) if aux0 is not None else 0),
                           L.ptr(bias), ctypes.c_float(alpha), L.stream_ptr())
    L.check(rc, "gemm")
    if sync:
        torch.cuda.synchronize()

def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()

ok = True
def report(name, err, tol=1e-2):
    global ok
    good = err < tol
    ok &= good
    print(f"{name:40s} rel_err={err:.3e} {'OK' if good else 'FAIL'}", flush=True)

for (M, N, K) in [(256, 128, 64), (1000, 512, 512), (20304, 1536, 512), (20304, 512, 1408)]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().t()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    gemm(A, 0, B, 0, M, N, K, L.EPI_BF16, out); report(f"KK bf16 {M}x{N}x{K}", rel(out, ref))
    # B MN-major: stored [K, N]
    Bt = B.t().contiguous()
    out.zero_(); gemm(A, 0, Bt, 1, M, N, K, L.EPI_BF16, out); report(f"K,MN bf16 {M}x{N}x{K}", rel(out, ref))
    At = A.t().contiguous()
    out.zero_(); gemm(At, 1, Bt, 1, M, N, K, L.EPI_BF16, out); report(f"MN,MN bf16 {M}x{N}x{K}", rel(out, ref))
    out.zero_(); gemm(At, 1, B, 0, M, N, K, L.EPI_BF16, out); report(f"MN,K bf16 {M}x{N}x{K}", rel(out, ref))
    # fp32 + bias + alpha
    bias = torch.randn(N, device=dev)
    o32 = torch.zeros(1, M, N, device=dev)
    gemm(A, 0, B, 0, M, N, K, L.EPI_F32, o32, bias=bias, alpha=0.5); report(f"f32 bias alpha {M}x{N}x{K}", rel(o32[0], 0.5 * ref + bias), 1e-5)
    # residual
    res = torch.randn(M, N, device=dev); o2 = torch.zeros(M, N, device=dev); o2b = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    gemm(A, 0, B, 0, M, N, K, L.EPI_RESID, o2, aux0=res); report(f"resid {M}x{N}x{K}", rel(o2, ref + res), 1e-5)

# split-K dW shape: dW[N_w, K_w] = dY^T X, tokens = 20304
T = 20304
for (NW, KW, splits) in [(512, 512, 9), (2816, 512, 3), (512, 1408, 4)]:
    dY = torch.randn(T, NW, device=dev).bfloat16(); X = torch.randn(T, KW, device=dev).bfloat16()
    ref = dY.float().t() @ X.float()
    eff = lib.mca_gemm_effective_splits(T, splits)
    part = torch.zeros(eff, NW, KW, device=dev)
    gemm(dY, 1, X, 1, NW, KW, T, L.EPI_F32, part, splits=splits)
    report(f"dW splitK={eff} {NW}x{KW}", rel(part.sum(0), ref), 1e-5)

# GEGLU fwd / bwd, interleaved layout
M, I, IP, D = 1000, 1365, 1408, 512
x = torch.randn(M, D, device=dev).bfloat16()
W1 = (torch.randn(2 * I, D, device=dev) * 0.05)
W1i = torch.zeros(2 * IP, D, device=dev)
vidx = torch.arange(IP, device=dev); blk = vidx // 64; off = vidx % 64
rows_v = blk * 128 + off; rows_g = rows_v + 64
W1i[rows_v[:I]] = W1[:I]; W1i[rows_g[:I]] = W1[I:]
W1i = W1i.bfloat16()
u = torch.zeros(M, 2 * IP, device=dev, dtype=torch.bfloat16); h = torch.zeros(M, IP, device=dev, dtype=torch.bfloat16)
gemm(x, 0, W1i, 0, M, 2 * IP, D, L.EPI_GEGLU, h, out1=u)
uref = x.float() @ W1i.float().t()
val = uref[:, rows_v].clone().requires_grad_(True); gate = uref[:, rows_g].clone().requires_grad_(True)
href = torch.nn.functional.gelu(gate) * val
report("geglu h", rel(h, href))
report("geglu a=gelu(g)", rel(u.float()[:, rows_v], torch.nn.functional.gelu(gate)))
# bwd: dh = dx4 @ W2 (B MN-major = W2 stored [D, IP])
W2 = (torch.randn(D, IP, device=dev) * 0.05).bfloat16(); dx4 = torch.randn(M, D, device=dev).bfloat16()
du = torch.zeros(M, 2 * IP, device=dev, dtype=torch.bfloat16)
gemm(dx4, 0, W2, 1, M, IP, D, L.EPI_GEGLU_BWD, du, aux0=u)
dh = dx4.float() @ W2.float()
href.backward(dh)
report("geglu bwd dval", rel(du.float()[:, rows_v], val.grad)); report("geglu bwd dgate", rel(du.float()[:, rows_g], gate.grad))

# timing of the main shapes
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for (M, N, K) in [(20304, 1536, 512), (20304, 2816, 512), (20304, 512, 1408), (20304, 512, 512)]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ms = bench(lambda: gemm(A, 0, B, 0, M, N, K, L.EPI_BF16, out, sync=False))
    ms_t = bench(lambda: torch.matmul(A, B.t()))
    print(f"time bf16 {M}x{N}x{K}: ours {ms*1e3:.1f} us ({2*M*N*K/ms/1e9:.0f} TFLOP/s)  torch {ms_t*1e3:.1f} us ({2*M*N*K/ms_t/1e9:.0f} TFLOP/s)", flush=True)
    if N == 512:
        res = torch.randn(M, N, device=dev); o2 = torch.zeros(M, N, device=dev)
        ms = bench(lambda: gemm(A, 0, B, 0, M, N, K, L.EPI_RESID, o2, aux0=res, sync=False))
        print(f"time resid {M}x{N}x{K}: ours {ms*1e3:.1f} us ({2*M*N*K/ms/1e9:.0f} TFLOP/s, {(M*K*2+2*M*N*4)/ms/1e6:.0f} GB/s)", flush=True)
M = 20304
x = torch.randn(M, D, device=dev).bfloat16(); u = torch.zeros(M, 2 * IP, device=dev, dtype=torch.bfloat16); h = torch.zeros(M, IP, device=dev, dtype=torch.bfloat16)
ms = bench(lambda: gemm(x, 0, W1i, 0, M, 2 * IP, D, L.EPI_GEGLU, h, out1=u, sync=False))
print(f"time geglu fwd {M}x{2*IP}x{D}: {ms*1e3:.1f} us ({2*M*2*IP*D/ms/1e9:.0f} TFLOP/s)", flush=True)
dx4 = torch.randn(M, D, device=dev).bfloat16(); du = torch.zeros(M, 2 * IP, device=dev, dtype=torch.bfloat16)
ms = bench(lambda: gemm(dx4, 0, W2, 1, M, IP, D, L.EPI_GEGLU_BWD, du, aux0=u, sync=False))
print(f"time geglu bwd {M}x{IP}x{D}: {ms*1e3:.1f} us ({2*M*IP*D/ms/1e9:.0f} TFLOP/s)", flush=True)
dY = torch.randn(M, 2816, device=dev).bfloat16(); X = torch.randn(M, 512, device=dev).bfloat16()
eff = lib.mca_gemm_effective_splits(M, 3); part = torch.zeros(eff, 2816, 512, device=dev)
ms = bench(lambda: gemm(dY, 1, X, 1, 2816, 512, M, L.EPI_F32, part, splits=3, sync=False))
print(f"time dW 2816x512x{M} splitK={eff}: {ms*1e3:.1f} us ({2*M*2816*512/ms/1e9:.0f} TFLOP/s)", flush=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
