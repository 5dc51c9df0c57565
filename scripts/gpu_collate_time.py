"""Device collate at CMU_config1 shape: time of DeviceCollator (pinned varlen staging + H2D +
expand kernels) against the reference algorithm on the host (oracle/collate_oracle.py port of MultimodalCollator, then a
dense pinned H2D), for full-length and for 40 %-dropout ragged samples; kernel-only time and GB/s of mca_collate_rows."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.collate import DeviceCollator
from mca_paper_b200.ops import P, call
from oracle import collate_oracle as CO

dev = "cuda"
cfg = C.named_config("CMU_config1")
B = cfg["batch_size"]
mc = {n: {"type": "embedded_sequence", "pad_len": e["max_tokens"], "data_col_name": "data", "pad_token": -10000,
          "embedding_size": e["input_size"]} for n, e in cfg["encoder_configs"].items()}
res = {}
for variant in ("full", "dropout_ragged"):
    dense = S.make_batch(cfg, seed=1, variant=variant)
    samples = []
    for b in range(B):
        s = {}
        for n in mc:
            live = int((~dense[n]["attention_mask"][b]).sum())
            s[n] = {"data": dense[n]["tokens"][b, :live].clone() if live else None}
        samples.append(s)
    col = DeviceCollator(mc, B, dev)
    for _ in range(3):
        col(samples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        col(samples)
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / 20
    # reference algorithm on the host + dense H2D from pinned memory
    pinned = None
    def host_path():
        global pinned
        out = CO.multimodal_collate(mc, samples)
        if pinned is None:
            pinned = {m: {k: torch.empty_like(v).pin_memory() for k, v in d.items()} for m, d in out.items()}
            host_path.devb = {m: {k: torch.empty_like(v, device=dev) for k, v in d.items()} for m, d in out.items()}
        for m, d in out.items():
            for k, v in d.items():
                pinned[m][k].copy_(v)
                host_path.devb[m][k].copy_(pinned[m][k], non_blocking=True)
    for _ in range(2):
        host_path()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        host_path()
    torch.cuda.synchronize()
    t_host = (time.perf_counter() - t0) / 10
    pinned = None
    # kernel only: OpenFace modality (450 x 713), rows already on the device
    n = "OpenFace"
    rows, off = col._buf[n + ".rows"][1], col._buf[n + ".off"][1]
    tok, msk = col._buf[n + ".tokens"], col._buf[n + ".mask"]
    L, E = mc[n]["pad_len"], mc[n]["embedding_size"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        call("mca_collate_rows", P(rows), P(off), B, L, E, 0.0, 1, P(tok), P(msk), torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    live_rows = int(off[-1].item())
    alg_bytes = live_rows * E * 4 + B * L * E * 4 + B * L
    res[variant] = {"device_collate_ms": t_dev * 1e3, "host_collate_plus_dense_h2d_ms": t_host * 1e3,
                    "h2d_bytes_device_path": col.h2d_bytes, "h2d_bytes_dense": sum(v.numel() * v.element_size() for d in dense.values() for v in d.values()),
                    "kernel_openface_us": us, "kernel_openface_gbs": alg_bytes / us / 1e3, "kernel_openface_algorithmic_bytes": alg_bytes}
print(json.dumps(res, indent=1))
