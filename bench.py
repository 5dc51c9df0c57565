#!/usr/bin/env python
"""bench.py — train samples/s of the MCA hot path at CMU_config1 shape (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU implementation of the same step (oracle port)

A step = one full training pass over one synthetic CMU_config1 batch (B = 8 per rank, all four modalities at full
length): encoders -> 5 x (LN, QKV, masked attention, out-proj, LN, GEGLU-FF) -> LN -> attention
pooling -> 14-pair InfoNCE -> full backward -> (DP all-reduce) -> clip 2.0 + AdamW + cosine LR -> weight re-pack.
`value` times K steps with the batch resident in HBM; `e2e` times the same K steps from HOST buffers (pinned ->
device copy of the 14.8 MB batch and a device -> host read of the loss inside the timed region, every step).
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec at CMU_config1 shape"
UNIT = "samples/s"

# stdout carries exactly ONE JSON line.  Libraries print to it behind Python's back (NCCL's "NCCL version ..." banner on the
# first communicator, torchrun's OMP notice): file descriptor 1 is pointed at stderr for the whole run and the line is
# written to a duplicate of the real stdout.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm_sorted = sorted(sm)
            out["sm_mhz"] = statistics.median(sm_sorted[len(sm_sorted) // 2:])  # median of the busier half = under load
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------------------------- reference arm
def _to_dev(obj, dev):
    if isinstance(obj, torch.Tensor):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: _to_dev(v, dev) for k, v in obj.items()}
    return obj


def cpu_training_step_fn(cfg_name: str, batch_size: int, device: str = "cpu", variant: str = "full", seed: int = 1):
    """The reference algorithm (oracle port, oracle/mca_oracle.py: plain torch fp32 ops + autograd) as a training step
    closure.  device="cpu": the reference's CPU path (cpu_baseline / --impl reference); device="cuda": the same eager
    fp32 program on the B200 (eager_gpu_baseline, the bar SURVEY.md §6 names)."""
    from mca_paper_b200 import config as C, synthetic as S
    from oracle import mca_oracle as O

    cfg = C.named_config(cfg_name)
    cfg["batch_size"] = batch_size
    kw = C.get_model_config(cfg)
    from mca_paper_b200.model import EAO, MCA

    torch.manual_seed(43)
    model = (EAO if kw.get("eao") else MCA)(**kw)  # parameter container only (CPU); arithmetic below is the oracle's
    oracle_forward = O.eao_forward if kw.get("eao") else O.mca_forward
    sd = {k: v.detach().clone().to(device) for k, v in model.state_dict().items()}
    names = [k for k, _ in model.named_parameters()]
    params = [sd[k].requires_grad_(True) for k in names]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    tables = _to_dev(O.static_tables(kw), device)
    batch = _to_dev(S.make_batch(cfg, seed=seed, variant=variant, batch_size=batch_size), device)
    state = {"step": 0}

    def step():
        state["step"] += 1
        for p in params:
            p.grad = None
        out = oracle_forward(sd, kw, batch, tables=tables)
        out["loss"].backward()
        with torch.no_grad():
            O.clip_adamw_step([p for p in params], [p.grad if p.grad is not None else torch.zeros_like(p) for p in params],
                              m, v, state["step"], lr=1e-4, max_norm=2.0)
        return float(out["loss"])

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bs = 8  # the configured batch (configs/CMU_config1.yaml:14): the same step the B200 arm times
    step = cpu_training_step_fn(args.config, bs, variant=args.variant)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps * bs / dt
    sample = (f"every step = one full training step (fwd+bwd+clip+AdamW, fp32, torch CPU ops on {cores} threads) on the same "
              f"{bs}-sample {args.config} batch the B200 arm uses")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.config, args.variant), "global_batch": bs, "batch_per_step": bs},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference cannot be installed (torchmultimodal/yacs/accelerate absent, no network); this arm times "
                "oracle/mca_oracle.py, the CPU restatement pinned against the live reference (tests/golden)",
    }
    emit(line)


def workload_string(cfg_name: str, variant: str) -> str:
    """Workload description from the config itself (modalities, shapes, losses), identical for both arms."""
    from mca_paper_b200 import config as C
    from mca_paper_b200.plan import EAOPlan, StaticPlan

    cfg = C.named_config(cfg_name)
    kw = C.get_model_config(cfg)
    enc = kw["encoder_configs"]
    mods = ", ".join(f"{n} {e['max_tokens']}x{e.get('input_size', 1)}" for n, e in enc.items())
    if kw.get("eao"):
        pl = EAOPlan(enc, list(kw.get("fusion_combos", (4, 5))), kw.get("fcl", False), kw.get("zorro", False),
                     kw.get("no_fusion", True), kw.get("bimodal_contrastive", False), kw.get("non_fusion_fcl", False))
        return (f"{cfg_name} EAO training step, B={cfg['batch_size']} per GPU ({mods}; {pl.R} passes stacked "
                f"block-diagonally -> N={pl.N} tokens per sample, d=512, {kw['depth']} layers, mean pooling, {pl.n_pairs} InfoNCE "
                f"pairs), variant={variant}")
    pl = StaticPlan(enc, kw.get("num_fusion_tokens", 16), list(kw.get("fusion_combos", (4, 5))), kw.get("fcl", False),
                    kw.get("zorro", False), kw.get("no_fusion", False), kw.get("bimodal_contrastive", False),
                    kw.get("non_fusion_fcl", False))
    kind = "MMA" if kw.get("zorro") else "MCA"
    return (f"{cfg_name} {kind} training step, B={cfg['batch_size']} per GPU ({mods}; {pl.F} fusion tokens -> N={pl.N}, d=512, "
            f"{kw['depth']} layers, {pl.n_pairs} InfoNCE pairs), variant={variant}")


def live_allowed_pairs(plan, host_batch) -> float:
    """Mean over the batch of the (query, key) pairs parity needs: statically allowed AND key not padded (every query row
    is computed, padded ones included: SURVEY.md H1).  Equals plan.allowed_pairs for an unpadded batch."""
    import numpy as np

    pads = []
    for name in plan.names:
        d = host_batch[name]
        if "attention_mask" in d:
            pads.append(d["attention_mask"].to(torch.bool).numpy())
        else:  # PatchEncoder derives its own mask: treat as live
            pads.append(np.zeros((next(iter(d.values())).shape[0], plan.lengths[plan.names.index(name)]), dtype=bool))
    B = pads[0].shape[0]
    if getattr(plan, "eao", False):
        return float(plan.allowed_pairs)
    pad = np.concatenate(pads + [np.zeros((B, plan.F), dtype=bool)], axis=1)       # [B, N]
    allowed_per_key = (~plan.attn_mask).sum(axis=0).astype(np.float64)              # queries that may see key k
    return float(((~pad) * allowed_per_key[None, :]).sum() / B)


def ncu_traffic(entry_point: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the entry point's main kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by scripts/ncu_traffic.py); None when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(entry_point, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- B200 arm
def algorithmic_flops(plan, B, H, depth, I):
    """only mask-allowed (q,k) pairs at full length, 2 FLOP/MAC, backward = 2x forward."""
    M = B * plan.N
    d = 512
    lin = 2 * M * d * (3 * d) + 2 * M * d * d + 2 * M * d * (2 * I) + 2 * M * I * d
    attn = plan.allowed_pairs * B * H * 4 * 64
    return {"linear_layer_fwd": lin, "attn_layer_fwd": attn, "qkv": 2 * M * d * 3 * d, "ff1": 2 * M * d * 2 * I,
            "ff2": 2 * M * I * d, "out": 2 * M * d * d}


def run_infer(args):
    """Config 5 (infer_accel_gpu.py:88-111): eval-mode forward of the MMA model incl. its losses, every returned embedding
    read back to the host; N ranks = N independent replicas on disjoint shards (the script asserts world_size == 1)."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from mca_paper_b200 import config as C, synthetic as S
    from mca_paper_b200.model import MCA

    name = args.config if args.config != "CMU_config1" else "CMU_config1_z_12i"
    cfg = C.named_config(name)
    torch.manual_seed(int(cfg["seed"]))
    model = MCA(**C.get_model_config(cfg)).to(dev).eval()
    eng = model.engine
    eng.ensure_flat()
    eng.pack_weights()
    host = S.make_batch(cfg, seed=1 + rank, variant=args.variant)
    pinned = {m: {k: v.pin_memory() for k, v in d.items()} for m, d in host.items()}
    devb = {m: {k: torch.empty_like(v, device=dev) for k, v in d.items()} for m, d in host.items()}
    out_pin = torch.empty(eng.B, eng.R, 512, pin_memory=True)

    def fwd():
        pooled = eng.trunk_forward(devb)
        eng.loss_forward(pooled)

    def h2d():
        for m, d in pinned.items():
            for k, v in d.items():
                devb[m][k].copy_(v, non_blocking=True)

    h2d()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        fwd(), fwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        fwd()
    for _ in range(max(args.warmup, 3)):
        g.replay()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    for _ in range(args.steps):
        g.replay()
    e[1].record()
    torch.cuda.synchronize()
    e[2].record()
    for _ in range(args.steps):
        h2d()
        g.replay()
        out_pin.copy_(eng.ws["pooled"], non_blocking=True)
    e[3].record()
    torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        h2d_bytes = sum(v.numel() * v.element_size() for d in host.values() for v in d.values())
        emit(({
            "metric": "embedding inference samples/sec (infer_accel_gpu.py forward incl. losses)", "unit": UNIT, "n_gpus": world,
            "value": world * eng.B * args.steps / (float(t[0]) * 1e-3), "steps": args.steps, "ms_per_step": float(t[0]) / args.steps,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{name} MMA forward, eval/no_grad, B=8 per replica, variant={args.variant}",
                       "parallelism": f"{world} independent replicas (no collective)"},
            "e2e": {"value": world * eng.B * args.steps / (float(t[1]) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": out_pin.numel() * 4, "ms_per_step": float(t[1]) / args.steps}}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="CMU_config1")
    ap.add_argument("--variant", default="full")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--varlen", default=None, choices=["exact", "fast", "off"],
                    help="varlen query skipping of the attention kernels (default: the engine's, MCA_VARLEN or 'exact')")
    ap.add_argument("--profile-step", action="store_true",
                    help="for ncu --profile-from-start off: warm up, then bracket exactly ONE eager training step with "
                         "cudaProfilerStart/Stop and exit (no timing, no JSON line: a number taken under a profiler is not a bench value)")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="infer = SURVEY.md §8d config 5: eval() forward + losses of infer_accel_gpu.py:97-111, replicas only")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "infer":
        return run_infer(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)

    from mca_paper_b200 import config as C, ops, synthetic as S
    from mca_paper_b200.model import EAO, MCA
    from mca_paper_b200.trainer import Trainer

    cfg = C.named_config(args.config)
    kw = C.get_model_config(cfg)
    torch.manual_seed(int(cfg["seed"]))
    model = (EAO if kw.get("eao") else MCA)(**kw).to(dev)   # train_accel_gpu.py:47-52
    trainer = Trainer(model, lr=float(cfg["lr"]), clip=float(cfg["clip"]), schedule="cosine",
                      warmup_steps=int(cfg["num_warmup_steps"]), total_steps=100000, use_graphs=not args.no_graphs)
    eng = trainer.eng
    if args.varlen is not None:
        eng.set_varlen(args.varlen)
    host_batch = S.make_batch(cfg, seed=1 + rank, variant=args.variant)
    B = eng.B
    live_tokens = 1.0 - float(sum(float(d["attention_mask"].to(torch.bool).sum()) for d in host_batch.values()
                                  if "attention_mask" in d)) / float(B * eng.N)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the graphs)
    ops.COUNT["n"] = 0
    trainer.stage(host_batch)
    trainer.step_staged()
    launches_per_step = ops.COUNT["n"] // (3 if not args.no_graphs else 1)  # 2 warm-up passes + 1 capture pass
    for _ in range(args.warmup):
        trainer.stage(host_batch)
        trainer.step_staged()
    torch.cuda.synchronize()
    if args.profile_step:
        torch.cuda.cudart().cudaProfilerStart()
        trainer._run_eager()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    loss_pinned = torch.empty(4, pin_memory=True)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- timed region 1: inputs resident in HBM
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        trainer.step_staged()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    # ---- timed region 2: end to end from host buffers.  Every step's batch is copied pinned-host -> device inside the
    # timed region (on a copy stream, one step ahead of the compute) and every step's loss summary is read back to
    # the host inside it (the read of step i completes while step i+1 runs).
    loss_pin2 = [torch.empty(4, pin_memory=True) for _ in range(2)]
    loss_evt = [torch.cuda.Event() for _ in range(2)]
    trainer._ensure_pipeline()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    trainer.prefetch(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            trainer.prefetch((i + 1) & 1)
        summary = trainer.step_from_slot(i & 1)
        loss_pin2[i & 1].copy_(summary, non_blocking=True)
        loss_evt[i & 1].record()
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
    loss_evt[(args.steps - 1) & 1].synchronize()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    loss_pinned = loss_pin2[(args.steps - 1) & 1]
    clocks = sampler.stop() if sampler is not None else None
    final_loss = float(loss_pinned[0])
    h2d_bytes = trainer.h2d_bytes

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel.  Every entry-point call of one eager pass of the same step is recorded,
    # then each kernel family is re-launched back to back on the device (all of its calls of the step, in order, so
    # consecutive launches touch different layers' buffers: > L2) between two CUDA events.  This removes the host
    # launch gaps that per-call events in an eager pass include.
    roof, per_kernel = None, None
    ops.RECORD["on"], ops.RECORD["calls"] = True, []
    ops.PROFILE["on"], ops.PROFILE["events"] = world > 1, []
    trainer._run_eager()  # on every rank: the step contains the three collectives
    torch.cuda.synchronize()
    ops.RECORD["on"] = False
    ops.PROFILE["on"] = False
    # exchange kernels wait for their peers, so they cannot be replayed on one rank: they are timed in place (CUDA events
    # around each call of the eager pass; the time includes waiting for the slowest rank = the exposed exchange time)
    exchange = {}
    for name, tag, ea, eb in ops.PROFILE["events"]:
        if name.startswith(("mca_dp_", "mca_p2p_", "mca_xgpu_barrier", "mca_contrastive_")):
            d = exchange.setdefault(name, {"ms_total_per_step": 0.0, "launches_per_step": 0})
            d["ms_total_per_step"] += ea.elapsed_time(eb)
            d["launches_per_step"] += 1
    for d in exchange.values():
        d["us_avg"] = d["ms_total_per_step"] / d["launches_per_step"] * 1e3
    if rank == 0:
        calls = ops.RECORD["calls"]
        groups = {}
        for name, tag, a in calls:
            groups.setdefault(f"{name}:{tag}" if tag else name, []).append((name, a))
        replay_ok = ("mca_gemm_bf16", "mca_attn_fwd", "mca_attn_bwd", "mca_layernorm512_fwd", "mca_layernorm512_bwd",
                     "mca_add_layernorm512_fwd", "mca_batchsum_rows", "mca_broadcast_rows", "mca_cast_f32_bf16",
                     "mca_pool_attn_fwd", "mca_pool_attn_bwd", "mca_small_gemm_f32", "mca_colsum", "mca_pack_weights",
                     "mca_unpack_grads", "mca_layernorm_in_fwd", "mca_layernorm_in_param_bwd", "mca_build_offsets")
        per_kernel = {}
        for key, lst in groups.items():
            if not key.startswith(replay_ok):
                continue
            reps = max(1, 8 // len(lst))
            for name, a in lst:  # warm
                ops.fn(name)(*a)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                for name, a in lst:
                    ops.fn(name)(*a)
            a1.record()
            torch.cuda.synchronize()
            ms = a0.elapsed_time(a1) / (reps * len(lst))
            per_kernel[key] = {"ms_total_per_step": ms * len(lst), "launches_per_step": len(lst), "us_avg": ms * 1e3}
        for k, v in exchange.items():
            per_kernel[k + " (in place)"] = v
        per_kernel = dict(sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms_total_per_step"]))
        if os.environ.get("MCA_BENCH_TABLE"):
            os.makedirs(os.path.dirname(os.environ["MCA_BENCH_TABLE"]) or ".", exist_ok=True)
            with open(os.environ["MCA_BENCH_TABLE"], "w") as f:
                json.dump(per_kernel, f, indent=1)
        peaks = read_peaks()
        fl = algorithmic_flops(eng.plan, B, eng.H, eng.depth, eng.I)
        live_pairs = live_allowed_pairs(eng.plan, host_batch)  # = plan.allowed_pairs for the unpadded workload
        fl["attn_layer_fwd"] = live_pairs * B * eng.H * 4 * 64
        top = next(k for k in per_kernel if not k.endswith("(in place)"))
        tname = top.split(":")[0]
        flops_map = {"mca_attn_bwd": 2 * fl["attn_layer_fwd"], "mca_attn_fwd": fl["attn_layer_fwd"]}
        if tname == "mca_gemm_bf16":
            tg = top.split(":")[1].split("_")
            gm, gn, gk = int(tg[1][1:]), int(tg[2][1:]), int(tg[3][1:])
            flops_map["mca_gemm_bf16"] = 2.0 * gm * gn * gk
        if tname in flops_map:
            f = flops_map[tname]
            dur = per_kernel[top]["us_avg"] * 1e-6
            ach = f / dur / 1e12
            # the kernel is timed ALONE (back-to-back launches of this one kernel, a few ms): the burst cuBLAS figure is the
            # denominator; the fraction of the sustained figure (what a seconds-long loop under the power cap reaches) is
            # stated beside it
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops"], "frac_of_sustained_peak": ach / peaks["bf16_tflops_sustained"],
                    "live_allowed_pairs_per_sample_head": live_pairs, "static_allowed_pairs": eng.plan.allowed_pairs,
                    "traffic": ncu_traffic(tname) if args.config == "CMU_config1" else None,  # captured on that workload
                    "algorithmic_flops_per_launch": f, "avg_launch_us": per_kernel[top]["us_avg"],
                    "launches_per_step": per_kernel[top]["launches_per_step"],
                    "peak_source": peaks["source"] + " (bf16_tflops, burst: kernel timed in isolation)",
                    "how": "CUDA events around back-to-back device launches of every call of this kernel in one step "
                           "(recorded from an eager pass of the same step, same process); attention FLOPs count only "
                           "mask-allowed (q,k) pairs whose key is live in this batch; the launch also covers its memset / prep / "
                           "cast helpers"}
        else:
            roof = {"kernel": top, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None,
                    "traffic": None, "avg_launch_us": per_kernel[top]["us_avg"], "peak_source": peaks["source"]}
        gemm_ms = sum(v["ms_total_per_step"] for k, v in per_kernel.items() if k.startswith("mca_gemm_bf16"))
        roof["gemm_family"] = {"ms_per_step": gemm_ms, "algorithmic_tflops": 3 * eng.depth * fl["linear_layer_fwd"] / (gemm_ms * 1e-3) / 1e12}
        step_flops = 3 * (eng.depth * (fl["linear_layer_fwd"] + fl["attn_layer_fwd"]))
        roof["step_algorithmic_tflops"] = step_flops / (ms_dev / args.steps * 1e-3) / 1e12

    # ---- the fp32-parity forward mode of the same step (north_star's 1e-3 bar; Engine.set_precision), N = 1 only: eager
    # launches (the captured graph holds the bf16 path), CUDA events around 5 steps after 2 warm-up
    fp32_mode = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not kw.get("eao"):
        try:
            eng.set_precision("fp32")
            for _ in range(2):
                trainer._run_eager()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(5):
                trainer._run_eager()
            f1.record()
            torch.cuda.synchronize()
            ms = f0.elapsed_time(f1) / 5
            fp32_mode = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                         "what": "same training step with the fp32-parity forward (3-term bf16 split products on the tcgen05 GEMM, "
                                 "fp32 attention / GEGLU / pooling; loss and embeddings within 1e-3 of the fp32 reference: "
                                 "tests/test_gpu_exact.py), regular bf16 backward; eager launches"}
        except Exception as exc:
            fp32_mode = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        finally:
            eng.set_precision("bf16")

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        bs, n_timed = B, 2
        step = cpu_training_step_fn(args.config, bs, variant=args.variant)
        step()  # warm-up (allocator, thread pool)
        t0 = time.perf_counter()
        for _ in range(n_timed):
            step()
        dt = (time.perf_counter() - t0) / n_timed
        cpu_base = {"value": bs / dt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{n_timed} timed CPU training steps (fwd+bwd+clip+AdamW, fp32, after 1 warm-up) on the same {bs}-sample "
                              f"batch (same config, same batch size as the B200 arm), oracle/mca_oracle.py"}
    # ---- eager fp32 baseline on the SAME B200 (SURVEY.md §6 / §8d "the real bar"): the oracle port (= the reference's
    # op sequence: cuBLAS fp32 SIMT GEMMs with TF32 off, materialised [B,h,N,N] scores, ATen softmax / layer_norm) on cuda
    eager_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            del trainer
            torch.cuda.empty_cache()
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            step = cpu_training_step_fn(args.config, B, device="cuda", variant=args.variant)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_g = 5
            g0.record()
            for _ in range(n_g):
                step()
            g1.record()
            torch.cuda.synchronize()
            ms = g0.elapsed_time(g1) / n_g
            eager_gpu = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "dtype": "f32 (TF32 off, the reference default)",
                         "kind": "port", "sample": f"{n_g} timed steps after 2 warm-up, same {B}-sample batch, torch eager on cuda:0 "
                                                   "(oracle/mca_oracle.py + torch autograd + clip + AdamW)",
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        except Exception as exc:  # an out-of-memory eager baseline must not lose the measured line
            eager_gpu = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    if rank == 0:
        value = world * B * args.steps / (ms_dev * 1e-3)
        e2e = world * B * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_string(args.config, args.variant),
                       "global_batch": world * B, "parallelism": f"dp{world}", "cuda_graphs": not args.no_graphs,
                       "dp_exchange": ("none" if world == 1 else ("peer memory (push/pull kernels + flag barriers, one graph)"
                                                                 if eng._p2p is not None else "nccl")),
                       "varlen": eng.varlen, "live_token_fraction": live_tokens,
                       "l2": "per-step working set (~2 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "precision": "bf16 tensor-core operands, fp32 accumulate, fp32 master weights/residual stream/LN/softmax/loss"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "every step copies its batch from PINNED host buffers (a copy stream, one step ahead) and reads its "
                            "loss summary back to pinned host memory inside the timed region; the pageable -> pinned memcpy a "
                            "DataLoader worker would do is outside it"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "cpu_baseline": cpu_base, "eager_gpu_baseline": eager_gpu, "fp32_parity_mode": fp32_mode, "final_loss": final_loss,
            "per_kernel_ms_per_step": {k: round(v["ms_total_per_step"], 4) for k, v in list(per_kernel.items())[:16]} if per_kernel else None,
        }
        emit(line)
    if world > 1:
        eng.check_p2p()  # raises if a peer GPU missed one of the flag barriers (would invalidate the number)
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
