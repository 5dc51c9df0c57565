"""Checkpoints in the directory layout of accelerate's `save_state` / `load_state`, which is what the reference writes
every epoch and restarts / infers from (train_accel_gpu.py:97-99,122-123,134; infer_accel_gpu.py:90-92; the published
checkpoints of README.md:44-53 are tarballs of such directories) — SURVEY.md §8(f) rank 2:

    <dir>/model.safetensors     model.state_dict()  (older accelerate: pytorch_model.bin, read as well)
    <dir>/optimizer.bin         torch.optim.AdamW.state_dict() of the one parameter group, parameters in
                                named_parameters() order
    <dir>/scheduler.bin         state_dict() of the LambdaLR that transformers.get_scheduler builds
    <dir>/random_states_<r>.pkl per-rank RNG states (written for completeness, only `step` is read back)

accelerate itself is not needed (it is not installed here); the files are plain safetensors / torch.save payloads.
The fused step keeps weights, moments and the step counter in flat device buffers (engine.Engine); the (un)flattening
to the per-parameter layout happens here, outside the timed path.  Under peer-memory data parallelism each rank holds the
moments of its shard only: `save_state` gathers them (every rank must call it), rank 0 writes.
"""
from __future__ import annotations

import json
import os
import pickle
import random
from typing import Dict, Optional

import torch

MODEL_NAME = "model.safetensors"
MODEL_NAME_BIN = "pytorch_model.bin"
OPTIMIZER_NAME = "optimizer.bin"
SCHEDULER_NAME = "scheduler.bin"
RNG_NAME = "random_states_{rank}.pkl"
STATE_NAME = "mca_b200_state_{rank}.json"   # plain-data side file: optimiser step + dropout-stream position


def scheduler_state_dict(eng, step: int) -> Dict:
    """What `lr_scheduler.state_dict()` of the reference loop holds after `step` optimiser steps: a LambdaLR that
    has been advanced step * stride times (accelerate steps it once per process)."""
    stride = max(1, int(eng.adamw_cfg.sched_stride))
    return {"base_lrs": [float(eng.adamw_cfg.lr)], "last_epoch": step * stride, "_step_count": step * stride + 1,
            "_is_initial": False, "_get_lr_called_within_step": False, "_last_lr": [eng.lr_at(step + 1)],
            "lr_lambdas": [{}]}


def model_state_dict(model) -> Dict[str, torch.Tensor]:
    """CPU copy of state_dict() (the nn.Parameters are views of the engine's flat fp32 buffer, always current)."""
    return {k: v.detach().to("cpu").contiguous() for k, v in model.state_dict().items()}


def save_model(model, output_dir: str, safe_serialization: bool = True) -> str:
    os.makedirs(output_dir, exist_ok=True)
    sd = model_state_dict(model)
    if safe_serialization:
        from safetensors.torch import save_file
        path = os.path.join(output_dir, MODEL_NAME)
        save_file({k: v.clone() for k, v in sd.items()}, path, metadata={"format": "pt"})
    else:
        path = os.path.join(output_dir, MODEL_NAME_BIN)
        torch.save(sd, path)
    return path


def load_model(model, input_dir: str, strict: bool = True):
    """Load model.safetensors (or pytorch_model.bin) written by accelerate / by save_model into `model`; keys of a
    DistributedDataParallel-wrapped save ("module." prefix) are accepted."""
    p = os.path.join(input_dir, MODEL_NAME)
    if os.path.exists(p):
        from safetensors.torch import load_file
        sd = load_file(p)
    else:
        p = os.path.join(input_dir, MODEL_NAME_BIN)
        if not os.path.exists(p):
            raise FileNotFoundError(f"no {MODEL_NAME} or {MODEL_NAME_BIN} under {input_dir}")
        sd = torch.load(p, map_location="cpu", weights_only=True)
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
    res = model.load_state_dict(sd, strict=strict)
    eng = getattr(model, "engine", None)
    if eng is not None and getattr(eng, "flat", None) is not None:
        eng.ensure_flat()     # parameters may have been re-assigned
        eng.pack_weights()    # refresh the bf16 kernel-layout copies the fused step reads
    return res


def save_state(trainer, output_dir: str, safe_serialization: bool = True) -> Optional[str]:
    """accelerator.save_state(output_dir) for the fused trainer.  Collective under data parallelism (the optimiser
    shards are gathered); only rank 0 writes the model / optimiser / scheduler files."""
    eng = trainer.eng
    eng.check_p2p()  # never checkpoint a run in which a peer missed an exchange barrier
    opt = trainer.optimizer_state_dict()  # gathers the shards: every rank takes part
    step = int(float(opt["state"][0]["step"])) if opt["state"] else 0
    os.makedirs(output_dir, exist_ok=True)
    rng = {"step": step, "random_state": random.getstate(), "torch_manual_seed": torch.get_rng_state(),
           "drop_ctr": int(eng.ws["drop_ctr"].item())}  # position of the counter-based dropout stream
    try:
        import numpy as np
        rng["numpy_random_seed"] = np.random.get_state()
    except Exception:  # numpy is optional for the product path
        pass
    if torch.cuda.is_available():
        rng["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
    # accelerate's per-rank RNG file, written for layout compatibility only: load_state never unpickles it
    with open(os.path.join(output_dir, RNG_NAME.format(rank=eng.rank)), "wb") as f:
        pickle.dump(rng, f)
    # what THIS trainer needs back on resume, as plain JSON (no code execution on load)
    with open(os.path.join(output_dir, STATE_NAME.format(rank=eng.rank)), "w") as f:
        json.dump({"step": step, "drop_ctr": rng["drop_ctr"]}, f)
    if eng.rank != 0:
        return None
    save_model(trainer.model, output_dir, safe_serialization)
    torch.save(opt, os.path.join(output_dir, OPTIMIZER_NAME))
    torch.save(scheduler_state_dict(eng, step), os.path.join(output_dir, SCHEDULER_NAME))
    return output_dir


def load_state(trainer, input_dir: str, strict: bool = True) -> int:
    """accelerator.load_state(input_dir): weights, AdamW moments + step, and the scheduler position (checked against the
    optimiser's step: the device schedule is a function of that one counter).  Returns the restored step count."""
    eng = trainer.eng
    load_model(trainer.model, input_dir, strict)
    p = os.path.join(input_dir, OPTIMIZER_NAME)
    step = 0
    if os.path.exists(p):
        opt = torch.load(p, map_location="cpu", weights_only=True)
        trainer.load_optimizer_state_dict(opt)
        steps = [int(float(s["step"])) for s in opt["state"].values()]
        step = max(steps) if steps else 0
    p = os.path.join(input_dir, SCHEDULER_NAME)
    if os.path.exists(p):
        sch = torch.load(p, map_location="cpu", weights_only=True)   # a plain dict of numbers / lists (no pickled code)
        stride = max(1, int(eng.adamw_cfg.sched_stride))
        if int(sch.get("last_epoch", step * stride)) != step * stride:
            raise ValueError(f"scheduler.bin is at scheduler step {sch.get('last_epoch')} but the optimiser state is at "
                             f"step {step} x {stride} scheduler steps per step: resume with the world size (or "
                             "scheduler_stride) the checkpoint was written with")
    # the dropout-stream position comes from our own JSON side file; random_states_<rank>.pkl (accelerate's pickled
    # host RNG states) is deliberately NOT read: unpickling a downloaded checkpoint would execute arbitrary code, and the
    # fused step draws nothing from the host generators
    p = os.path.join(input_dir, STATE_NAME.format(rank=eng.rank))
    if os.path.exists(p):
        with open(p) as f:
            st = json.load(f)
        eng.ws["drop_ctr"].fill_(int(st.get("drop_ctr", 0)))
    return step
