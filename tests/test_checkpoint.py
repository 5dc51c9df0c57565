"""Checkpoint layout: accelerate's save_state / load_state directory — model.safetensors,
optimizer.bin, scheduler.bin — as written by train_accel_gpu.py:122-123,134 and read by :97-99 / infer_accel_gpu.py:90-92.
CPU part: file formats and the learning-rate schedule against transformers' own scheduler; GPU part: resume parity."""
import os
import types

import pytest
import torch

from mca_paper_b200 import checkpoint as K, config as C, ops, synthetic as S
from mca_paper_b200.engine import Engine
from mca_paper_b200.model import MCA
from tests import helpers as H


def _fake_engine(lr, warmup, total, stride):
    eng = types.SimpleNamespace(adamw_cfg=ops.AdamWCfg(lr, 0.9, 0.999, 1e-8, 0.01, 2.0, 1, warmup, total, stride))
    eng.lr_at = lambda step: Engine.lr_at(eng, step)
    return eng


def _hf_scheduler(lr, warmup, total):
    from transformers import get_scheduler
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=lr)
    return opt, get_scheduler(name="cosine", optimizer=opt, num_warmup_steps=warmup, num_training_steps=total)


@pytest.mark.parametrize("stride", [1, 8])
def test_host_schedule_matches_transformers_cosine(stride):
    """lr used by optimiser step s == the reference's scheduler after (s-1) * num_processes scheduler steps
    (train_accel_gpu.py:81-86,119 under accelerate's prepared scheduler)."""
    lr, warmup, total = 1e-4, 12, 100
    eng = _fake_engine(lr, warmup, total, stride)
    opt, sch = _hf_scheduler(lr, warmup, total)
    for step in range(1, 14):
        assert abs(eng.lr_at(step) - opt.param_groups[0]["lr"]) <= 1e-12 + 1e-7 * lr, step
        opt.step()
        for _ in range(stride):
            sch.step()


@pytest.mark.parametrize("stride", [1, 2])
def test_scheduler_bin_loads_into_transformers_scheduler(tmp_path, stride):
    lr, warmup, total, step = 3e-4, 5, 40, 7
    eng = _fake_engine(lr, warmup, total, stride)
    torch.save(K.scheduler_state_dict(eng, step), tmp_path / K.SCHEDULER_NAME)
    opt, sch = _hf_scheduler(lr, warmup, total)
    ref_opt, ref = _hf_scheduler(lr, warmup, total)
    for _ in range(step * stride):
        ref_opt.step()
        ref.step()
    want = ref.state_dict()
    got = torch.load(tmp_path / K.SCHEDULER_NAME, weights_only=False)
    assert set(got) >= {"base_lrs", "last_epoch", "_step_count", "_last_lr", "lr_lambdas"}
    assert got["last_epoch"] == want["last_epoch"] and got["_step_count"] == want["_step_count"]
    assert abs(got["_last_lr"][0] - want["_last_lr"][0]) < 1e-9
    sch.load_state_dict(got)
    assert sch.last_epoch == step * stride and abs(sch.get_last_lr()[0] - ref.get_last_lr()[0]) < 1e-9
    opt.step()
    sch.step()
    ref_opt.step()
    ref.step()
    assert abs(sch.get_last_lr()[0] - ref.get_last_lr()[0]) < 1e-9


@pytest.mark.parametrize("safe", [True, False])
def test_model_file_roundtrip_cpu(tmp_path, safe):
    """model.safetensors / pytorch_model.bin hold exactly state_dict() (persistent buffers included) and load back,
    also when the keys carry DDP's `module.` prefix."""
    cfg = C.tiny_config("cmu", fcl=True)
    torch.manual_seed(0)
    a = MCA(**C.get_model_config(cfg))
    torch.manual_seed(1)
    b = MCA(**C.get_model_config(cfg))
    path = K.save_model(a, str(tmp_path), safe_serialization=safe)
    assert os.path.basename(path) == (K.MODEL_NAME if safe else K.MODEL_NAME_BIN)
    if safe:
        from safetensors.torch import load_file, save_file
        on_disk = load_file(path)
    else:
        on_disk = torch.load(path, weights_only=True)
    assert set(on_disk) == set(a.state_dict())
    res = K.load_model(b, str(tmp_path))
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    # DDP-prefixed file
    d2 = tmp_path / "ddp"
    d2.mkdir()
    pref = {"module." + k: v.clone() for k, v in a.state_dict().items()}
    if safe:
        save_file(pref, str(d2 / K.MODEL_NAME))
    else:
        torch.save(pref, str(d2 / K.MODEL_NAME_BIN))
    torch.manual_seed(2)
    c = MCA(**C.get_model_config(cfg))
    K.load_model(c, str(d2))
    for k, v in a.state_dict().items():
        assert torch.equal(v, c.state_dict()[k]), k


def test_load_model_missing_dir_raises(tmp_path):
    cfg = C.tiny_config("cmu", fcl=True)
    with pytest.raises(FileNotFoundError):
        K.load_model(MCA(**C.get_model_config(cfg)), str(tmp_path))


@pytest.mark.gpu
def test_save_state_load_state_resumes_identically(tmp_path):
    """3 steps -> save_state -> fresh model + trainer -> load_state -> the 4th step equals the uninterrupted run's; the
    files are what accelerate writes: optimizer.bin loads into a stock torch.optim.AdamW, scheduler.bin into
    transformers' scheduler."""
    from mca_paper_b200.trainer import Trainer
    dev = torch.device("cuda:0")
    cfg = C.tiny_config("cmu", fcl=True)
    kw = C.get_model_config(cfg)
    batch = S.make_batch(cfg, seed=1, variant="dropout_ragged")
    torch.manual_seed(0)
    model = MCA(**kw).to(dev)
    mk = lambda m: Trainer(m, lr=1e-3, clip=2.0, schedule="cosine", warmup_steps=2, total_steps=20, use_graphs=False)
    tr = mk(model)
    for _ in range(3):
        tr.step(batch)
    out = K.save_state(tr, str(tmp_path / "ckpt"))
    assert sorted(os.listdir(out)) == ["mca_b200_state_0.json", "model.safetensors", "optimizer.bin",
                                       "random_states_0.pkl", "scheduler.bin"]
    s4 = tr.step(batch).clone()
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}

    torch.manual_seed(123)
    model2 = MCA(**kw).to(dev)
    tr2 = mk(model2)
    tr2.step(batch)                      # make it dirty: load_state must overwrite weights, moments and the step
    assert K.load_state(tr2, out) == 3
    s4b = tr2.step(batch).clone()
    assert abs(float(s4[0]) - float(s4b[0])) <= 2e-3 * abs(float(s4[0]))          # atomics order only
    for k in after:
        if after[k].dtype.is_floating_point:
            assert H.rel_err(model2.state_dict()[k], after[k]) < 1e-3, k
    ref_opt = torch.optim.AdamW(model2.parameters(), lr=1e-3)
    ref_opt.load_state_dict(torch.load(os.path.join(out, K.OPTIMIZER_NAME), weights_only=True))
    _, sch = _hf_scheduler(1e-3, 2, 20)
    sch.load_state_dict(torch.load(os.path.join(out, K.SCHEDULER_NAME), weights_only=False))
    assert sch.last_epoch == 3
    # a checkpoint written at another world size / scheduler stride is refused rather than silently re-timed
    tr3 = Trainer(model2, lr=1e-3, clip=2.0, schedule="cosine", warmup_steps=2, total_steps=20, use_graphs=False,
                  scheduler_stride=4)
    with pytest.raises(ValueError):
        K.load_state(tr3, out)


@pytest.mark.gpu
def test_device_schedule_matches_transformers_cosine():
    """With zero gradients AdamW's update is p *= 1 - lr * weight_decay, so the learning rate the kernel used can be read
    back exactly; compare it with transformers' cosine schedule advanced `stride` times per step (accelerate)."""
    import ctypes
    call, P = ops.call, ops.P
    dev = torch.device("cuda:0")
    lr, warmup, total, stride, wd = 1e-2, 6, 40, 3, 0.5
    n = 4096
    p = torch.ones(n, device=dev)
    g = torch.zeros(n, device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, device=dev, dtype=torch.int64)
    sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
    tn = torch.zeros(1, device=dev)
    cfg = ops.AdamWCfg(lr, 0.9, 0.999, 1e-8, wd, 2.0, 1, warmup, total, stride)
    opt, sch = _hf_scheduler(lr, warmup, total)
    for it in range(1, 13):
        before = p[0].item()
        call("mca_clip_adamw_step", P(p), P(g), P(m), P(v), n, P(sumsq), P(step), P(tn), 1.0, ctypes.addressof(cfg),
             torch.cuda.current_stream().cuda_stream)
        used = (1.0 - p[0].item() / before) / wd
        want = opt.param_groups[0]["lr"]
        assert abs(used - want) <= 5e-7, (it, used, want)      # fp32 rounding of p bounds the read-back at ~1e-7
        opt.step()
        for _ in range(stride):
            sch.step()
