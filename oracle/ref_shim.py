"""TEST INFRASTRUCTURE ONLY — loader for the *live* reference model (never imported by the product path).

The reference (/root/reference, read-only, only present in the build container) cannot be imported as shipped:
model.py:6 imports the loss from `torchmultimodal`, which is not installed.  Following SURVEY.md Appendix C this
module registers stub `torchmultimodal...contrastive_loss_with_temperature` modules whose
`ContrastiveLossWithTemperature` is oracle.mca_oracle's restatement of the reference's own adapted copy
(utils/contrastive_loss_with_temperature.py:40-108,178-195), and neutralises the debug `torch.save` inside
Attention.forward (model.py:94) while reference forwards run.  Nothing is copied from the reference; it is imported
in place and only used to (a) validate oracle/mca_oracle.py and (b) generate tests/golden/ fixtures
(oracle/make_golden.py).  On the GPU box /root/reference does not exist and `available()` is False.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MCA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _install_stub_modules():
    from oracle import mca_oracle

    names = [
        "torchmultimodal",
        "torchmultimodal.modules",
        "torchmultimodal.modules.losses",
        "torchmultimodal.modules.losses.contrastive_loss_with_temperature",
        "torchmultimodal.utils",
        "torchmultimodal.utils.distributed",
    ]
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules[names[3]].ContrastiveLossWithTemperature = mca_oracle.ContrastiveLossWithTemperature


_ref_model_module = None


def load_reference():
    """Import the reference's model.py in place and return the module (cached)."""
    global _ref_model_module
    if _ref_model_module is not None:
        return _ref_model_module
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stub_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference has top-level modules called `model`, `encoders`, `utils`; import them under their own names
    # but make sure our package-local names never shadow them
    import importlib

    _ref_model_module = importlib.import_module("model")
    assert os.path.dirname(os.path.abspath(_ref_model_module.__file__)) == os.path.abspath(REFERENCE_ROOT)
    return _ref_model_module


@contextlib.contextmanager
def no_debug_save():
    """model.py:94 calls torch.save(...) inside every attention forward; make it a no-op (declared in DESIGN.md)."""
    real = torch.save
    torch.save = lambda *a, **k: None
    try:
        yield
    finally:
        torch.save = real


def build_reference_model(model_kwargs: dict, state_dict=None):
    """Construct reference MCA(**model_kwargs) (stdout noise suppressed) and optionally load a state_dict."""
    ref = load_reference()
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        m = (ref.EAO if model_kwargs.get("eao") else ref.MCA)(**model_kwargs)  # train_accel_gpu.py:47-52
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m


def reference_forward(model, batch, no_loss=False):
    with no_debug_save(), contextlib.redirect_stdout(open(os.devnull, "w")):
        return model(batch, no_loss=no_loss)


_ref_metrics_module = None


def load_reference_metrics():
    """Import the reference's utils/metrics.py in place (cached).  It subclasses torchmetrics.Metric, which is not
    installed: a stub `Metric` providing add_state (list states) and `dim_zero_cat` = torch.cat is registered so the
    module imports; lalign / lunif / get_rank_metrics themselves are plain torch."""
    global _ref_metrics_module
    if _ref_metrics_module is not None:
        return _ref_metrics_module
    path = os.path.join(REFERENCE_ROOT, "utils", "metrics.py")
    if not os.path.isfile(path):
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "torchmetrics" not in sys.modules:
        tm, tmu, tmd = (types.ModuleType(n) for n in ("torchmetrics", "torchmetrics.utilities", "torchmetrics.utilities.data"))

        class Metric:
            def __init__(self, **kwargs):
                self._defaults = {}

            def add_state(self, name, default, dist_reduce_fx=None):
                self._defaults[name] = default
                setattr(self, name, list(default) if isinstance(default, list) else default)

            def reset(self):
                for k, v in self._defaults.items():
                    setattr(self, k, list(v) if isinstance(v, list) else v)

        tm.Metric = Metric
        tmd.dim_zero_cat = lambda x: torch.cat(list(x), 0) if isinstance(x, (list, tuple)) else x
        tm.utilities, tmu.data = tmu, tmd
        sys.modules.update({"torchmetrics": tm, "torchmetrics.utilities": tmu, "torchmetrics.utilities.data": tmd})
    import importlib.util

    spec = importlib.util.spec_from_file_location("_mca_reference_metrics", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ref_metrics_module = mod
    return mod
