"""ctypes binding of libmca_b200.so (the C ABI declared in include/mca_b200.h).

There is deliberately no fallback: if the shared library is missing the import of any kernel-backed op raises,
so a GPU box can never silently run an eager PyTorch path (the product path is the CUDA extension or nothing).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCA_LIB selects another build of the SAME library (e.g. the -DMCA_TRACE debug build); there is still no fallback
LIB_PATH = os.environ.get("MCA_LIB") or os.path.join(_HERE, "csrc", "libmca_b200.so")

MCA_OK = 0
ERR_NAMES = {1: "MCA_ERR_SHAPE", 2: "MCA_ERR_ALIGN", 3: "MCA_ERR_CUDA", 4: "MCA_ERR_NONFINITE", 5: "MCA_ERR_ARG"}

EPI_BF16, EPI_F32, EPI_RESID, EPI_GEGLU, EPI_GEGLU_BWD, EPI_F32_ACC = range(6)

_lib = None


class MCAKernelError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MCAKernelError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/eager fallback for the mca_paper_b200 hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def check(rc: int, what: str) -> None:
    if rc == MCA_OK:
        return
    name = ERR_NAMES.get(rc, str(rc))
    if rc == 1:
        raise AssertionError(f"{what}: {name}")
    raise MCAKernelError(f"{what}: {name}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as a c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
