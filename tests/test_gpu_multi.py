"""GPU (-m gpu), needs >= 2 GPUs on the box (skipped otherwise): the data-parallel step — pooled-block all-gather, gradient
reduce-scatter, sharded clip + AdamW, parameter push — against `oracle.mca_forward_ranks` / `oracle.clip_adamw_step` on
every rank, for both exchange forms (peer-memory kernels and NCCL).  The checks themselves live in
scripts/gpu_dp_check.py (it runs under torchrun, one process per GPU; its docstring lists them); logs of the N = 2 and
N = 8 runs are under profiles/."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, mode, p2p, port):
    env = dict(os.environ, MCA_P2P="1" if p2p else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "gpu_dp_check.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DP CHECK OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode,p2p", [("tiny", True), ("tiny", False), ("tiny_z", True)])
def test_data_parallel_step_against_oracle_2gpu(mode, p2p):
    _run(2, mode, p2p, 29531)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs 8 GPUs")
def test_data_parallel_step_against_oracle_8gpu():
    _run(8, "tiny", True, 29532)
