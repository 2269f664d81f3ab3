"""Multi-GPU MPPI on real devices: spawns tests/mgpu_worker.py under torchrun, one rank per GPU.
Needs >= 2 CUDA devices; the single-GPU box the driver uses for `-m gpu` skips it (the host-side logic of the
exchange is covered on CPU by tests/test_host_cpu.py::test_two_rank_exchange_over_gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_mppi_matches_oracle_and_single_gpu(gpu_required):
    from mpc_rs_b200 import _abi as A
    n = A.lib().mpcb_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on one box")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "mgpu ok: transport=nccl precision=f32" in r.stdout
