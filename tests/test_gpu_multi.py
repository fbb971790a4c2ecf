"""Multi-GPU parity (runs where at least two GPUs are visible, skipped otherwise): one rank per GPU under torchrun,
NCCL all-reduce of the sufficient statistics; every rank must end with the identical proposal, equal to the single-GPU
iteration over the same global sample (tools/mgpu_parity.py)."""
import os
import subprocess
import sys

import pytest

from pmclib_b200.lib import load_library

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg,port", [("c3", 29561), ("c5", 29562)])
def test_two_ranks_equal_one_gpu(cfg, port):
    if load_library().pmcgpu_device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MGPU_CONFIG=cfg, MGPU_N="200003" if cfg == "c5" else "2000003")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_parity.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MGPU PARITY OK" in r.stdout
