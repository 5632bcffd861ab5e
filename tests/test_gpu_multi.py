"""GPU, >= 2 devices: the NCCL-sharded search equals the single-GPU search (tools/dist_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_search_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tools", "dist_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DIST_CHECK OK %d" % world in out.stdout


def test_single_process_driver_equals_single_gpu():
    """theta_mgpu_optimize (one process, host thread + stream per GPU, ncclAllReduce(min)) == do_optimization_single, on
    every device count the box offers -- including 1, so the driver path is exercised on a one-GPU box too."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mgpu_check.py")], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "MGPU_CHECK OK 1" in out.stdout
    import torch
    if torch.cuda.device_count() >= 2:
        assert "MGPU_CHECK OK 2" in out.stdout
