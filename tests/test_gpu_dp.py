"""Multi-GPU data-parallel correctness as a test (it was only a script).  Needs >= 2 GPUs: skipped on a one-GPU
box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu -s` (log kept under profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_step_two_gpus(capsys):
    """scripts/check_dp.py under torchrun on 2 GPUs: identical weights on every rank after bucketed / overlapped steps,
    overlap == plain hook, and the all-reduced gradient == the mean of the per-shard ORACLE gradients (fp32, 1e-4)."""
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_dp.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    with capsys.disabled():
        print("\n" + r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ranks identical = True" in r.stdout and "DP gradient vs oracle" in r.stdout
