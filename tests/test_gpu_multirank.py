"""Multi-process row bands over NCCL (one process per GPU).  Needs >= 2 GPUs: run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_nccl_bands_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered on one GPU by test_gpu_bands.py through the same iteration structure)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join("tests", "_nccl_band_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-identical") == 2
