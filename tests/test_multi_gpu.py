"""Real-GPU multi-rank parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).
The gloo test (tests/test_sharding_gloo.py) covers the host logic on CPU; this one runs the CUDA path on every rank under
torchrun + NCCL and compares the concatenated output with the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_ranks_equal_the_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("multi-GPU parity ok") == 2, out.stdout
