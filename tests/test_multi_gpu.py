"""Multi-GPU parity on a box with >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`):
launches tests/multi_gpu_worker.py under torchrun.  Skipped on single-GPU boxes; the sharding logic itself is
covered on CPU by tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_levels_sharded_over_two_gpus_p2p_and_nccl(build_lib):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(HERE, "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "exchange p2p" in out.stdout and "exchange nccl" in out.stdout
