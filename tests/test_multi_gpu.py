"""N > 1 on real GPUs (skipped on a single-GPU box): launches tests/multi_gpu_worker.py under torchrun."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2])
def test_slq_and_moments_across_ranks(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "tests" / "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi_gpu_worker ok" in r.stdout
