"""GPU (>= 2 devices): the multi-GPU exchange step - NCCL broadcast and the fused peer-store scatter - must hand
every rank exactly the spectrum the ingest rank computed (bit-identical PCM vs a local recomputation)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_rank_spectrum_exchange(gpu_required):
    if gpu_required < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29621", str(ROOT / "tests" / "mgpu_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_EXCHANGE_OK" in out.stdout
