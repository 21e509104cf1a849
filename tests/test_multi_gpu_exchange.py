"""On-hardware multi-rank correctness of the view-parallel gradient exchange (needs >= 2 GPUs; the CPU
suite covers the host logic with gloo in test_parallel_gloo.py). Spawns `torch.distributed.run` with one
rank per GPU over NCCL; the worker asserts sparse exchange == dense all-reduce == single-rank accumulation,
that the statistics travel with it and that the replicas stay bit-identical after Adam + MCMC noise."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sparse_exchange_equals_dense_equals_single_rank(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, {_gpus()} visible")
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "workers" / "exchange_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    tail = (p.stdout + p.stderr)[-4000:]
    assert p.returncode == 0, tail
    assert "total failures 0" in p.stdout, tail
