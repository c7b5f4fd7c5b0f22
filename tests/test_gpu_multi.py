"""Multi-GPU parity: N ranks (one per GPU) against the oracle's N emulated ranks.  Skipped with < 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("coll", ["nccl", "p2p"])
@pytest.mark.parametrize("dep", [0, 3])
def test_two_or_more_gpus_match_oracle_ranks(dep, coll):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + dep + (10 if coll == "p2p" else 0)),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), str(dep), coll]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
