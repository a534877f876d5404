"""Sharded (one process per GPU, NCCL) extraction == single-GPU extraction, bit for bit.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p2p,mc", [("1", "1"), ("1", "0"), ("0", "1")])
def test_sharded_extraction_matches_single_gpu(lib_built, p2p, mc):
    """p2p: peer-memory exchange kernels (1) or NCCL collectives (0); mc: sample / count exchanges over several blocks (1) or one."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PR_P2P=p2p, PR_P2P_MC=mc))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU check ok" in r.stdout
    assert ("peer-memory kernels" if p2p == "1" else "NCCL") in r.stdout
