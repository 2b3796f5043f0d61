"""-m gpu: N-GPU parity (needs >= 2 GPUs on the box; skipped otherwise).  Spawns
tests/mgpu_check.py under torchrun, one rank per GPU over NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("layout", ["dense", "gapped"])
@pytest.mark.parametrize("nproc", [2, 4])
def test_n_ranks_equal_one_rank(nproc, layout):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + 11*nproc), os.path.join(HERE, "mgpu_check.py")]
    env = dict(os.environ, MGPU_GAPPED="1" if layout == "gapped" else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
