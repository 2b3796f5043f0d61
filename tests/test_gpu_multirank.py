"""-m gpu: N-rank parity ("N ranks == 1 rank", reference tests/test_skeletor.py:142-150).
Spawns tests/mgpu_check.py under torchrun: one rank per GPU over NCCL (+ the NVLink
peer-memory exchange) when the box has enough GPUs; otherwise the N ranks SHARE cuda:0
and exchange their messages over gloo (host-staged), so that the slab decomposition,
the migration kernels (pack / classify / multi-hop forwarding / edge-rank y wrap) and
the halo kernels are still checked against the single-rank golden fixtures."""
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
    shared = torch.cuda.device_count() < nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + 11*nproc), os.path.join(HERE, "mgpu_check.py")]
    env = dict(os.environ, MGPU_GAPPED="1" if layout == "gapped" else "0")
    if shared:
        env["SKELETOR_B200_BACKEND"] = "gloo"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
