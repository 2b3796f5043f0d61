"""CPU, world_size 2 over gloo: the communicator shim (skeletor_b200/comm.py) that
replaces mpi4py sendrecv (reference field.py:52-58) and the MPI calls inside
cppmove2 (pplib2.c:741-753, 873) — ring exchange of halo rows, variable-size
particle exchange with a count handshake, scalar reductions."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, size, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(size))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from skeletor_b200 import comm as C
        from oracle import oracle as orc
        world = C.TorchComm()
        assert (world.rank, world.size) == (rank, size)
        # scalar reductions (sources.py:59, pplib2.c:873, horowitz.py:119)
        assert world.allreduce(rank + 1, op=C.SUM) == size*(size + 1)//2
        assert world.allreduce(float(rank), op=C.MAX) == float(size - 1)
        assert world.allgather({"r": rank}) == [{"r": r} for r in range(size)]
        assert world.bcast("x" if rank == 0 else None) == "x"
        # halo rows: `up` goes to rank+1, `down` to rank-1 (field.py:15-16)
        up = torch.full((2, 5), float(10*rank + 1), dtype=torch.float64)
        dn = torch.full((2, 5), float(10*rank + 2), dtype=torch.float64)
        fb, fa = world.ring_exchange(up, dn)
        below, above = (rank - 1) % size, (rank + 1) % size
        assert torch.all(fb == 10*below + 1) and torch.all(fa == 10*above + 2)
        # device-tensor collectives of the steppers / Poisson solve (host tensors here)
        acc = torch.tensor([float(rank + 1)], dtype=torch.float64)
        world.allreduce_tensor_(acc)
        assert float(acc) == size*(size + 1)/2
        slab = torch.full((3, 4), float(rank), dtype=torch.float64)
        full = world.allgather_tensor(slab)
        assert full.shape == (3*size, 4)
        assert all(torch.all(full[3*r:3*r + 3] == r) for r in range(size))
        # object sendrecv (mpi4py lowercase API)
        got = world.sendrecv(np.arange(4.0) + rank, dest=above, source=below)
        assert np.array_equal(got, np.arange(4.0) + below)

        # migration protocol on host tensors vs the oracle's N-slab cppmove2
        nx = ny = 16
        grids = [orc.Grid(nx, ny, rank=r, size=size) for r in range(size)]
        rng = np.random.default_rng(5)
        parts, Ns = [], []
        for g in grids:
            n = 400
            p = np.zeros(600, orc.Particle)
            p["x"][:n] = rng.uniform(0, nx, n)
            p["y"][:n] = rng.uniform(g.edges[0] - 1.5, g.edges[1] + 1.5, n)
            parts.append(p)
            Ns.append(n)
        exp, expN = orc.move([p.copy() for p in parts], list(Ns), grids)
        g = grids[rank]
        mine = parts[rank][:Ns[rank]]
        a = np.ascontiguousarray(mine).view(np.float64).reshape(-1, 5)
        dn_m = a[:, 1] < g.edges[0]
        up_m = ~dn_m & (a[:, 1] >= g.edges[1])
        sd, su = a[dn_m].copy(), a[up_m].copy()
        if rank == 0:
            sd[:, 1] += ny
        if rank == size - 1:
            su[:, 1] -= ny
        nb, na = world.exchange_counts(len(su), len(sd))
        rb = torch.empty((max(nb, 1), 5), dtype=torch.float64)
        ra = torch.empty((max(na, 1), 5), dtype=torch.float64)
        pad = lambda x: torch.as_tensor(np.concatenate([x, np.zeros((1, 5))])[:max(len(x), 1)])
        fb, fa = world.ring_exchange(pad(su), pad(sd), rb, ra)
        new = np.concatenate([a[~dn_m & ~up_m], fb.numpy()[:nb], fa.numpy()[:na]])
        assert len(new) == expN[rank]
        e = np.ascontiguousarray(exp[rank][:expN[rank]]).view(np.float64).reshape(-1, 5)
        srt = lambda z: z[np.lexsort(z.T[::-1])]
        assert np.array_equal(srt(new), srt(e))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_torchcomm_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_selfcomm_is_identity():
    from skeletor_b200 import comm as C
    c = C.COMM_SELF
    assert (c.rank, c.size) == (0, 1)
    assert c.allreduce(7) == 7 and c.allgather(3) == [3] and c.bcast(2) == 2
    t = torch.ones(3)
    assert c.ring_exchange(t, 2*t)[0] is t
    assert c.exchange_counts(4, 5) == (4, 5)
