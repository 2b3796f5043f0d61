"""Where does the per-step host/communication overhead go at N GPUs?  Wall-clock per
phase with a device sync between phases (torchrun, one rank per GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import skeletor_b200 as sk
    comm = sk.comm.init_world()
    nx = ny = 2048
    ppc = 256
    m = sk.Manifold(nx, ny, comm, Lx=1.0, Ly=1.0)
    n = nx*m.nyp*ppc
    ions = sk.Particles(m, int(1.05*n) + 4096, nbmax=max(n//100, 1 << 16))
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + comm.rank)
    d = ions._data
    d[0, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*nx
    d[1, :n] = m.noff + torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*m.nyp
    d[2:5, :n] = torch.randn((3, n), generator=gen, device="cuda", dtype=torch.float64)
    ions.N = n
    dt = 0.1*m.dx
    E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3); B.fill((0., 0., 1.)); B.copy_guards()
    src = sk.Sources(m)
    phases = {}

    def tick(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        phases.setdefault(name, []).append(time.perf_counter() - t0)
        return r

    for it in range(12):
        comm.barrier()
        tick("push_kernel", lambda: ions._push_kernel(E, B, dt, False, count=True))
        tick("move", lambda: ions.move())
        tick("sort", lambda: ions.sort(precounted=False))
        tick("deposit_kernel+zero", lambda: (src.t.zero_(), sk._lib.call(
            "skb_deposit", ions._c, ions.N, src.ptr, m.c, 1, 0.0, ions._tiling_c(),
            torch.cuda.current_stream().cuda_stream)))
        src.boundaries_set = False
        tick("normalize", lambda: src.normalize(ions))
        tick("add_guards", lambda: src.add_guards())
        tick("copy_guards", lambda: src.copy_guards())
    if comm.rank == 0:
        for k, v in phases.items():
            print("%-22s %.3f ms (median of %d)" % (k, 1e3*np.median(v[2:]), len(v) - 2))
    comm.barrier()


if __name__ == "__main__":
    main()
