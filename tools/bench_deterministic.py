"""cost of Particles.deterministic at config 5 (one GPU): canonical-order pass and the
atomic-free deposit next to the default kernels"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import skeletor_b200 as sk
from skeletor_b200 import _lib

nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ppc = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = sk.Manifold(nx, ny, sk.COMM_SELF)
n = nx*ny*ppc
ions = sk.Particles(m, int(1.05*n) + 4096, nbmax=max(n//100, 1 << 16))
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
d = ions._data
d[0, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*nx
d[1, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*ny
d[2:5, :n] = torch.randn((3, n), generator=gen, device="cuda", dtype=torch.float64)
ions.N = n
src = sk.Sources(m)
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ions.sort()
t_sort = timeit(ions.sort)
t_dep = timeit(lambda: _lib.call("skb_deposit", ions._c, ions.N, src.ptr, m.c, 1, 0.0,
                                 ions._tiling_c(), st()))
cs = ions._cellsums()
t_det = timeit(lambda: _lib.call("skb_deposit_deterministic", ions._c, ions.N, src.ptr, m.c,
                                 1, 0.0, ions._tiling_c(), cs.data_ptr(), st()))
t_can = timeit(lambda: _lib.call("skb_canonical_cells", ions._c, ions._soa(ions._alt),
                                 ions._cell_counts.data_ptr(), m.c, 4, 4, st()))
print("n=%d ppc=%d: full sort %.2f ms | deposit (atomics) %.2f ms | deposit (deterministic) "
      "%.2f ms | canonical order pass %.2f ms" % (n, ppc, t_sort, t_dep, t_det, t_can))
