"""Tuning probe: does the (issue-bound) gapped push overlap with the (HBM-bound) deposit when
the two kernels run on different streams over independent particle sets?  Prints the serial
and the concurrent time of push(set 1) + deposit(set 2)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import skeletor_b200 as sk
from skeletor_b200 import _lib
from skeletor_b200.field import _stream

nx, ny, ppc = 2048, 1024, 256
m = sk.Manifold(nx, ny, sk.COMM_SELF, Lx=1.0, Ly=0.5)
n = nx*ny*ppc
nmax = int(1.36*n) + 4096
nmax += nmax & 1
E = sk.Field(m, dtype=sk.Float3)
B = sk.Field(m, dtype=sk.Float3)
xg, yg = np.meshgrid(m.x, m.y)
E['x'].active = 0.01*np.sin(2*np.pi*xg/m.Lx)
E['y'].active = 0.01*np.cos(2*np.pi*yg/m.Ly)
B['z'].active = 1.0
E.copy_guards(); B.copy_guards()
dt = 0.1*m.dx
sets = []
for k in range(2):
    ions = sk.Particles(m, nmax, order=1, nbmax=max(n//100, 1 << 16))
    ions.gapped = True
    ions.fuse_deposit = False
    gen = torch.Generator(device="cuda"); gen.manual_seed(77 + k)
    d = ions._data
    d[0, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*nx
    d[1, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*ny
    d[2:5, :n] = torch.randn((3, n), generator=gen, device="cuda", dtype=torch.float64)
    ions.N = n
    ions._sorted = False
    assert ions._to_gapped()
    sets.append(ions)
src = sk.Sources(m)
a, b = sets


def push(ions):
    ions.push(E, B, dt)


def dep(ions):
    src.t.zero_()
    _lib.call("skb_deposit", ions._c, ions._N, src.ptr, src.grid.c, ions.order, 0.0,
              ions._tiling_c(), _stream())


def ev():
    return torch.cuda.Event(enable_timing=True)


for _ in range(2):
    push(a); dep(b)
torch.cuda.synchronize()
for mode in ("serial", "concurrent", "concurrent_dep_hi", "concurrent_push_hi"):
    res = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        if mode == "serial":
            e0.record()
            push(a); dep(b)
            e1.record()
        else:
            pa, pb = {"concurrent": (0, 0), "concurrent_dep_hi": (0, -1),
                      "concurrent_push_hi": (-1, 0)}[mode]
            s1 = torch.cuda.Stream(priority=pa)
            s2 = torch.cuda.Stream(priority=pb)
            cur = torch.cuda.current_stream()
            e0.record()
            s1.wait_stream(cur); s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                push(a)
            with torch.cuda.stream(s2):
                dep(b)
            cur.wait_stream(s1); cur.wait_stream(s2)
            e1.record()
        torch.cuda.synchronize()
        res.append(round(e0.elapsed_time(e1), 3))
    print(mode, res, flush=True)
# the parts alone
for name, f, arg in (("push", push, a), ("deposit", dep, b)):
    res = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(); f(arg); e1.record()
        torch.cuda.synchronize()
        res.append(round(e0.elapsed_time(e1), 3))
    print(name, res, flush=True)
