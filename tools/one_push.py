"""Tuning aid: time ONE gapped push launch (fresh particles every time) for the library
named by SKELETOR_B200_LIB.  Prints ms (CUDA events) of 3 independent first launches."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import skeletor_b200 as sk

nx = int(os.environ.get("NXY", "2048"))
ppc = 256
m = sk.Manifold(nx, nx, sk.COMM_SELF, Lx=1.0, Ly=1.0)
n = nx*nx*ppc
nmax = int(1.36*n) + 4096
nmax += nmax & 1
ions = sk.Particles(m, nmax, order=1, nbmax=max(n//100, 1 << 16))
ions.gapped = True
ions.fuse_deposit = False
E = sk.Field(m, dtype=sk.Float3)
B = sk.Field(m, dtype=sk.Float3)
xg, yg = np.meshgrid(m.x, m.y)
E['x'].active = 0.01*np.sin(2*np.pi*xg/m.Lx)
E['y'].active = 0.01*np.cos(2*np.pi*yg/m.Ly)
B['z'].active = 1.0
E.copy_guards(); B.copy_guards()
dt = 0.1*m.dx
out = []
for rep in range(3):
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + rep)
    ions._rep = "dense"
    d = ions._data
    d[0, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*nx
    d[1, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*nx
    d[2:5, :n] = torch.randn((3, n), generator=gen, device="cuda", dtype=torch.float64)
    ions.N = n
    ions._sorted = False
    assert ions._to_gapped()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cnt = ions._gap_kernel(E, B, dt, False, False)
    e1.record()
    torch.cuda.synchronize()
    out.append(round(e0.elapsed_time(e1), 3))
print(os.path.basename(os.environ.get("SKELETOR_B200_LIB", "default")), out, flush=True)

if os.environ.get("REPEAT"):
    # steady state: consecutive pushes of the same particles, kernel time of each
    times = []
    orig = ions._gap_kernel

    def timed(*a, **k):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        r = orig(*a, **k)
        a1.record()
        times.append((a0, a1))
        return r
    ions._gap_kernel = timed
    for it in range(int(os.environ["REPEAT"])):
        ions.push(E, B, dt)
        if os.environ.get("PAUSE"):
            torch.cuda.synchronize()
            import time
            time.sleep(float(os.environ["PAUSE"]))
    torch.cuda.synchronize()
    print("steady", [round(a.elapsed_time(b), 3) for a, b in times], "rep", ions._rep,
          "fail", getattr(ions, "_gap_fail", None), flush=True)
