"""where the non-kernel time of a gapped push() goes: CUDA events + host clock"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import skeletor_b200 as sk
from skeletor_b200 import _lib

nx = ny = int(os.environ.get("NXY", 2048)); ppc = 256
m = sk.Manifold(nx, ny, sk.COMM_SELF)
n = nx*ny*ppc
ions = sk.Particles(m, int(1.36*n) + 4096, nbmax=max(n//100, 1 << 16))
ions.gapped = True
sk.InitialCondition(ppc, vt=1.0, on_device=True, seed=1)(m, ions)
E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
B = sk.Field(m, dtype=sk.Float3); B.fill((0., 0., 1.)); B.copy_guards()
src = sk.Sources(m)
dt = 0.1*m.dx
for _ in range(4):
    ions.push(E, B, dt); src.deposit(ions)
ev = lambda: torch.cuda.Event(enable_timing=True)
marks = []
orig_call = _lib.call
def traced(name, *a):
    e0 = ev(); e0.record(); h0 = time.perf_counter()
    r = orig_call(name, *a)
    e1 = ev(); e1.record(); h1 = time.perf_counter()
    marks.append((name, e0, e1, h0, h1))
    return r
import skeletor_b200.particles as P, skeletor_b200.sources as S
for it in range(3):
    marks.clear()
    torch.cuda.synchronize()
    P._lib.call = traced; S._lib.call = traced
    t0 = ev(); t0.record(); h0 = time.perf_counter()
    ions.push(E, B, dt)
    t1 = ev(); t1.record(); h1 = time.perf_counter()
    src.deposit(ions)
    t2 = ev(); t2.record()
    torch.cuda.synchronize(); h2 = time.perf_counter()
    P._lib.call = orig_call; S._lib.call = orig_call
    print("iter %d push() gpu %.3f ms host %.3f ms | deposit() gpu %.3f ms | host total %.3f" % (
        it, t0.elapsed_time(t1), (h1-h0)*1e3, t1.elapsed_time(t2), (h2-h0)*1e3))
    for name, e0, e1, a, b in marks:
        print("   %-22s start +%.3f ms  gpu %.3f ms   host call %.3f ms (at +%.3f)" % (
            name, t0.elapsed_time(e0), e0.elapsed_time(e1), (b-a)*1e3, (a-h0)*1e3))
