"""where does the step time go?  CUDA-event time of the public calls vs the kernels"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import skeletor_b200 as sk

nx = ny = 2048
ppc = 256
m = sk.Manifold(nx, ny, sk.COMM_SELF)
n = nx*ny*ppc
ions = sk.Particles(m, int(1.05*n) + 4096, nbmax=max(n//100, 1 << 16))
sk.InitialCondition(ppc, vt=1.0, on_device=True, seed=1)(m, ions)
E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
B = sk.Field(m, dtype=sk.Float3); B.fill((0., 0., 1.)); B.copy_guards()
src = sk.Sources(m)
dt = 0.1*m.dx
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {"push()": [], "deposit()": [], "add_guards()": [], "copy_guards()": [], "step": []}
for it in range(8):
    t = [ev() for _ in range(5)]
    torch.cuda.synchronize()
    t[0].record(); ions.push(E, B, dt)
    t[1].record(); src.deposit(ions)
    t[2].record(); src.add_guards()
    t[3].record(); src.copy_guards()
    t[4].record()
    torch.cuda.synchronize()
    for k, i in zip(acc, range(4)):
        acc[k].append(t[i].elapsed_time(t[i + 1]))
    acc["step"].append(t[0].elapsed_time(t[4]))
for k, v in acc.items():
    print("%-14s median %.3f ms  min %.3f" % (k, np.median(v[2:]), min(v[2:])))
