"""step time of the gapped layout (Particles.gapped) next to the default tile-sort step"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import skeletor_b200 as sk

nx = ny = int(os.environ.get("NXY", 2048))
ppc = int(os.environ.get("PPC", 256))
steps = int(os.environ.get("STEPS", 12))
for gapped in ((True,) if os.environ.get("ONLY_GAPPED") else (True, False)):
    m = sk.Manifold(nx, ny, sk.COMM_SELF)
    n = nx*ny*ppc
    nmax = int((1.36 if gapped else 1.05)*n) + 4096
    ions = sk.Particles(m, nmax, nbmax=max(n//100, 1 << 16))
    ions.gapped = gapped
    ions.mover_fraction = 0.18
    sk.InitialCondition(ppc, vt=1.0, on_device=True, seed=1)(m, ions)
    E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3); B.fill((0., 0., 1.)); B.copy_guards()
    src = sk.Sources(m)
    dt = 0.1*m.dx
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc = {"push()": [], "deposit()": [], "step": []}
    reps = []
    for it in range(steps):
        t = [ev() for _ in range(4)]
        torch.cuda.synchronize()
        t[0].record(); ions.push(E, B, dt)
        t[1].record(); src.deposit(ions)
        t[2].record(); src.add_guards(); src.copy_guards()
        t[3].record()
        torch.cuda.synchronize()
        acc["push()"].append(t[0].elapsed_time(t[1]))
        acc["deposit()"].append(t[1].elapsed_time(t[2]))
        acc["step"].append(t[0].elapsed_time(t[3]))
        reps.append(ions._rep[0])
    print("gapped" if gapped else "dense ", "reps", "".join(reps), "N", ions.N)
    for k, v in acc.items():
        print("  %-10s median %.3f ms  min %.3f  all %s" % (
            k, np.median(v[3:]), min(v[3:]), " ".join("%.1f" % x for x in v)))
    rho = src["t"] if hasattr(src, "__getitem__") else None
    tot = float(src.t[m.lby:m.uby, m.lbx:m.ubx, 0].sum().item())
    print("  total charge %.10g (cells %d)" % (tot, nx*ny))
    del ions, src, E, B
    torch.cuda.empty_cache()
