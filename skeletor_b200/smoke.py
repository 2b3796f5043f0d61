"""One small invocation of the hot path on cuda:0, checked against the CPU oracle:
push (CIC) + boundary epilogue + migration + ordering + deposit + guards on a
32x32 grid with 16 particles per cell, once on the dense layout (tile sort) and once
on the gapped layout (per-cell slot ranges)."""
import numpy as np


def run():
    import torch
    import skeletor_b200 as sk
    from oracle import oracle as orc

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    torch.cuda.set_device(0)
    nx = ny = 32
    npc = 16
    rng = np.random.default_rng(0)
    n = nx*ny*npc
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    vx, vy, vz = rng.normal(0, 0.3, (3, n))

    def on_gpu(gapped):
        m = sk.Manifold(nx, ny, sk.COMM_SELF, lbx=2, lby=2)
        ions = sk.Particles(m, int((4.0 if gapped else 1.5)*n), charge=1.0, mass=1.0,
                            order=1)
        ions.gapped = gapped
        ions.initialize(x, y, vx, vy, vz)
        E = sk.Field(m, dtype=sk.Float3)
        B = sk.Field(m, dtype=sk.Float3)
        xg, yg = np.meshgrid(m.x, m.y)
        E['x'].active = 0.1*np.sin(2*np.pi*xg)
        E['y'].active = 0.1*np.cos(2*np.pi*yg)
        B['z'].active = 1.0 + 0.1*np.sin(2*np.pi*(xg + yg))
        E.copy_guards()
        B.copy_guards()
        src = sk.Sources(m)
        for it in range(3):
            ions.push(E, B, 0.5*m.dx)
        assert ions._rep == ("gapped" if gapped else "dense")
        src.deposit(ions, set_boundaries=True)
        torch.cuda.synchronize()
        return m, ions, src, E, B

    m, ions, src, E, B = on_gpu(False)
    dt = 0.5*m.dx

    # the same on the CPU oracle
    g = orc.Grid(nx, ny, lbx=2, lby=2)
    p = np.zeros(int(1.5*n), orc.Particle)
    p["x"][:n], p["y"][:n] = x/g.dx, y/g.dy
    p["vx"][:n], p["vy"][:n], p["vz"][:n] = vx, vy, vz
    Eo, Bo = np.asarray(E).copy(), np.asarray(B).copy()
    parts, N = [p], [n]
    for it in range(3):
        orc.push(parts[0][:N[0]], Eo, Bo, g, 1, 1.0*dt/2, dt)
        parts, N = orc.move(parts, N, [g])
        orc.periodic_x(parts[0][:N[0]], g)
    so = g.field(orc.Float4)
    orc.deposit(parts[0][:N[0]], so, g, 1)
    orc.normalize([so], [g], N, 1.0, 1.0)
    orc.add_guards([so], [g])
    orc.copy_guards([so], [g])

    def rows(a):
        a = np.ascontiguousarray(a).view(np.float64).reshape(-1, 5)
        return a[np.lexsort(a.T[::-1])]
    for layout in ("dense", "gapped"):
        if layout == "gapped":
            m, ions, src, E, B = on_gpu(True)
        assert ions.N == N[0]
        got, exp = rows(np.asarray(ions[:ions.N])), rows(parts[0][:N[0]])
        assert np.array_equal(got, exp), "particles differ from the oracle (%s)" % layout
        a = np.asarray(src).view(np.float64)
        b = so.view(np.float64)
        err = np.abs(a - b).max()/np.abs(b).max()
        assert err < 1e-12, "sources differ from the oracle (%s): %g" % (layout, err)
        print("smoke OK (%s layout): %d particles bit-exact, sources rel err %.2e" % (
            layout, ions.N, err))
