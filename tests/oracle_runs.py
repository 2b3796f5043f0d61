"""The scenarios of tests/scenarios.py restated on the function-level CPU oracle
(oracle/oracle.py), for any number of y-slabs ("N-slab emulator").  Used to pin
the oracle's multi-step / multi-rank behaviour against tests/golden/*.npz and as
the expectation for the multi-GPU tests.  Test infrastructure.
"""
import numpy as np

from oracle import oracle as orc
import scenarios as sc


class Slabs:
    """nslabs oracle grids + per-slab particle arrays (reference Particles.__new__,
    particles.py:10-57: Nmax live-prefix arrays, nbmax = 0.1 Nmax, ntmax = 2 nbmax)"""

    def __init__(self, nslabs, Nmax, order=1, charge=1.0, mass=1.0, n0=1.0,
                 **gk):
        self.n = nslabs
        self.grids = [orc.Grid(rank=r, size=nslabs, **gk) for r in range(nslabs)]
        self.order, self.charge, self.mass, self.n0 = order, charge, mass, n0
        self.parts = [np.zeros(Nmax, orc.Particle) for _ in range(nslabs)]
        self.N = [0]*nslabs
        self.time = 0.0

    def initialize(self, x, y, vx, vy, vz):
        """Particles.initialize, particles.py:77-102"""
        for r, g in enumerate(self.grids):
            ind = np.logical_and(y >= g.y0 + g.edges[0]*g.dy,
                                 y < g.y0 + g.edges[1]*g.dy)
            n = int(ind.sum())
            p = self.parts[r]
            p["x"][:n] = (x[ind] - g.x0)/g.dx
            p["y"][:n] = (y[ind] - g.y0)/g.dy
            p["vx"][:n], p["vy"][:n], p["vz"][:n] = vx[ind], vy[ind], vz[ind]
            self.N[r] = n

    def live(self, r):
        return self.parts[r][:self.N[r]]

    def periodic_y(self):
        self.parts, self.N = orc.move(self.parts, self.N, self.grids)

    def push(self, E, B, dt, modified=False):
        """Particles.push / push_modified, particles.py:159-188, 233-257"""
        self.time += dt
        qtmh = self.charge/self.mass*dt/2
        g0 = self.grids[0]
        for r, g in enumerate(self.grids):
            orc.push(self.live(r), E[r], B[r], g, self.order, qtmh, dt,
                     modified, getattr(g0, "Omega", 0.0), getattr(g0, "S", 0.0))
        if g0.shear:
            for r, g in enumerate(self.grids):
                orc.shear_periodic_y(self.live(r), g, g0.S, self.time)
        self.periodic_y()
        for r, g in enumerate(self.grids):
            orc.periodic_x(self.live(r), g)

    def deposit(self, time=0.0, set_boundaries=False):
        """Sources.deposit, sources.py:27-50 (+ optional set_boundaries)"""
        g0 = self.grids[0]
        S = getattr(g0, "S", 0.0)
        src = [g.field(orc.Float4) for g in self.grids]
        for r, g in enumerate(self.grids):
            orc.deposit(self.live(r), src[r], g, self.order, S)
        orc.normalize(src, self.grids, self.N, self.charge, self.n0)
        if set_boundaries:
            orc.add_guards(src, self.grids, time)
            orc.copy_guards(src, self.grids, time)
        return src

    def gathered(self):
        return np.concatenate([self.live(r) for r in range(self.n)])


def fields(grids, dtype, fill=0.0):
    out = []
    for g in grids:
        f = g.field(dtype)
        for d in f.dtype.names:
            f[d] = fill
        out.append(f)
    return out


def active_cat(fs, grids):
    return np.concatenate([f[g.lby:g.uby, g.lbx:g.ubx] for f, g in zip(fs, grids)])


def sorted_particles(p):
    a = np.ascontiguousarray(p).view(np.float64).reshape(-1, 5)
    return a[np.lexsort((a[:, 4], a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]


def ohm_all(src, B, E, grids, **kw):
    for r, g in enumerate(grids):
        orc.ohm(src[r], B[r], E[r], g, **kw)


def ionacoustic(nslabs=1, nx=16, ny=16, npc=8, nt=6, order=1, lb=1, seed=11):
    charge, mass, Te = 0.5, 1.0, 1.0
    N = nx*ny*npc
    s = Slabs(nslabs, int(1.5*N/nslabs) + 16, order=order, charge=charge,
              mass=mass, nx=nx, ny=ny, lbx=lb, lby=lb, Lx=1.0, Ly=1.0)
    g = s.grids
    dt = 0.5*g[0].dx
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 0.05, seed)
    vx = vx + 0.1*np.sin(2*np.pi*x)
    s.initialize(x, y, vx, vy, vz)
    E = fields(g, orc.Float3)
    B = fields(g, orc.Float3)

    def fields_update():
        src = s.deposit(set_boundaries=True)
        ohm_all(src, B, E, g, charge=charge, temperature=Te)
        orc.copy_guards(E, g)
        return src
    src = fields_update()
    for it in range(nt):
        s.push(E, B, dt)
        src = fields_update()
    return dict(slabs=s, sources=src, E=E, grids=g)


def sheared(nslabs=1, nx=32, ny=16, npc=4, nt=6, order=1, seed=13, Omega=1.0):
    S = -1.5
    N = nx*ny*npc
    s = Slabs(nslabs, int(1.5*N/nslabs) + 16, order=order, nx=nx, ny=ny, lbx=2,
              lby=2, S=S, Omega=Omega, Lx=2.0, Ly=1.0, x0=-1.0, y0=-0.5)
    g = s.grids
    dt = 0.2*g[0].dx
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 0.3, seed, Lx=2.0, Ly=1.0,
                                     x0=-1.0, y0=-0.5)
    vx = vx - S*y
    s.initialize(x, y, vx, vy, vz)
    E = fields(g, orc.Float3)
    B = fields(g, orc.Float3)
    t = 0.0
    for it in range(nt):
        s.push(E, B, dt, modified=True)
        t += dt
        src = s.deposit(time=t, set_boundaries=True)
    return dict(slabs=s, sources=src, grids=g)


def gyro_fields(nslabs=1, nx=16, ny=32, npc=4, nt=8, order=1, lb=2, seed=12):
    N = nx*ny*npc
    gk = dict(nx=nx, ny=ny, lbx=lb, lby=lb, Lx=2.0, Ly=1.0, x0=-1.0, y0=0.25)
    s = Slabs(nslabs, int(1.5*N/nslabs) + 16, order=order, charge=1.0, mass=2.0,
              **gk)
    g = s.grids
    dt = 0.4*g[0].dy
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 0.2, seed, Lx=2.0, Ly=1.0,
                                     x0=-1.0, y0=0.25)
    s.initialize(x, y, vx, vy, vz)

    def smooth(amp, kind):
        fs = fields(g, orc.Float3)
        for f, gr in zip(fs, g):
            xg, yg = np.meshgrid(gr.x, gr.y)
            kx, ky, ph = 2*np.pi/gr.Lx, 2*np.pi/gr.Ly, 0.3*seed
            a = (slice(gr.lby, gr.uby), slice(gr.lbx, gr.ubx))
            if kind == "E":
                f["x"][a] = amp*np.sin(kx*xg + ph)*np.cos(ky*yg)
                f["y"][a] = amp*np.cos(kx*xg)*np.sin(ky*yg + ph)
                f["z"][a] = 0.5*amp*np.cos(kx*xg + ky*yg)
            else:
                f["x"][a] = 0.2*amp*np.sin(ky*yg + ph)
                f["y"][a] = 0.2*amp*np.sin(kx*xg + ph)
                f["z"][a] = amp*(1.0 + 0.1*np.cos(kx*xg)*np.cos(ky*yg))
        orc.copy_guards(fs, g)
        return fs
    E = smooth(0.3, "E")
    B = smooth(2.0, "B")
    for it in range(nt):
        s.push(E, B, dt)
    src = s.deposit(set_boundaries=True)
    return dict(slabs=s, sources=src, grids=g)
