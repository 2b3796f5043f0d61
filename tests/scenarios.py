"""Small end-to-end scenarios written against skeletor's PUBLIC API.

The same function runs on the reference package (oracle/ref.py: package(), build
container only — used by oracle/make_golden.py to produce tests/golden/*.npz) and
on skeletor_b200 (the GPU tests), so the parity tests read like the reference's
own tests (tests/test_ionacoustic.py:160-178, tests/test_sheared_burgers.py,
tests/test_deposit.py, tests/test_fastwave.py).

`sk` is a namespace with: Manifold, ShearingManifold, Particles, Sources, Field,
Ohm, Faraday, State, Float3, comm, and the two time-stepper classes
(HorowitzStepper, PredictorCorrectorStepper).
"""
import numpy as np


def host(a):
    """structured/plain ndarray copy of a reference or skeletor_b200 array"""
    return np.array(np.asarray(a))


def maxwellian(nx, ny, npc, vth, seed, Lx=1.0, Ly=1.0, x0=0.0, y0=0.0):
    """positions in PHYSICAL units, as Particles.initialize expects"""
    rng = np.random.default_rng(seed)
    n = nx*ny*npc
    x = x0 + rng.uniform(0, Lx, n)
    y = y0 + rng.uniform(0, Ly, n)
    vx, vy, vz = rng.normal(0, vth, (3, n))
    return x, y, vx, vy, vz


def smooth_field(sk, manifold, amp, kind, seed=0):
    """A Float3 field with smooth sinusoidal components, guards set."""
    f = sk.Field(manifold, dtype=sk.Float3)
    f.fill((0.0, 0.0, 0.0))
    xg, yg = np.meshgrid(manifold.x, manifold.y)
    kx = 2*np.pi/manifold.Lx
    ky = 2*np.pi/manifold.Ly
    ph = 0.3*seed
    if kind == "E":
        f['x'].active = amp*np.sin(kx*xg + ph)*np.cos(ky*yg)
        f['y'].active = amp*np.cos(kx*xg)*np.sin(ky*yg + ph)
        f['z'].active = 0.5*amp*np.cos(kx*xg + ky*yg)
    else:
        f['x'].active = 0.2*amp*np.sin(ky*yg + ph)
        f['y'].active = 0.2*amp*np.sin(kx*xg + ph)
        f['z'].active = amp*(1.0 + 0.1*np.cos(kx*xg)*np.cos(ky*yg))
    f.copy_guards()
    return f


def ionacoustic(sk, nx=16, ny=16, npc=8, nt=6, order=1, lb=1, seed=11):
    """push -> deposit -> add_guards -> copy_guards -> Ohm -> copy_guards,
    the loop body of tests/test_ionacoustic.py:160-178."""
    charge, mass, Te = 0.5, 1.0, 1.0
    m = sk.Manifold(nx, ny, sk.comm, lbx=lb, lby=lb, Lx=1.0, Ly=1.0)
    dt = 0.5*m.dx
    N = nx*ny*npc
    ions = sk.Particles(m, int(1.5*N/sk.comm.size) + 16, charge=charge,
                        mass=mass, order=order)
    x, y, vx, vy, vz = maxwellian(nx, ny, npc, 0.05, seed)
    vx = vx + 0.1*np.sin(2*np.pi*x)
    ions.initialize(x, y, vx, vy, vz)
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, 0.0))
    B.copy_guards()
    sources = sk.Sources(m)
    ohm = sk.Ohm(m, temperature=Te, charge=charge)
    sources.deposit(ions)
    sources.add_guards()
    sources.copy_guards()
    ohm(sources, B, E)
    E.copy_guards()
    for it in range(nt):
        ions.push(E, B, dt)
        sources.deposit(ions)
        sources.add_guards()
        sources.copy_guards()
        ohm(sources, B, E)
        E.copy_guards()
    return dict(particles=host(ions[:ions.N]), N=np.int64(ions.N),
                sources=host(sources), E=host(E))


def gyro_fields(sk, nx=16, ny=32, npc=4, nt=8, order=1, lb=2, seed=12):
    """push with non-trivial E and B (Boris rotation), then deposit+guards."""
    m = sk.Manifold(nx, ny, sk.comm, lbx=lb, lby=lb, Lx=2.0, Ly=1.0,
                    x0=-1.0, y0=0.25)
    dt = 0.4*m.dy
    N = nx*ny*npc
    ions = sk.Particles(m, int(1.5*N/sk.comm.size) + 16, charge=1.0, mass=2.0,
                        order=order)
    x, y, vx, vy, vz = maxwellian(nx, ny, npc, 0.2, seed, Lx=2.0, Ly=1.0,
                                  x0=-1.0, y0=0.25)
    ions.initialize(x, y, vx, vy, vz)
    E = smooth_field(sk, m, 0.3, "E", seed)
    B = smooth_field(sk, m, 2.0, "B", seed)
    sources = sk.Sources(m)
    for it in range(nt):
        ions.push(E, B, dt)
    sources.deposit(ions, set_boundaries=True)
    return dict(particles=host(ions[:ions.N]), N=np.int64(ions.N),
                sources=host(sources))


def sheared(sk, nx=32, ny=16, npc=4, nt=6, order=1, seed=13, Omega=1.0):
    """push_modified + shear-periodic particles + deposit with S + sheared
    add_guards/copy_guards (tests/test_sheared_burgers.py:280-300,
    tests/test_sheared_disturbance.py)."""
    S = -1.5
    m = sk.ShearingManifold(nx, ny, sk.comm, lbx=2, lby=2, S=S, Omega=Omega,
                            Lx=2.0, Ly=1.0, x0=-1.0, y0=-0.5)
    dt = 0.2*m.dx
    N = nx*ny*npc
    ions = sk.Particles(m, int(1.5*N/sk.comm.size) + 16, charge=1.0, mass=1.0,
                        order=order)
    x, y, vx, vy, vz = maxwellian(nx, ny, npc, 0.3, seed, Lx=2.0, Ly=1.0,
                                  x0=-1.0, y0=-0.5)
    vx = vx - S*y            # background shear flow u = -S y
    ions.initialize(x, y, vx, vy, vz)
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, 0.0))
    sources = sk.Sources(m)
    t = 0.0
    for it in range(nt):
        ions.push_modified(E, B, dt)
        t += dt
        sources.deposit(ions)
        sources.time = t
        sources.add_guards()
        sources.copy_guards()
    return dict(particles=host(ions[:ions.N]), N=np.int64(ions.N),
                sources=host(sources), time=np.float64(ions.time))


def guards_only(sk, shear, seed=14, nx=16, ny=8, lbx=1, lby=2, time=0.37):
    """add_guards / copy_guards on a random Float4 field and copy_guards on a
    scalar (tests/test_deposit.py:66-89, tests/test_extended_grid.py,
    tests/test_copy_guards_with_shear.py)."""
    if shear:
        m = sk.ShearingManifold(nx, ny, sk.comm, lbx=lbx, lby=lby, S=-1.5,
                                Omega=0.0, Lx=2.0, Ly=1.0)
    else:
        m = sk.Manifold(nx, ny, sk.comm, lbx=lbx, lby=lby)
    rng = np.random.default_rng(seed)
    src = sk.Sources(m)
    src.time = time
    for d in ('t', 'x', 'y', 'z'):
        src[d][...] = rng.uniform(-1, 1, (m.myp, m.mx))
    src.add_guards()
    added = host(src)
    src.copy_guards()
    copied = host(src)
    f = sk.Field(m, dtype=np.float64)
    f.time = time
    f[...] = rng.uniform(-1, 1, (m.myp, m.mx))
    f.copy_guards()
    return dict(added=added, copied=copied, scalar=host(f))


def ohm_only(sk, seed=15, nx=16, ny=16, lb=1, eta=0.05):
    m = sk.Manifold(nx, ny, sk.comm, lbx=lb, lby=lb, Lx=1.0, Ly=2.0)
    rng = np.random.default_rng(seed)
    src = sk.Sources(m)
    src['t'][...] = rng.uniform(0.5, 1.5, (m.myp, m.mx))
    for d in ('x', 'y', 'z'):
        src[d][...] = rng.uniform(-1, 1, (m.myp, m.mx))
    src.copy_guards()
    B = sk.Field(m, dtype=sk.Float3)
    for d in ('x', 'y', 'z'):
        B[d][...] = rng.uniform(-1, 1, (m.myp, m.mx))
    B.copy_guards()
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    ohm = sk.Ohm(m, temperature=0.7, charge=1.3, eta=eta)
    ohm(src, B, E, set_boundaries=True)
    far = sk.Faraday(m)
    far(E, B, 0.01, set_boundaries=True)
    return dict(E=host(E), B=host(B))


def poisson_only(sk, nx=32, ny=64, seed=18):
    """E = grad del^-2 rho (tests/test_poisson.py): smooth multi-mode charge density"""
    m = sk.Manifold(nx, ny, sk.comm, lbx=1, lby=1, Lx=1.0, Ly=2.0, ax=0.0, ay=0.0)
    xg, yg = np.meshgrid(m.x, m.y)
    rho = sk.Field(m, dtype=np.float64)
    rho.fill(0.0)
    rho.active = (np.sin(2*np.pi*xg) * np.cos(2*np.pi*yg/2.0)
                  + 0.3*np.cos(2*np.pi*3*xg + 0.4) + 0.2*np.sin(2*np.pi*2*yg/2.0))
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    # (sk.Poisson(m)(rho, E) calls exactly this and drops the return value, poisson.py:9-12)
    ttp, we = m.grad_inv_del(rho, E)
    E.copy_guards()
    return dict(E=host(E), we=np.float64(we))


def quiet_lattice(nx, ny, sq, Lx=1.0, Ly=1.0):
    """sq x sq particles per cell on a regular sub-lattice (quiet start)"""
    ax = (np.arange(nx*sq) + 0.5)/(nx*sq)*Lx
    ay = (np.arange(ny*sq) + 0.5)/(ny*sq)*Ly
    x, y = np.meshgrid(ax, ay)
    return x.ravel().copy(), y.ravel().copy()


def _stepper_setup(sk, nx, ny, sq, order, lb, seed):
    """fast magnetosonic wave set-up in the spirit of tests/test_fastwave.py"""
    A = 1e-3
    m = sk.Manifold(nx, ny, sk.comm, lbx=lb, lby=lb, Lx=1.0, Ly=1.0)
    N = nx*ny*sq*sq
    ions = sk.Particles(m, int(1.5*N/sk.comm.size) + 16, charge=1.0, mass=1.0,
                        order=order)
    x, y = quiet_lattice(nx, ny, sq)
    kx, ky = 2*np.pi, 2*np.pi
    ph = kx*x + ky*y
    rng = np.random.default_rng(seed)
    vx = -A*np.sin(ph) + 1e-4*rng.normal(size=N)
    vy = -A*np.sin(ph) + 1e-4*rng.normal(size=N)
    vz = 1e-4*rng.normal(size=N)
    ions.initialize(x, y, vx, vy, vz)
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, 1.0))
    xg, yg = np.meshgrid(m.x, m.y)
    B['x'].active = 0.05*np.sin(2*np.pi*yg)
    B['y'].active = 0.05*np.sin(2*np.pi*xg)
    B.copy_guards()
    ohm = sk.Ohm(m, temperature=0.05, charge=1.0)
    state = sk.State(ions, B)
    return m, ions, state, ohm


def predictor_corrector(sk, nx=16, ny=8, sq=3, nt=3, order=2, lb=2, seed=16):
    """tests/test_fastwave.py: prepare + iterate (2 particle sweeps per step)."""
    m, ions, state, ohm = _stepper_setup(sk, nx, ny, sq, order, lb, seed)
    e = sk.PredictorCorrectorStepper(state, ohm, m)
    dt = 0.1*m.dx
    e.prepare(dt)
    for it in range(nt):
        e.iterate(dt)
    return dict(particles=host(ions[:ions.N]), N=np.int64(ions.N),
                E=host(e.E), B=host(e.B), sources=host(e.sources),
                t=np.float64(e.t))


def horowitz(sk, nx=16, ny=8, sq=3, nt=3, order=1, lb=2, seed=17):
    """tests/test_circular.py: Horowitz iterate."""
    m, ions, state, ohm = _stepper_setup(sk, nx, ny, sq, order, lb, seed)
    e = sk.HorowitzStepper(state, ohm, m)
    dt = 0.1*m.dx
    e.prepare(dt)
    for it in range(nt):
        e.iterate(dt)
    return dict(particles=host(ions[:ions.N]), N=np.int64(ions.N),
                E=host(e.E), B=host(e.B), sources=host(e.sources),
                t=np.float64(e.t))


SCENARIOS = {
    "ionacoustic_cic": lambda sk: ionacoustic(sk, order=1, lb=1),
    "ionacoustic_tsc": lambda sk: ionacoustic(sk, order=2, lb=2),
    "gyro_cic": lambda sk: gyro_fields(sk, order=1),
    "gyro_tsc": lambda sk: gyro_fields(sk, order=2),
    "sheared_cic": lambda sk: sheared(sk, order=1),
    "sheared_tsc": lambda sk: sheared(sk, order=2, Omega=0.0),
    "guards_plain": lambda sk: guards_only(sk, shear=False),
    "guards_shear": lambda sk: guards_only(sk, shear=True),
    "ohm_faraday": lambda sk: ohm_only(sk),
    "poisson": lambda sk: poisson_only(sk),
    "predictor_corrector_tsc": lambda sk: predictor_corrector(sk),
    "horowitz_cic": lambda sk: horowitz(sk),
}
