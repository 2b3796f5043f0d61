"""-m gpu: physics acceptance tests, the reference's own tests run on skeletor_b200
with the reference's tolerances:

  ion-acoustic wave   reference tests/test_ionacoustic.py:12-201 (BASELINE config 1)
  gyromotion          reference tests/test_gyromotion.py
  E x B drift         reference tests/test_EcrossBdrift_along_x.py
  shearing epicycle   reference tests/test_shearing_epicycle.py (its check sits inside
                      `if plot:`; here it is asserted)
  Landau damping      reference example/landau_ions.py (no assert there; the fitted
                      damping rate is compared with the kinetic dispersion relation)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ionacoustic():
    import skeletor_b200 as sk
    comm = sk.COMM_SELF
    nx, ny, npc = 32, 32, 256
    charge, mass, Te, A = 0.5, 1.0, 1.0, 0.001
    ikx = iky = 1
    cfl = 0.5
    manifold = sk.Manifold(nx, ny, comm, Lx=1.0, Ly=1.0)
    xg, yg = np.meshgrid(manifold.x, manifold.y)
    cs = np.sqrt(Te/mass)
    dt = cfl/cs*manifold.dx
    N = npc*nx*ny
    kx, ky = 2*np.pi*ikx/manifold.Lx, 2*np.pi*iky/manifold.Ly
    k = np.sqrt(kx*kx + ky*ky)
    omega = k*cs
    nt = int(2*np.pi/omega/dt)

    def rho_an(x, y, t):
        return charge*(1 + A*np.cos(kx*x + ky*y)*np.sin(omega*t))

    def u_an(x, y, t, kk):
        return -omega/k*A*np.sin(kx*x + ky*y)*np.cos(omega*t)*kk/k

    ions = sk.Particles(manifold, int(1.5*N), charge=charge, mass=mass)
    sk.InitialCondition(npc, quiet=True)(manifold, ions)
    x = ions['x']*manifold.dx
    y = ions['y']*manifold.dy
    ions['vx'] = u_an(x, y, dt/2, kx)
    ions['vy'] = u_an(x, y, dt/2, ky)
    assert ions.N == N
    E = sk.Field(manifold, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    E.copy_guards()
    B = sk.Field(manifold, dtype=sk.Float3)
    B.fill((0.0, 0.0, 0.0))
    B.copy_guards()
    sources = sk.Sources(manifold)
    ohm = sk.Ohm(manifold, temperature=Te, charge=charge)
    sources.deposit(ions)
    assert np.isclose(sources.rho.sum(), ions.N*charge/npc)
    sources.add_guards()
    sources.copy_guards()
    assert np.isclose(sources.rho.trim().sum(), N*charge/npc)
    ohm(sources, B, E)
    E.copy_guards()
    t, diff2 = 0.0, 0.0
    for it in range(nt):
        ions.push(E, B, dt)
        t += dt
        sources.deposit(ions)
        sources.add_guards()
        sources.copy_guards()
        ohm(sources, B, E)
        E.copy_guards()
        diff2 += ((rho_an(xg, yg, t) - sources.rho.trim())**2).mean()
    # reference tests/test_ionacoustic.py:201
    assert np.sqrt(diff2/nt) < 4e-5*charge


def test_gyromotion():
    import skeletor_b200 as sk
    dt, tend = 1e-3, 10
    nt = int(tend/dt)
    bz = 1.0
    og, phi, ampl = bz, 0.0, 8/32
    x0, y0 = 8/32, 33/32
    x_an = lambda t: -ampl*np.cos(og*t + phi) + x0
    y_an = lambda t: +ampl*np.sin(og*t + phi) + y0
    vx = og*ampl*np.sin(phi)*np.ones(1)
    vy = og*ampl*np.cos(phi)*np.ones(1)
    x = np.array([x_an(-dt/2)]) + vx*dt/2
    y = np.array([y_an(-dt/2)]) + vy*dt/2
    m = sk.Manifold(32, 64, sk.COMM_SELF, Lx=1.0, Ly=2.0)
    ions = sk.Particles(m, 1, charge=1, mass=1)
    ions.initialize(x, y, vx, vy, np.zeros(1))
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, bz))
    B.copy_guards()
    t = 0.0
    for it in range(nt):
        ions.push(E, B, dt)
        t += dt
        if it % 50 == 0 or it == nt - 1:
            assert ions.N == 1
            p = np.asarray(ions[:1])
            err = max(abs(p['x'][0]*m.dx - x_an(t)), abs(p['y'][0]*m.dy - y_an(t)))/ampl
            assert err < 5.0e-3          # reference tests/test_gyromotion.py:142


def test_ExB_drift_and_periodic_wrap():
    """uniform E_y, B_z: drift along x at E_y/B_z through the periodic x boundary and
    with migration through the y boundary (single rank: wraps onto itself)"""
    import skeletor_b200 as sk
    m = sk.Manifold(32, 64, sk.COMM_SELF, Lx=1.0, Ly=2.0)
    ey, bz, dt = 0.05, 1.0, 2e-3
    vd = ey/bz
    ions = sk.Particles(m, 1, charge=1, mass=1)
    # at rest in the drift frame: pure drift, no gyration
    x0, y0 = 0.9, 1.97
    ions.initialize(np.array([x0 + vd*dt/2]), np.array([y0]), np.array([vd]),
                    np.zeros(1), np.zeros(1))
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, ey, 0.0))
    E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, bz))
    B.copy_guards()
    nt = 4000
    for it in range(nt):
        ions.push(E, B, dt)
    p = np.asarray(ions[:1])
    x_exp = (x0 + vd*(nt*dt + dt/2)) % m.Lx
    assert ions.N == 1
    assert abs(p['x'][0]*m.dx - x_exp) < 1e-3
    assert abs(p['y'][0]*m.dy - y0) < 1e-3
    assert abs(p['vx'][0] - vd) < 1e-9


def test_shearing_epicycle():
    import skeletor_b200 as sk
    dt = 0.5e-3
    nt = int(2*np.pi/dt)
    Omega, S = 1.0, -1.5
    Sz = 2.0*Omega
    og = np.sqrt(Sz*(Sz + S))
    phi = np.pi/2
    nx, ny, Lx, Ly = 64, 32, 2.0, 1.0
    ampl = Lx/3
    x0, y0 = Lx/2, Ly/2
    y_an = lambda t: ampl*np.cos(og*t + phi) + y0
    x_an = lambda t: (Sz/og)*ampl*np.sin(og*t + phi) + x0 - S*t*y0
    vy_an = lambda t: -og*ampl*np.sin(og*t + phi)
    vx_an = lambda t: Sz*ampl*np.cos(og*t + phi) - S*y0
    vx, vy = np.array([vx_an(0.0)]), np.array([vy_an(0.0)])
    x = np.array([x_an(-dt/2)]) + vx*dt/2
    y = np.array([y_an(-dt/2)]) + vy*dt/2
    m = sk.ShearingManifold(nx, ny, sk.COMM_SELF, S=S, Omega=Omega, Lx=Lx, Ly=Ly)
    ions = sk.Particles(m, 1, charge=1, mass=1)
    ions.initialize(x, y, vx, vy, np.zeros(1))
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, 0.0))

    def wrapped(t):
        """analytic orbit folded back with the shearing-periodic boundary rules"""
        xx, yy, vxx = x_an(t), y_an(t), vx_an(t)
        while yy < 0:
            xx -= S*Ly*t; vxx -= S*Ly; yy += Ly
        while yy >= Ly:
            xx += S*Ly*t; vxx += S*Ly; yy -= Ly
        return xx % Lx, yy, vxx
    t = 0.0
    for it in range(nt):
        ions.push_modified(E, B, dt)
        t += dt
        if it % 100 == 0 or it == nt - 1:
            assert ions.N == 1
            p = np.asarray(ions[:1])
            xa, ya, vxa = wrapped(t)
            dx_ = abs(p['x'][0]*m.dx - xa)
            dx_ = min(dx_, Lx - dx_)
            err = max(dx_, abs(p['y'][0]*m.dy - ya))/ampl
            assert err < 2e-2            # reference tests/test_shearing_epicycle.py:218


def test_poisson_solves_gauss_law():
    """E = grad del^-2 rho: analytic check (reference tests/test_poisson.py)"""
    import skeletor_b200 as sk
    nx, ny = 32, 64
    m = sk.Manifold(nx, ny, sk.COMM_SELF, Lx=1.0, Ly=2.0)
    xg, yg = np.meshgrid(m.x, m.y)
    kx, ky = 2*np.pi*2/m.Lx, 2*np.pi*3/m.Ly
    rho = sk.Field(m, dtype=np.float64)
    rho.active = np.sin(kx*xg + ky*yg)
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    sk.Poisson(m)(rho, E)
    k2 = kx*kx + ky*ky
    assert np.abs(E['x'].active - (-kx/k2*np.cos(kx*xg + ky*yg))).max() < 5e-8
    assert np.abs(E['y'].active - (-ky/k2*np.cos(kx*xg + ky*yg))).max() < 5e-8


def test_landau_damping_of_ion_acoustic_wave():
    """reference example/landau_ions.py (BASELINE config 2's driver): 32 x 1 grid,
    2^16 particles per cell, Ti/Te = 1/5, noisy start.  The example only plots; here the
    decay rate fitted to the maxima of the density amplitude is compared with the
    kinetic dispersion relation Ti/Te + W(vph/vt) = 0 (example/landau_ions.py:233-242)."""
    import skeletor_b200 as sk
    from scipy.optimize import newton
    from scipy.signal import argrelextrema
    from scipy.special import wofz
    nx, ny, npc = 32, 1, 2**16
    charge = mass = Te = 1.0
    Ti, A = 1/5, 0.01
    cs = np.sqrt(Te/mass)
    N = npc*nx*ny
    m = sk.Manifold(nx, ny, sk.COMM_SELF)
    kx = 2*np.pi/m.Lx
    omega = kx*cs
    dt = 0.5*m.dx/cs
    nt = int(2*np.pi*3.0/omega/dt)
    rng = np.random.default_rng(2024)
    x = m.Lx*rng.uniform(size=N)
    y = m.Ly*rng.uniform(size=N)
    vx = -omega/kx*A*np.sin(kx*x) + np.sqrt(Ti/mass)*rng.normal(size=N)
    ions = sk.Particles(m, int(1.5*N), charge=charge, mass=mass)
    ions.initialize(x, y, vx, np.zeros(N), np.zeros(N))
    assert ions.N == N
    xg, yg = np.meshgrid(m.x, m.y)
    S, C = np.sin(kx*xg)/(nx*ny), np.cos(kx*xg)/(nx*ny)
    E = sk.Field(m, dtype=sk.Float3)
    E.fill((0.0, 0.0, 0.0))
    E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3)
    B.fill((0.0, 0.0, 0.0))
    B.copy_guards()
    sources = sk.Sources(m)
    ohm = sk.Ohm(m, temperature=Te, charge=charge)
    sources.deposit(ions, set_boundaries=True)
    ohm(sources, B, E)
    E.copy_guards()
    ampl, time, t = [], [], 0.0
    for it in range(nt):
        ions.push(E, B, dt)
        t += dt
        sources.deposit(ions, set_boundaries=True)
        ohm(sources, B, E)
        E.copy_guards()
        rho = sources.rho.active
        ampl.append(np.sqrt((S*rho).sum()**2 + (C*rho).sum()**2))
        time.append(t)
    ampl, time = np.array(ampl), np.array(time)
    peaks = argrelextrema(ampl, np.greater)[0]
    peaks = peaks[ampl[peaks] > 5e-4]          # above the particle noise floor
    assert len(peaks) >= 3
    gamma_fit = np.polyfit(time[peaks[:4]], np.log(ampl[peaks[:4]]), 1)[0]

    def W(z):
        return 1. + 1j*np.sqrt(0.5*np.pi)*z*wofz(np.sqrt(0.5)*z)
    vt = np.sqrt(Ti/Te)*cs
    vph = newton(lambda v: Ti/Te + W(v/vt), cs + 0j)
    gamma_t = kx*vph.imag
    assert gamma_t < 0 and gamma_fit < 0
    assert abs(gamma_fit - gamma_t) < 0.2*abs(gamma_t), (gamma_fit, gamma_t)


def test_density_perturbation_quiet_start():
    """reference tests/test_density_perturbation.py (quiet start, :79): the y-averaged
    deposited density of the displaced lattice matches 1 + ampl cos(kx x)"""
    import skeletor_b200 as sk
    nx, ny, npc, ampl = 128, 8, 256, 0.1
    Lx, Ly = 4, 1
    m = sk.Manifold(nx, ny, sk.COMM_SELF, Lx=Lx, Ly=Ly, x0=-Lx/2, y0=-Ly/2)
    N = npc*nx*ny
    ions = sk.Particles(m, int(1.5*N), charge=1.0, mass=1.0)
    sk.DensityPertubation(npc, 1, 0, ampl, quiet=True, global_init=True)(m, ions)
    src = sk.Sources(m)
    src.deposit(ions)
    assert np.isclose(src.rho.sum(), ions.N*1.0/npc)
    src.add_guards()
    src.copy_guards()
    assert np.isclose(src.rho.trim().sum(), N*1.0/npc)
    rho = src.rho.trim().mean(axis=0)
    rho_exact = 1 + ampl*np.cos(2*np.pi/Lx*m.x)
    assert np.sqrt(np.mean((rho - rho_exact)**2)) < 0.005*ampl


def test_io_snapshots(tmp_path, monkeypatch):
    """reference skeletor/io.py: run directory, info.p, fields.NNNN.npz, log"""
    import pickle
    import skeletor_b200 as sk
    monkeypatch.chdir(tmp_path)
    m = sk.Manifold(16, 8, sk.COMM_SELF)
    src = sk.Sources(m)
    E = sk.Field(m, dtype=sk.Float3)
    rng = np.random.default_rng(0)
    src['t'][...] = rng.uniform(0, 1, (m.myp, m.mx))
    E['x'][...] = rng.uniform(0, 1, (m.myp, m.mx))
    nx, dt = 16, 0.5
    io = sk.IO(str(tmp_path/"run"), locals(), __file__, tag="t", comm=sk.COMM_SELF)
    io.set_outputrate(dt)
    io.output_fields(src, E, m, 0.0)
    io.output_fields(src, E, m, 0.5)
    io.log(1, 0.5, dt)
    io.finished()
    d = np.load(tmp_path/"run"/"fields.0001.npz")
    assert np.array_equal(d["rho"], src.rho.trim()) and np.array_equal(d["Ex"], E['x'].trim())
    assert d["t"] == 0.5 and np.array_equal(d["x"], m.x)
    info = pickle.load(open(tmp_path/"run"/"info.p", "rb"))
    assert info["nx"] == 16 and info["MPI"] == 1 and "seconds" in info
    assert (tmp_path/"run"/"skeletor.log").exists()


def test_on_device_initial_condition():
    import skeletor_b200 as sk
    m = sk.Manifold(64, 32, sk.COMM_SELF)
    npc = 32
    ions = sk.Particles(m, int(1.5*64*32*npc))
    sk.InitialCondition(npc, vt=0.3, on_device=True, seed=7)(m, ions)
    assert ions.N == 64*32*npc
    p = np.asarray(ions[:ions.N])
    assert (p['x'] >= 0).all() and (p['x'] < 64).all() and (p['y'] >= 0).all() and (p['y'] < 32).all()
    assert abs(p['vx'].std() - 0.3) < 0.01 and abs(p['x'].mean() - 32) < 0.5
    src = sk.Sources(m)
    src.deposit(ions, set_boundaries=True)
    assert np.isclose(src.rho.trim().sum(), 64*32)
