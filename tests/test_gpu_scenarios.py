"""-m gpu: the scenarios of tests/scenarios.py run through skeletor_b200's public
API (the reference's own API) and compared with the golden fixtures generated from
the unmodified reference (oracle/make_golden.py).

Tolerance: particle coordinates and fields <= 1e-12 relative (north_star): the
deposit's summation order differs from the reference's serial loop, which feeds
back into the particles through E after the first step; integer results (particle
counts) are exact."""
import os
import types

import numpy as np
import pytest

import scenarios as sc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


GAPPED_PUSHES = [0]


def namespace(gapped=False):
    import skeletor_b200 as sk
    from skeletor_b200.time_steppers.horowitz import TimeStepper as Horowitz
    from skeletor_b200.time_steppers.predictor_corrector import TimeStepper as PC

    class GappedParticles(sk.Particles):
        """same scenarios on the gapped particle layout (room for the slot ranges)"""

        def __init__(self, manifold, Nmax, **kw):
            super().__init__(manifold, 4*int(Nmax) + 8192, **kw)   # (empty cells own 4 slots each)
            self.gapped = True

        def _gap_kernel(self, *args, **kw):
            GAPPED_PUSHES[0] += 1
            return super()._gap_kernel(*args, **kw)

        def _push_and_deposit_gapped(self, *args, **kw):
            GAPPED_PUSHES[0] += 1
            return super()._push_and_deposit_gapped(*args, **kw)

    return types.SimpleNamespace(
        Manifold=sk.Manifold, ShearingManifold=sk.ShearingManifold,
        Particles=GappedParticles if gapped else sk.Particles, Sources=sk.Sources, Field=sk.Field, Ohm=sk.Ohm,
        Faraday=sk.Faraday, State=sk.State, Float3=sk.Float3, comm=sk.COMM_SELF,
        Poisson=sk.Poisson,
        HorowitzStepper=Horowitz, PredictorCorrectorStepper=PC)


def rows(p):
    a = np.ascontiguousarray(p).view(np.float64).reshape(-1, 5)
    return a[np.lexsort((a[:, 4], a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]


def check(name, got, exp, rtol=1e-12):
    g = np.ascontiguousarray(got).view(np.float64).ravel()
    e = np.ascontiguousarray(exp).view(np.float64).ravel()
    assert g.shape == e.shape, name
    scale = np.abs(e).max()
    err = np.abs(g - e).max()
    assert err <= rtol*max(scale, 1e-300), "%s: rel err %.3e" % (name, err/scale)


@pytest.mark.parametrize("layout", ["dense", "gapped"])
@pytest.mark.parametrize("name", sorted(sc.SCENARIOS))
def test_scenario_matches_reference(name, layout, capsys):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    if layout == "gapped" and "particles" not in gold.files:
        pytest.skip("no particles in this scenario")
    GAPPED_PUSHES[0] = 0
    res = sc.SCENARIOS[name](namespace(layout == "gapped"))
    assert set(res) == set(gold.files)
    if layout == "gapped":
        assert GAPPED_PUSHES[0] > 0, "the gapped push never ran"
    # north star: <= 1e-12 relative.  The steppers iterate Ohm / Faraday (log(rho), a
    # division by rho, a convergence loop) and amplify the last-bit differences of the
    # deposit a little: measured against the single-rank golden fixtures E 6.6e-13
    # (horowitz_cic) and 2.5e-13 (predictor_corrector_tsc), B 4.4e-16, particles 2e-16
    # (profiles/mgpu_check_r02_*.log), i.e. inside the bound with a factor of two to
    # spare for other summation orders
    rtol = 2e-12 if ("horowitz" in name or "predictor" in name) else 1e-12
    if name == "poisson":
        # the reference truncates the spectrum to float32 (operators.pyx:6-8, 97-101;
        # SURVEY.md Q3): parity is defined at that level
        rtol = 2e-6
    for key in gold.files:
        if key == "N":
            assert int(res[key]) == int(gold[key])
        elif key == "particles":
            check(name + ".particles", rows(res[key]), rows(gold[key]), rtol)
        else:
            check(name + "." + key, res[key], gold[key], rtol)


def test_deposit_conservation_and_guards():
    """reference tests/test_deposit.py:63,79-80 and tests/test_extended_grid.py:48-54"""
    import skeletor_b200 as sk
    nx, ny, npc = 64, 32, 32
    m = sk.Manifold(nx, ny, sk.COMM_SELF, lbx=1, lby=2)
    n = nx*ny*npc
    ions = sk.Particles(m, int(1.25*n), charge=0.7)
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 1.0, 3)
    ions.initialize(x, y, vx, vy, vz)
    src = sk.Sources(m)
    src.deposit(ions)
    assert np.isclose(src.rho.sum(), ions.N*ions.charge/npc)
    src.add_guards()
    assert np.isclose(src.rho.trim().sum(), n*ions.charge/npc)
    assert src.rho.sum() == src.rho.trim().sum()        # guards are zero
    src.copy_guards()
    assert np.isclose(src.rho.trim().sum(), n*ions.charge/npc)


def test_user_writes_keep_working():
    """tests poke particle storage directly (tests/test_ionacoustic.py:87-92); a
    write invalidates the tile ordering but never the results"""
    import skeletor_b200 as sk
    m = sk.Manifold(32, 32, sk.COMM_SELF)
    ions = sk.Particles(m, 5000)
    x, y, vx, vy, vz = sc.maxwellian(32, 32, 4, 0.0, 4)
    ions.initialize(x, y, vx, vy, vz)
    src = sk.Sources(m)
    src.deposit(ions, set_boundaries=True)
    assert ions._sorted
    xp = ions['x']*m.dx
    ions['vx'] = 0.25*np.sin(2*np.pi*xp)
    assert not ions._sorted
    assert np.allclose(np.asarray(ions['vx'])[:ions.N],
                       0.25*np.sin(2*np.pi*np.asarray(ions['x'])[:ions.N]*m.dx))
    ions['x'][:10] = np.arange(10) + 0.5
    assert np.array_equal(np.asarray(ions[:10]['x']), np.arange(10) + 0.5)
    src.deposit(ions, set_boundaries=True)
    assert np.isclose(src.rho.trim().sum(), 32*32)


def test_reference_style_script_runs_through_compat_imports(tmp_path):
    """compat/: `from skeletor import ...` + `from mpi4py.MPI import COMM_WORLD` (the
    imports of every reference test) resolve to the B200 path; run in a subprocess so the
    names do not clash with the oracle's serial mpi4py stand-in"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ref_style.py"
    script.write_text('''
from skeletor import Float3, Field, Particles, Sources, Ohm, InitialCondition
from skeletor.manifolds.second_order import Manifold
import numpy as np
from mpi4py import MPI
from mpi4py.MPI import COMM_WORLD as comm

nx, ny, npc = 32, 32, 16
manifold = Manifold(nx, ny, comm, Lx=1.0, Ly=1.0)
N = npc*nx*ny
ions = Particles(manifold, int(1.5*N/comm.size), charge=0.5, mass=1.0)
InitialCondition(npc, quiet=True)(manifold, ions)
x = ions['x']*manifold.dx
ions['vx'] = 1e-3*np.sin(2*np.pi*x)
assert comm.allreduce(ions.N, op=MPI.SUM) == N
E = Field(manifold, dtype=Float3); E.fill((0.0, 0.0, 0.0)); E.copy_guards()
B = Field(manifold, dtype=Float3); B.fill((0.0, 0.0, 0.0)); B.copy_guards()
sources = Sources(manifold)
ohm = Ohm(manifold, temperature=1.0, charge=0.5)
for it in range(5):
    ions.push(E, B, 0.5*manifold.dx)
    sources.deposit(ions)
    sources.add_guards()
    sources.copy_guards()
    ohm(sources, B, E)
    E.copy_guards()
assert np.isclose(comm.allreduce(sources.rho.trim().sum(), op=MPI.SUM), N*0.5/npc)
print("compat OK")
''')
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "compat"), root]))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True,
                       env=env, timeout=300)
    assert r.returncode == 0 and "compat OK" in r.stdout, r.stderr[-2000:]


def test_piecewise_guard_api_equals_fused():
    """the reference's piecewise guard methods (copy_guards_x/_y, add_guards_x/_y, the
    *_old test helpers; field.py:73-151, sources.py:91-192) give the fused results —
    the bitwise new-vs-old check of reference tests/test_deposit.py:77,89"""
    import skeletor_b200 as sk
    m = sk.Manifold(32, 16, sk.COMM_SELF, lbx=1, lby=2)
    rng = np.random.default_rng(8)
    vals = {d: rng.uniform(-1, 1, (m.myp, m.mx)) for d in 'txyz'}

    def fresh():
        s = sk.Sources(m)
        for d in 'txyz':
            s[d][...] = vals[d]
        return s
    a, b, c = fresh(), fresh(), fresh()
    a.add_guards()
    b.add_guards_old()
    c.add_guards_x()
    c.add_guards_y()
    c[:m.lby, :] = 0.0
    c[m.uby:, :] = 0.0
    c[:, m.ubx:] = 0.0
    c[:, :m.lbx] = 0.0
    assert np.all(np.asarray(a) == np.asarray(b)) and np.all(np.asarray(a) == np.asarray(c))
    a.copy_guards()
    b.copy_guards_old()
    c.copy_guards_y()
    c.copy_guards_x()
    assert np.all(np.asarray(a) == np.asarray(b)) and np.all(np.asarray(a) == np.asarray(c))
    a += b
    assert np.allclose(np.asarray(a['t']), 2*np.asarray(b['t']))


def test_particle_proxies_stay_live_across_pushes():
    """the reference returns NumPy views: `vx = ions['vx']` stays valid for the whole run
    (a proxy bound to one buffer went stale when push / sort swapped buffers)"""
    import skeletor_b200 as sk
    m = sk.Manifold(32, 32, sk.COMM_SELF)
    n = 32*32*8
    ions = sk.Particles(m, 4*n)
    x, y, vx, vy, vz = sc.maxwellian(32, 32, 8, 0.5, 5)
    ions.initialize(x, y, vx, vy, vz)
    E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3); B.copy_guards()
    px, pvx = ions['x'], ions['vx']
    for it in range(3):
        ions.push(E, B, 0.3*m.dx)
    now = np.asarray(ions[:ions.N])
    assert np.array_equal(np.asarray(px)[:ions.N], now['x'])
    pvx[:ions.N] = 0.25                      # a write through the old proxy lands
    assert np.all(np.asarray(ions[:ions.N])['vx'] == 0.25)


def test_kick_is_push_without_drift():
    """Particles.kick (north star): velocities bit-identical to the oracle's push,
    positions, time and count untouched"""
    import skeletor_b200 as sk
    from oracle import oracle as orc
    nx, ny, npc = 32, 16, 8
    n = nx*ny*npc
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 0.3, 9)
    for order in (1, 2):
        m = sk.Manifold(nx, ny, sk.COMM_SELF, lbx=2, lby=2)
        g = orc.Grid(nx, ny, lbx=2, lby=2)
        E = sc.smooth_field(sk, m, 0.2, "E")
        B = sc.smooth_field(sk, m, 1.0, "B")
        dt = 0.2*m.dx
        ions = sk.Particles(m, 4*n, charge=0.7, mass=1.3, order=order)
        ions.initialize(x, y, vx, vy, vz)
        p = np.zeros(n, orc.Particle)
        p["x"], p["y"] = x/g.dx, y/g.dy
        p["vx"], p["vy"], p["vz"] = vx, vy, vz
        exp = p.copy()
        orc.push(exp, np.asarray(E).copy(), np.asarray(B).copy(), g, order, 0.7/1.3*dt/2, dt)
        exp["x"], exp["y"] = p["x"], p["y"]          # kick: no drift
        ions.kick(E, B, dt)
        assert ions.time == 0.0 and ions.N == n
        assert np.array_equal(rows(np.asarray(ions[:n])), rows(exp)), order


def test_on_device_quiet_start_and_density_perturbation():
    """initial_condition.py:20-62, 65-143 generated in device memory: the quiet start's
    sub-lattice is bit-identical to the host path, the perturbed positions agree to
    rounding of sin / cos (1e-13 cells)"""
    import skeletor_b200 as sk
    nx, ny, npc = 32, 16, 16
    m = sk.Manifold(nx, ny, sk.COMM_SELF)
    n = nx*ny*npc
    host = sk.Particles(m, 2*n)
    dev = sk.Particles(m, 2*n)
    np.random.seed(3)
    sk.InitialCondition(npc, quiet=True, vt=0.0)(m, host)
    sk.InitialCondition(npc, quiet=True, vt=0.0, on_device=True)(m, dev)
    assert host.N == dev.N == n
    a, b = np.asarray(host[:n]), np.asarray(dev[:n])
    assert np.array_equal(a['x'], b['x']) and np.array_equal(a['y'], b['y'])
    host2 = sk.Particles(m, 2*n)
    dev2 = sk.Particles(m, 2*n)
    sk.DensityPertubation(npc, 1, 0, 0.2, quiet=True, vt=0.0)(m, host2)
    sk.DensityPertubation(npc, 1, 0, 0.2, quiet=True, vt=0.0, on_device=True)(m, dev2)
    a, b = np.asarray(host2[:n]), np.asarray(dev2[:n])
    assert np.abs(a['x'] - b['x']).max() < 1e-12 and np.array_equal(a['y'], b['y'])
    src = sk.Sources(m)
    src.deposit(dev2, set_boundaries=True)
    rho = src.rho.trim()
    xg, yg = np.meshgrid(m.x, m.y)
    assert np.abs(rho - (1 + 0.2*np.cos(2*np.pi*xg/m.Lx))).max() < 0.02


@pytest.mark.parametrize("order", [1, 2])
def test_drift_on_gapped_layout_equals_oracle(order):
    """Particles.drift (particles.py:259-265) without leaving the gapped layout
    (skb_drift_gapped): bit-exact against the oracle's drift + cppmove2 + periodic_x"""
    import skeletor_b200 as sk
    from oracle import oracle as orc
    nx, ny, npc = 64, 32, 16
    n = nx*ny*npc
    m = sk.Manifold(nx, ny, sk.COMM_SELF, lbx=2, lby=2)
    g = orc.Grid(nx, ny, lbx=2, lby=2)
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 0.8, 21)
    ions = sk.Particles(m, 4*n + 8192, order=order)
    ions.gapped = True
    ions.initialize(x, y, vx, vy, vz)
    p = np.zeros(2*n, orc.Particle)
    p["x"][:n], p["y"][:n] = x/g.dx, y/g.dy
    p["vx"][:n], p["vy"][:n], p["vz"][:n] = vx, vy, vz
    parts, N = [p], [n]
    dt = 0.4*m.dx
    for it in range(3):
        ions.drift(dt)
        assert ions._rep == "gapped"
        orc.drift(parts[0][:N[0]], g, dt)
        orc.periodic_x(parts[0][:N[0]], g)
        parts, N = orc.move(parts, N, [g])
        assert ions.N == N[0]
    assert np.array_equal(rows(np.asarray(ions[:ions.N])), rows(parts[0][:N[0]]))


def test_overlapped_migration_equals_immediate():
    """Particles.overlap_migration (default with more than one rank): push() returns with
    its migration still in flight and Sources.deposit overlaps it - same particles (bit
    for bit), same sources (summation order) as the immediate path"""
    import skeletor_b200 as sk
    nx, ny, npc = 64, 32, 16
    n = nx*ny*npc
    x, y, vx, vy, vz = sc.maxwellian(nx, ny, npc, 0.8, 31)
    out = {}
    for overlap in (False, True):
        m = sk.Manifold(nx, ny, sk.COMM_SELF, lbx=2, lby=2)
        E = sc.smooth_field(sk, m, 0.2, "E")
        B = sc.smooth_field(sk, m, 1.0, "B")
        ions = sk.Particles(m, 4*n + 8192)
        ions.gapped = True
        ions.overlap_migration = overlap
        ions.initialize(x, y, vx, vy, vz)
        src = sk.Sources(m)
        for it in range(4):
            ions.push(E, B, 0.4*m.dx)
            assert (ions._pending is not None) == overlap
            src.deposit(ions)
            assert ions._pending is None
            src.add_guards()
            src.copy_guards()
        ions.push(E, B, 0.4*m.dx)                 # finished by the read below
        out[overlap] = (rows(np.asarray(ions[:ions.N])), np.asarray(src).view(np.float64).copy(),
                        ions.N)
    assert out[True][2] == out[False][2] == n
    assert np.array_equal(out[True][0], out[False][0])
    check("sources", out[True][1], out[False][1], 1e-12)
