"""Pin the CPU oracle (oracle/skeletor_oracle.c + oracle/oracle.py) against the
UNMODIFIED reference compiled into oracle/_ref (oracle/build_ref.py).

Bit-exact comparisons: the oracle keeps the reference's operation order, both
are gcc -O2 without FMA.  Skipped when oracle/_ref is not built.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref
from refutil import bits, random_field, random_particles, ref_grid

pytestmark = pytest.mark.skipif(not ref.available(),
                                reason="oracle/_ref not built")

GRIDS = [
    dict(nx=32, ny=32, lbx=1, lby=1),
    dict(nx=16, ny=64, lbx=2, lby=2, Lx=2.0, Ly=1.0, x0=-0.5, y0=-0.25),
    dict(nx=64, ny=32, rank=1, size=4, lbx=2, lby=3, Lx=1.0, Ly=3.0),
]


@pytest.mark.parametrize("gk", GRIDS)
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("modified", [False, True])
def test_push_bitexact(gk, order, modified):
    if order == 2 and gk["lbx"] < 2:
        pytest.skip("TSC needs two guard layers")
    k = ref.kernels()
    g = orc.Grid(**gk)
    rng = np.random.default_rng(1)
    p = random_particles(g, 5000, rng)
    E = random_field(g, orc.Float3, rng)
    B = random_field(g, orc.Float3, rng)
    qtmh, dt = 0.37*0.05/2, 0.05
    Omega, S = 1.0, -1.5
    a, b = p.copy(), p.copy()
    orc.push(a, E, B, g, order, qtmh, dt, modified, Omega, S)
    name = ("modified_" if modified else "") + "boris_push_" + \
        ("cic" if order == 1 else "tsc")
    fn = getattr(k.particle_push, name)
    if modified:
        fn(b, E, B, qtmh, dt, ref_grid(g), Omega, S)
    else:
        fn(b, E, B, qtmh, dt, ref_grid(g))
    assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("gk", GRIDS)
def test_drift_and_boundaries_bitexact(gk):
    k = ref.kernels()
    g = orc.Grid(**gk)
    rg = ref_grid(g)
    rng = np.random.default_rng(2)
    p = random_particles(g, 4000, rng, vth=30.0)
    a, b = p.copy(), p.copy()
    orc.drift(a, g, 0.1)
    k.particle_push.drift(b, 0.1, rg)
    assert np.array_equal(bits(a), bits(b))
    # shear boost uses global 0 / ny
    orc.shear_periodic_y(a, g, -1.5, 0.7)
    k.particle_boundary.shear_periodic_y(b, rg, -1.5, 0.7)
    assert np.array_equal(bits(a), bits(b))
    orc.periodic_x(a, g)
    k.particle_boundary.periodic_x(b, rg)
    assert np.array_equal(bits(a), bits(b))
    assert (a["x"] >= 0).all() and (a["x"] < g.nx).all()
    for ntmax in (8000, 10):      # normal and overflowing hole list
        ia = np.zeros(ntmax + 1, np.int32)
        ib = np.zeros(ntmax + 1, np.int32)
        orc.calculate_ihole(a, ia, g)
        k.particle_boundary.calculate_ihole(b, ib, rg)
        assert np.array_equal(ia, ib)
    assert ib[0] < 0          # the small list overflowed: -count


@pytest.mark.parametrize("gk", GRIDS)
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("S", [0.0, -1.5])
def test_deposit_bitexact(gk, order, S):
    if order == 2 and gk["lbx"] < 2:
        pytest.skip("TSC needs two guard layers")
    k = ref.kernels()
    g = orc.Grid(**gk)
    rng = np.random.default_rng(3)
    p = random_particles(g, 6000, rng)
    a = g.field(orc.Float4)
    b = g.field(orc.Float4)
    orc.deposit(p, a, g, order, S)
    fn = k.deposit.deposit_cic if order == 1 else k.deposit.deposit_tsc
    fn(p, b, ref_grid(g), S)
    assert np.array_equal(bits(a), bits(b))
    assert np.isclose(a["t"].sum(), p.shape[0])


@pytest.mark.parametrize("gk", GRIDS)
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("update", [True, False])
def test_push_and_deposit_bitexact(gk, order, update):
    if order == 2 and gk["lbx"] < 2:
        pytest.skip("TSC needs two guard layers")
    k = ref.kernels()
    g = orc.Grid(**gk)
    rng = np.random.default_rng(4)
    p = random_particles(g, 5000, rng, vth=1.0)
    E = random_field(g, orc.Float3, rng, -0.1, 0.1)
    B = random_field(g, orc.Float3, rng)
    qtmh, dt = 0.5*0.01/2, 0.01*g.dx     # keeps |v| dt/2/dx well below 1/2
    pa, pb = p.copy(), p.copy()
    ca, cb = g.field(orc.Float4), g.field(orc.Float4)
    ia, ib = np.zeros(1001, np.int32), np.zeros(1001, np.int32)
    orc.push_and_deposit(pa, E, B, g, order, qtmh, dt, ia, ca, 0.0, update)
    fn = (k.push_and_deposit.push_and_deposit_cic if order == 1
          else k.push_and_deposit.push_and_deposit_tsc)
    fn(pb, E, B, qtmh, dt, ref_grid(g), ib, cb, 0.0, update)
    assert np.array_equal(bits(pa), bits(pb))
    assert np.array_equal(bits(ca), bits(cb))
    assert np.array_equal(ia, ib)
    if not update:
        assert np.array_equal(bits(pa), bits(p))


def test_push_and_deposit_cfl_flag():
    """> half a cell in half a step sets ihole[0] = -1 (push_and_deposit.pyx:66-68)"""
    k = ref.kernels()
    g = orc.Grid(nx=32, ny=32)
    rng = np.random.default_rng(5)
    p = random_particles(g, 100, rng, vth=0.01, margin=3.0)
    p["vx"][7] = 40.0
    E = g.field(orc.Float3)
    B = g.field(orc.Float3)
    dt = g.dx
    for fn_is_ref in (False, True):
        q, c = p.copy(), g.field(orc.Float4)
        ih = np.zeros(50, np.int32)
        if fn_is_ref:
            k.push_and_deposit.push_and_deposit_cic(
                q, E, B, 0.0, dt, ref_grid(g), ih, c, 0.0, False)
        else:
            orc.push_and_deposit(q, E, B, g, 1, 0.0, dt, ih, c, 0.0, False)
        assert ih[0] == -1


@pytest.mark.parametrize("gk", GRIDS)
def test_finite_differences_bitexact(gk):
    k = ref.kernels()
    fd = k.finite_difference
    g = orc.Grid(**gk)
    rg = ref_grid(g)
    rng = np.random.default_rng(6)
    f = random_field(g, orc.Float3, rng)
    s = rng.uniform(0.5, 1.5, (g.myp, g.mx))
    a, b = g.field(orc.Float3), g.field(orc.Float3)
    orc.gradient(s, a, g)
    fd.gradient(s, b, rg)
    assert np.array_equal(bits(a), bits(b))
    for down, fn in ((True, fd.curl_down), (False, fd.curl_up)):
        a, b = g.field(orc.Float3), g.field(orc.Float3)
        orc.curl(f, a, g, down=down)
        fn(f["x"], f["y"], f["z"], b, rg)
        assert np.array_equal(bits(a), bits(b))
    a, b = g.field(orc.Float3), g.field(orc.Float3)
    orc.unstagger(f, a, g)
    fd.unstagger(f["x"], f["y"], f["z"], b, rg)
    assert np.array_equal(bits(a), bits(b))
    a, b = g.field(orc.Float3), g.field(orc.Float3)
    orc.stagger(f, a, g)
    fd.stagger(f["x"], f["y"], f["z"], b, rg)
    assert np.array_equal(bits(a), bits(b))
    da, db = g.field(), g.field()
    orc.divergence(f, da, g)
    fd.divergence(f["x"], f["y"], db, rg)
    assert np.array_equal(bits(da), bits(db))


def test_move_matches_cppmove2_single_rank():
    """The migration SET and per-rank count of the oracle equal cppmove2's
    (order is not contractual: compare sorted, tests/test_skeletor.py:8-12)."""
    k = ref.kernels()
    g = orc.Grid(nx=32, ny=32)
    rng = np.random.default_rng(7)
    n, nmax = 3000, 4500
    p = np.zeros(nmax, orc.Particle)
    p[:n] = random_particles(g, n, rng, vth=20.0)
    orc.drift(p[:n], g, 0.01)          # some leave through y = 0 / ny
    orc.periodic_x(p[:n], g)
    nb = int(0.1*nmax)
    ih = np.zeros(2*nb, np.int32)
    orc.calculate_ihole(p[:n], ih, g)
    assert 0 < ih[0] < ih.shape[0]
    q = p.copy()
    bufs = [np.zeros(nb, orc.Particle) for _ in range(4)]
    info = np.zeros(7, np.int32)
    k.ppic2_wrapper.cppinit(k.MPI.COMM_WORLD)
    npp = k.ppic2_wrapper.cppmove2(q, n, bufs[0], bufs[1], bufs[2], bufs[3],
                                   ih, info, ref_grid(g))
    (mine,), (nmine,) = orc.move([p], [n], [g])
    assert npp == nmine == n
    key = lambda a: np.sort(a.view(np.float64).reshape(-1, 5), axis=0)
    assert np.array_equal(key(q[:npp]), key(mine[:nmine]))
    assert (mine["y"][:nmine] >= 0).all() and (mine["y"][:nmine] < 32).all()
