"""CPU oracle vs the committed golden fixtures (tests/golden/*.npz, generated from
the unmodified reference by oracle/make_golden.py), single slab and N slabs.

Single slab: particles bit-exact (same per-particle arithmetic), fields bit-exact
(same summation order).  N slabs: the reference's own invariant "N ranks == 1
rank" (reference tests/test_skeletor.py:142-150), here to 1e-12.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
import oracle_runs as runs
from refutil import bits

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def flat(a):
    return np.ascontiguousarray(a).view(np.float64).ravel()


def check_fields(a, b, rtol):
    a, b = flat(a), flat(b)
    scale = np.abs(b).max()
    assert np.abs(a - b).max() <= rtol*scale


@pytest.mark.parametrize("name,order,lb", [("ionacoustic_cic", 1, 1),
                                           ("ionacoustic_tsc", 2, 2)])
@pytest.mark.parametrize("nslabs", [1, 2, 4])
def test_ionacoustic(name, order, lb, nslabs):
    g = gold(name)
    r = runs.ionacoustic(nslabs, order=order, lb=lb)
    s = r["slabs"]
    assert sum(s.N) == int(g["N"])
    got = runs.sorted_particles(s.gathered())
    exp = runs.sorted_particles(g["particles"])
    a = (slice(lb, -lb), slice(lb, -lb))
    if nslabs == 1:
        assert np.array_equal(got, exp)
        assert np.array_equal(bits(r["sources"][0]), bits(g["sources"]))
        assert np.array_equal(bits(r["E"][0]), bits(g["E"]))
    else:
        assert np.abs(got - exp).max() < 1e-12
        check_fields(runs.active_cat(r["sources"], r["grids"]), g["sources"][a], 1e-12)
        check_fields(runs.active_cat(r["E"], r["grids"]), g["E"][a], 1e-11)


@pytest.mark.parametrize("name,order,Omega", [("sheared_cic", 1, 1.0),
                                              ("sheared_tsc", 2, 0.0)])
@pytest.mark.parametrize("nslabs", [1, 2, 4])
def test_sheared(name, order, Omega, nslabs):
    g = gold(name)
    r = runs.sheared(nslabs, order=order, Omega=Omega)
    s = r["slabs"]
    assert sum(s.N) == int(g["N"])
    assert s.time == float(g["time"])
    got = runs.sorted_particles(s.gathered())
    exp = runs.sorted_particles(g["particles"])
    a = (slice(2, -2), slice(2, -2))
    if nslabs == 1:
        assert np.array_equal(got, exp)
        assert np.array_equal(bits(r["sources"][0]), bits(g["sources"]))
    else:
        assert np.array_equal(got, exp)      # particle work is slab-independent
        check_fields(runs.active_cat(r["sources"], r["grids"]), g["sources"][a], 1e-12)


@pytest.mark.parametrize("name,order", [("gyro_cic", 1), ("gyro_tsc", 2)])
@pytest.mark.parametrize("nslabs", [1, 4])
def test_gyro(name, order, nslabs):
    g = gold(name)
    r = runs.gyro_fields(nslabs, order=order)
    s = r["slabs"]
    assert sum(s.N) == int(g["N"])
    got = runs.sorted_particles(s.gathered())
    exp = runs.sorted_particles(g["particles"])
    a = (slice(2, -2), slice(2, -2))
    if nslabs == 1:
        assert np.array_equal(got, exp)
        assert np.array_equal(bits(r["sources"][0]), bits(g["sources"]))
    else:
        # the gather offset lby - noff differs per slab, so y + offset rounds
        # differently than on one rank: N ranks == 1 rank only to rounding
        assert np.abs(got - exp).max() < 1e-12
        check_fields(runs.active_cat(r["sources"], r["grids"]), g["sources"][a], 1e-12)


@pytest.mark.parametrize("shear", [False, True])
def test_guards(shear):
    g = gold("guards_shear" if shear else "guards_plain")
    kw = dict(nx=16, ny=8, lbx=1, lby=2)
    if shear:
        grid = orc.Grid(S=-1.5, Omega=0.0, Lx=2.0, Ly=1.0, **kw)
    else:
        grid = orc.Grid(**kw)
    rng = np.random.default_rng(14)
    src = grid.field(orc.Float4)
    for d in "txyz":
        src[d][...] = rng.uniform(-1, 1, (grid.myp, grid.mx))
    orc.add_guards([src], [grid], 0.37)
    assert np.array_equal(bits(src), bits(g["added"]))
    orc.copy_guards([src], [grid], 0.37)
    assert np.array_equal(bits(src), bits(g["copied"]))
    f = grid.field()
    f[...] = rng.uniform(-1, 1, (grid.myp, grid.mx))
    orc.copy_guards([f], [grid], 0.37)
    assert np.array_equal(bits(f), bits(g["scalar"]))


def test_ohm_faraday():
    g = gold("ohm_faraday")
    grid = orc.Grid(nx=16, ny=16, lbx=1, lby=1, Lx=1.0, Ly=2.0)
    rng = np.random.default_rng(15)
    src = grid.field(orc.Float4)
    src["t"][...] = rng.uniform(0.5, 1.5, (grid.myp, grid.mx))
    for d in "xyz":
        src[d][...] = rng.uniform(-1, 1, (grid.myp, grid.mx))
    orc.copy_guards([src], [grid])
    B = grid.field(orc.Float3)
    for d in "xyz":
        B[d][...] = rng.uniform(-1, 1, (grid.myp, grid.mx))
    orc.copy_guards([B], [grid])
    E = grid.field(orc.Float3)
    orc.ohm(src, B, E, grid, charge=1.3, temperature=0.7, eta=0.05)
    orc.copy_guards([E], [grid])
    orc.faraday(E, B, grid, 0.01)
    orc.copy_guards([B], [grid])
    assert np.array_equal(bits(E), bits(g["E"]))
    assert np.array_equal(bits(B), bits(g["B"]))


def test_sort_keys_are_tile_major():
    grid = orc.Grid(nx=64, ny=32, lbx=2, lby=2)
    rng = np.random.default_rng(0)
    p = np.zeros(1000, orc.Particle)
    p["x"] = rng.uniform(0, 64, 1000)
    p["y"] = rng.uniform(0, 32, 1000)
    k = orc.cell_keys(p, grid, 1, (3, 2))
    assert k.min() >= 0
    perm = np.argsort(k, kind="stable")
    assert (np.diff(k[perm]) >= 0).all()
