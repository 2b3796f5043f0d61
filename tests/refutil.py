"""Helpers shared by the oracle-vs-reference tests (test infrastructure)."""
import numpy as np

from oracle import oracle as orc
from oracle import ref


def ref_grid(g):
    """reference grid_t (types.pxd:25-37) mirroring an oracle Grid."""
    k = ref.kernels()
    rg = k.types.grid_t()
    rg.nx, rg.ny = g.nx, g.ny
    rg.comm = k.MPI.COMM_WORLD
    rg.edges = list(g.edges)
    rg.nyp, rg.noff = g.nyp, g.noff
    rg.lbx, rg.lby, rg.ubx, rg.uby = g.lbx, g.lby, g.ubx, g.uby
    rg.dx, rg.dy, rg.Lx, rg.Ly, rg.x0, rg.y0 = g.dx, g.dy, g.Lx, g.Ly, g.x0, g.y0
    return rg


def random_particles(g, n, rng, vth=0.3, margin=0.0):
    """n particles uniformly inside the slab of grid g (positions in cells)."""
    p = np.zeros(n, orc.Particle)
    p["x"] = rng.uniform(margin, g.nx - margin, n)
    p["y"] = rng.uniform(g.edges[0] + margin, g.edges[1] - margin, n)
    p["vx"], p["vy"], p["vz"] = rng.normal(0, vth, (3, n))
    return p


def random_field(g, dtype, rng, lo=-1.0, hi=1.0):
    f = np.zeros((g.myp, g.mx), dtype)
    for d in f.dtype.names:
        f[d] = rng.uniform(lo, hi, (g.myp, g.mx))
    return f


def bits(a):
    """view a float64 / structured-float64 array as raw uint64 for bitwise compare"""
    return np.ascontiguousarray(a).view(np.uint64)
