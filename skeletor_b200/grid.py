"""Grid: slab geometry of the y-decomposed domain (reference skeletor/grid.py:5-85,
grid_t in skeletor/cython/types.pxd:25-37)."""
import numpy as np

from . import _lib


class Grid:

    def __init__(self, nx, ny, comm,
                 lbx=1, lby=1, Lx=1.0, Ly=1.0, x0=0.0, y0=0.0):
        # Number of grid points in x- and y-direction
        self.nx = nx
        self.ny = ny
        # Grid size, origin, cell size
        self.Lx = Lx
        self.Ly = Ly
        self.x0 = x0
        self.y0 = y0
        self.dx = self.Lx/self.nx
        self.dy = self.Ly/self.ny
        # communicator (mpi4py-like: skeletor_b200.comm)
        self.comm = comm
        # nyp = number of grid rows in this slab, noff = first global row
        self.nyp = ny//comm.size
        self.noff = self.nyp*comm.rank
        # edges[0:1] = lower:upper boundary of particle partition (floats)
        self.edges = [float(self.noff), float(self.noff + self.nyp)]
        # first active index / first upper guard index
        self.lbx = lbx
        self.ubx = lbx + self.nx
        self.lby = lby
        self.uby = lby + self.nyp
        # total (active plus guard) number of grid points in each subdomain
        self.mx = self.nx + 2*self.lbx
        self.myp = self.nyp + 2*self.lby

        if comm.size > self.ny:
            msg = "Too many processors requested: ny={}, comm.size={}"
            raise RuntimeError(msg.format(self.ny, comm.size))
        # the guard-cell kernels fold / copy whole guard layers in one pass
        assert self.nx >= 2*self.lbx and self.nyp >= self.lby, \
            "slab too small for its guard layers"

    @property
    def x(self):
        "One-dimensional x-coordinate array"
        return self.x0 + (np.arange(self.nx) + 0.5)*self.dx

    @property
    def y(self):
        "One-dimensional y-coordinate array"
        yrange = np.arange(self.noff, self.noff + self.nyp)
        return self.y0 + (yrange + 0.5)*self.dy

    @property
    def yg(self):
        "One-dimensional y-coordinate array including ghost"
        yrange = np.arange(self.noff - self.lby,
                           self.noff + self.nyp + self.lby)
        return self.y0 + (yrange + 0.5)*self.dy

    @property
    def c(self):
        """skb_grid_t for the C ABI"""
        g = _lib.GridT(self.nx, self.ny, self.nyp, self.noff, self.lbx, self.lby,
                       self.ubx, self.uby, self.dx, self.dy, self.Lx, self.Ly,
                       self.x0, self.y0)
        g.edges[0], g.edges[1] = self.edges
        return g
