"""Grid: geometry of one y-slab of the decomposed domain.

Attribute names and meaning follow the reference (skeletor/grid.py:5-85 and the
`grid_t` extension type, skeletor/cython/types.pxd:25-37); `Grid.c` packs them into the
`skb_grid_t` POD of the C ABI.
"""
import numpy as np

from . import _lib


class Grid:

    def __init__(self, nx, ny, comm,
                 lbx=1, lby=1, Lx=1.0, Ly=1.0, x0=0.0, y0=0.0):
        if comm.size > ny:
            msg = "Too many processors requested: ny={}, comm.size={}"
            raise RuntimeError(msg.format(ny, comm.size))
        self.comm = comm                       # mpi4py-like: skeletor_b200.comm
        # global box: cells, physical size, origin, cell size
        self.nx, self.ny = nx, ny
        self.Lx, self.Ly = Lx, Ly
        self.x0, self.y0 = x0, y0
        self.dx, self.dy = Lx/nx, Ly/ny
        # this rank's slab: nyp rows starting at global row noff; particles live in
        # edges[0] <= y < edges[1] (floats, compared against particle y directly)
        self.nyp = ny//comm.size
        self.noff = self.nyp*comm.rank
        self.edges = [float(self.noff), float(self.noff + self.nyp)]
        # guard layers: lb* = index of the first active cell, ub* = first upper guard
        self.lbx, self.lby = lbx, lby
        self.ubx, self.uby = lbx + nx, lby + self.nyp
        # extent of the stored arrays, guards included
        self.mx, self.myp = nx + 2*lbx, self.nyp + 2*lby
        # the guard-cell kernels fold / copy whole guard layers in one pass
        assert nx >= 2*lbx and self.nyp >= lby, "slab too small for its guard layers"

    def _centres(self, first, count, origin, spacing):
        return origin + (np.arange(first, first + count) + 0.5)*spacing

    @property
    def x(self):
        "cell-centre x coordinates of the active cells"
        return self._centres(0, self.nx, self.x0, self.dx)

    @property
    def y(self):
        "cell-centre y coordinates of this slab's active rows"
        return self._centres(self.noff, self.nyp, self.y0, self.dy)

    @property
    def yg(self):
        "cell-centre y coordinates of this slab's rows including the guard rows"
        return self._centres(self.noff - self.lby, self.myp, self.y0, self.dy)

    @property
    def c(self):
        """skb_grid_t for the C ABI"""
        g = _lib.GridT(self.nx, self.ny, self.nyp, self.noff, self.lbx, self.lby,
                       self.ubx, self.uby, self.dx, self.dy, self.Lx, self.Ly,
                       self.x0, self.y0)
        g.edges[0], g.edges[1] = self.edges
        return g
