"""Sources: charge and current density (rho, Jx, Jy, Jz) deposited from particles.

Mirrors skeletor.Sources (reference skeletor/sources.py:5-192)."""

from . import _lib
from .field import Field, _stream
from .types import Float4


class Sources(Field):

    def __init__(self, manifold, **kwds):
        kwds.pop("dtype", None)
        super().__init__(manifold, dtype=Float4, **kwds)

    @property
    def rho(self):
        return self['t']

    @property
    def Jx(self):
        return self['x']

    @property
    def Jy(self):
        return self['y']

    @property
    def Jz(self):
        return self['z']

    def deposit(self, particles, erase=True, set_boundaries=False):
        """sources.py:27-50"""
        if particles.order not in (1, 2):
            msg = 'Interpolation order {} not implemented.'
            raise RuntimeError(msg.format(particles.order))
        # Rate of shear
        S = getattr(self.grid, 'S', 0.0)
        if particles.fuse_deposit == "auto" and particles._last_op == "push":
            particles._fuse_next = True    # deposit follows push: fuse the next pair
        particles._last_op = "deposit"
        fused = particles._fused_for(self, S)
        if fused is not None:
            # the push already accumulated these sums (skb_push_deposit_gapped)
            if erase:
                self.t.copy_(fused)
            else:
                self.t.add_(fused)
            self.boundaries_set = False
            self.normalize(particles)
            if set_boundaries:
                self.set_boundaries()
            return
        if erase:
            self.t.zero_()
        if getattr(particles, "_pending", None) is not None and \
                particles._rep == "gapped" and not particles.deterministic:
            # the push's migration is still in flight on the side stream: deposit the
            # cells as the kernel left them while the neighbour exchange runs, then add
            # the rows that are inserted when it completes (summation order only)
            _lib.call("skb_deposit", particles._c, particles._N, self.ptr, self.grid.c,
                      particles.order, float(S), particles._tiling_c(), _stream())
            particles._finish_pending(deposit_into=self)
            self.boundaries_set = False
            self.normalize(particles)
            if set_boundaries:
                self.set_boundaries()
            return
        if particles.deterministic:
            particles._dense()         # the fixed-order deposit needs the dense ordering
        particles._ensure_sorted()
        if particles.deterministic and particles._sorted and particles.N > 0:
            _lib.call("skb_deposit_deterministic", particles._c, particles.N, self.ptr,
                      self.grid.c, particles.order, float(S), particles._tiling_c(),
                      particles._cellsums().data_ptr(), _stream())
        else:
            # (gapped layout: the tiling names the slot range of every cell, N unused)
            _lib.call("skb_deposit", particles._c, particles.N, self.ptr, self.grid.c,
                      particles.order, float(S), particles._tiling_c(), _stream())
            if particles._rep == "gapped" and particles._gap_nleft > 0:
                lo = particles._leftover
                _lib.call("skb_deposit", particles._soa(lo), particles._gap_nleft,
                          self.ptr, self.grid.c, particles.order, float(S), None,
                          _stream())
        self.boundaries_set = False
        self.normalize(particles)
        if set_boundaries:
            self.set_boundaries()

    def normalize(self, particles):
        """Normalize the charge and current densities such that the mean charge
        density is equal to particle.n0 (sources.py:52-63; guards included)."""
        N = particles.N_global()
        fac = particles.charge*particles.n0*self.grid.nx*self.grid.ny/N
        _lib.call("skb_scale", self.ptr, self.t.numel(), float(fac), _stream())

    def set_boundaries(self):
        self.add_guards()
        self.copy_guards()

    def __iadd__(self, other):
        """+= over all components (sources.py:72-89)"""
        self.t.add_(other.t)
        return self

    def add_guards_x(self):
        """Add the x guard cells to the corresponding active cells, all rows
        (sources.py:91-101)"""
        _lib.call("skb_add_guards", self.ptr, self.nc, self.grid.c, 0, None, None,
                  _stream())

    def add_guards_y(self):
        """Add the y guard rows to the corresponding active rows, active x only; the
        guards are NOT zeroed (sources.py:103-115)"""
        g = self.grid
        t = self.t
        ax = slice(g.lbx, g.ubx)
        up = t[g.uby:, ax].contiguous()       # my upper guards: for the rank above
        dn = t[:g.lby, ax].contiguous()       # my lower guards: for the rank below
        if g.comm.size > 1:
            from_below, from_above = self._halo_exchange(up, dn)
        else:
            from_below, from_above = up, dn
        t[g.uby:, ax] = from_below
        t[:g.lby, ax] = from_above
        for iy in range(g.lby):
            t[iy + g.nyp, ax] += t[iy, ax]
        for iy in range(g.uby + g.lby - 1, g.uby - 1, -1):
            t[iy - g.nyp, ax] += t[iy, ax]

    def add_guards_old(self):
        """reference test helper (sources.py:152-192): same result as add_guards when
        there is no shear"""
        shear, self.shear = self.shear, False
        try:
            self.add_guards()
        finally:
            self.shear = shear

    def add_guards(self):
        "Add data from guard cells to corresponding active cells (sources.py:117-150)."
        g = self.grid
        gc = g.c
        # x-boundaries, all rows
        _lib.call("skb_add_guards", self.ptr, self.nc, gc, 0, None, None, _stream())
        if self.shear:
            # Translate the y-ghostzones (sources.py:128-139)
            if g.comm.rank == g.comm.size - 1:
                self._translate_boundary(g.Ly*g.S*self.time, g.uby, g.lby)
            if g.comm.rank == 0:
                self._translate_boundary(-g.Ly*g.S*self.time, 0, g.lby)
        # y-boundaries (+ zeroing of all guards)
        if g.comm.size == 1:
            _lib.call("skb_add_guards", self.ptr, self.nc, gc, 1, None, None, _stream())
        else:
            # my upper guards go up, my lower guards go down (sources.py:110-111)
            up = self._pack_rows(g.uby, g.lby)
            dn = self._pack_rows(0, g.lby)
            from_below, from_above = self._halo_exchange(up, dn)
            _lib.call("skb_add_guards", self.ptr, self.nc, gc, 1,
                      from_below.data_ptr(), from_above.data_ptr(), _stream())
