"""Generalised Ohm's law of the hybrid model (reference skeletor/ohm.py:4-75):

    E = -(T_e/q) grad ln(rho) + eta J + ((J - J_i)/rho) x B_c,    J = curl B,

with B_c the staggered B interpolated to the cell centres where E lives.  The reference
evaluates this as ~25 whole-array NumPy passes (gradient, curl, copy_guards on the
scratch current, unstagger, three cross-product updates); here it is ONE kernel over the
active cells, `skb_ohm`, with the reference's operation order per cell.
"""
from . import _lib
from .field import Field, _stream
from .types import Float3


class Ohm:

    def __init__(self, manifold, charge=1.0, temperature=0.0, eta=0.0):
        self.manifold = manifold
        self.charge, self.temperature, self.eta = charge, temperature, eta
        # diagnostics the reference also keeps: electron "velocity" (J - J_i)/rho and the
        # cell-centred magnetic field of the last call
        self.Je = Field(manifold, dtype=Float3)
        self.B = Field(manifold, dtype=Float3)
        # the manifold's operators, exposed under the reference's attribute names
        for name in ("gradient", "log", "curl", "unstagger"):
            setattr(self, name, getattr(manifold, name))

    @property
    def alpha(self):
        """electron temperature over charge"""
        return self.temperature/self.charge

    def __call__(self, sources, B, E, set_boundaries=False):
        # the stencils read one guard layer of rho and B (second_order.py:25,36,51)
        assert sources.boundaries_set and B.boundaries_set, \
            'Boundaries need to be set on sources and B'
        _lib.call("skb_ohm", sources.ptr, B.ptr, E.ptr, self.Je.ptr, self.B.ptr,
                  self.manifold.c, float(self.alpha), float(self.eta), _stream())
        E.boundaries_set = False
        if set_boundaries:
            E.copy_guards()
