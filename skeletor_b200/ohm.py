"""Ohm's law E = -(T/q) grad ln rho + eta J + (J - J_i)/rho x B, J = curl B
(reference skeletor/ohm.py:4-75), as ONE fused kernel over the active cells
(skb_ohm) instead of ~25 whole-array NumPy passes."""
from . import _lib
from .field import Field, _stream
from .types import Float3


class Ohm:

    """Solve Ohm's law"""

    def __init__(self, manifold, charge=1.0, temperature=0.0, eta=0.0):
        self.manifold = manifold
        # operators kept for API compatibility (ohm.py:11-15)
        self.gradient = manifold.gradient
        self.log = manifold.log
        self.curl = manifold.curl
        self.unstagger = manifold.unstagger
        self.charge = charge
        self.temperature = temperature
        self.eta = eta
        # electron current and interpolated B-field (diagnostics, ohm.py:24-28)
        self.Je = Field(manifold, dtype=Float3)
        self.B = Field(manifold, dtype=Float3)

    @property
    def alpha(self):
        # Ratio of temperature to charge
        return self.temperature/self.charge

    def __call__(self, sources, B, E, set_boundaries=False):
        assert sources.boundaries_set and B.boundaries_set, \
            'Boundaries need to be set on sources and B'
        _lib.call("skb_ohm", sources.ptr, B.ptr, E.ptr, self.Je.ptr, self.B.ptr,
                  self.manifold.c, float(self.alpha), float(self.eta), _stream())
        E.boundaries_set = False
        # Set boundary condition on E?
        if set_boundaries:
            E.copy_guards()
