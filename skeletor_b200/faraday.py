"""Faraday's law B -= dt curl E (reference skeletor/faraday.py:4-30), one fused
kernel (skb_faraday)."""
from . import _lib
from .field import Field, _stream
from .types import Float3


class Faraday:

    def __init__(self, manifold):
        self.manifold = manifold
        self.curl = manifold.curl
        # Pre-allocate array for dB
        self.dB = Field(manifold, dtype=Float3)

    def __call__(self, E, B, dt, set_boundaries=False):
        assert E.boundaries_set, 'Boundaries need to be set on E'
        _lib.call("skb_faraday", E.ptr, B.ptr, self.dB.ptr, self.manifold.c,
                  float(dt), _stream())
        B.boundaries_set = False
        # Set boundary condition on B?
        if set_boundaries:
            B.copy_guards()
