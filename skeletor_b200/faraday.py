"""Faraday's law: advance B by -dt curl E (reference skeletor/faraday.py:4-30).

The reference computes curl_up(E) into a scratch field and then updates the three
components of B with whole-array NumPy operations; here both happen in one kernel over
the active cells (skb_faraday).  `dB` keeps the curl for diagnostics, as the reference's
attribute of the same name does."""
from . import _lib
from .field import Field, _stream
from .types import Float3


class Faraday:

    def __init__(self, manifold):
        self.manifold = manifold
        self.curl = manifold.curl            # kept for API compatibility
        self.dB = Field(manifold, dtype=Float3)

    def __call__(self, E, B, dt, set_boundaries=False):
        assert E.boundaries_set, 'Boundaries need to be set on E'
        _lib.call("skb_faraday", E.ptr, B.ptr, self.dB.ptr, self.manifold.c, float(dt),
                  _stream())
        B.boundaries_set = False             # the guards of B are stale now
        if set_boundaries:
            B.copy_guards()
