"""Manifold / ShearingManifold: the slab grid plus second-order finite-difference
operators (reference skeletor/manifolds/second_order.py:5-98), dispatching to the
stencil kernels of libskeletor_b200 (finite_difference.pyx:5-85)."""
import numpy as np
import torch

from .. import _lib
from ..field import Field, _stream
from ..grid import Grid


def _plane(f):
    """(device pointer, element stride in doubles) of a scalar field or of a
    component view of an interleaved field"""
    t = f.t
    g = f.grid
    assert t.dim() == 2 and t.shape == (g.myp, g.mx)
    es = t.stride(1)
    assert t.stride(0) == g.mx*es, "unsupported field layout"
    return t.data_ptr(), es


class Manifold(Grid):

    """Finite difference operators"""

    def __init__(self, nx, ny, comm,
                 ax=0.0, ay=0.0, custom_cppois22=True, **grid_kwds):

        super().__init__(nx, ny, comm, **grid_kwds)

        err = 'Not enough guard layers for second order finite difference.'
        assert self.lbx >= 1 and self.lby >= 1, err

        # Poisson solver (cuFFT), built on first use (second_order.py:18)
        self._poisson_args = dict(ax=ax, ay=ay, custom_cppois22=custom_cppois22)
        self._grad_inv_del = None

    @property
    def grad_inv_del(self):
        if self._grad_inv_del is None:
            from ..poisson import PoissonSolver
            self._grad_inv_del = PoissonSolver(self, **self._poisson_args)
        return self._grad_inv_del

    def gradient(self, f, grad):
        """Calculate the gradient of f"""
        msg = 'Boundaries need to be set on f for second order differences'
        assert f.boundaries_set, msg
        p, es = _plane(f)
        _lib.call("skb_gradient", p, es, grad.ptr, self.c, _stream())
        grad.boundaries_set = False

    def curl(self, f, curl, down=True):
        """Calculate the curl of f"""
        msg = 'Boundaries need to be set on f for second order differences'
        assert f.boundaries_set, msg
        (px, es), (py, _), (pz, _) = _plane(f['x']), _plane(f['y']), _plane(f['z'])
        _lib.call("skb_curl", px, py, pz, es, curl.ptr, self.c, int(bool(down)),
                  _stream())
        curl.boundaries_set = False

    def _interp(self, f, g, up, set_boundaries):
        msg = 'Boundaries need to be set on f for interpolation'
        assert f.boundaries_set, msg
        (px, es), (py, _), (pz, _) = _plane(f['x']), _plane(f['y']), _plane(f['z'])
        _lib.call("skb_interp", px, py, pz, es, g.ptr, self.c, up, _stream())
        g.boundaries_set = False
        if set_boundaries:
            g.copy_guards()

    def unstagger(self, f, g, set_boundaries=False):
        """Interpolate the staggered field f to cell centers"""
        self._interp(f, g, 0, set_boundaries)

    def stagger(self, f, g, set_boundaries=False):
        """Interpolate the cell-centered field f to cell corners"""
        self._interp(f, g, 1, set_boundaries)

    def divergence(self, f, g):
        """Calculate the divergence of the vector field f"""
        (px, es), (py, _) = _plane(f['x']), _plane(f['y'])
        _lib.call("skb_divergence", px, py, es, g.ptr, self.c, _stream())

    def log(self, f):
        """elementwise log of a scalar field, as a new field (guards included)"""
        out = Field(self, time=f.time, dtype=np.float64, _tensor=torch.log(f.t))
        out.boundaries_set = f.boundaries_set
        return out


class ShearingManifold(Manifold):

    """Finite difference operators in the shearing sheet"""

    def __init__(self, nx, ny, comm, S=0, Omega=0, **manifold_kwds):

        super().__init__(nx, ny, comm, **manifold_kwds)

        # Shear parameter
        self.S = S

        # Angular frequency
        self.Omega = Omega
