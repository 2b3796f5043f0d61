"""Poisson: E = grad del^-2 rho (reference skeletor/poisson.py:1-12,
manifolds/second_order.py:101-213).  Electrostatic tests only — scheduled after
the particle hot path."""


class PoissonSolver:

    def __init__(self, grid, ax=0.0, ay=0.0, custom_cppois22=True):
        self.grid = grid
        self.ax, self.ay = ax, ay

    def __call__(self, rho, E):
        raise NotImplementedError(
            "cuFFT Poisson solve is not built yet (SURVEY.md §8f item 2)")


class Poisson:

    """Solve Gauss' law ∇·E = ρ/ε0"""

    def __init__(self, manifold):
        self.manifold = manifold

    def __call__(self, rho, E, **kwds):
        self.manifold.grad_inv_del(rho, E, **kwds)
