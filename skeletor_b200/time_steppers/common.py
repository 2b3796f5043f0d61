"""Pieces shared by the two time steppers (reference
skeletor/time_steppers/horowitz.py, predictor_corrector.py)."""
import math


from ..faraday import Faraday
from ..field import Field
from ..sources import Sources
from ..types import Float3


class StepperBase:
    """Common state: sources, E, B and scratch fields; deposit-all-species,
    half-step `step`, L2 difference.  Field arithmetic between whole arrays is done
    with torch elementwise ops on the device tensors in the reference's operation
    order (so values match NumPy's to the bit); the field solves are the fused
    Ohm/Faraday kernels and the particle sweeps are push_and_deposit."""

    extra_fields = ()

    def __init__(self, state, ohm, manifold):
        self.state = state
        # Numerical grid with differential operators
        self.manifold = manifold
        # Ohm's law / Faraday's law
        self.ohm = ohm
        self.faraday = Faraday(manifold)
        # Initialize sources
        self.sources = Sources(manifold)
        # Set the electric field to zero
        self.E = Field(manifold, dtype=Float3)
        self.E.fill((0.0, 0.0, 0.0))
        self.E.copy_guards()
        self.B = state.B
        for name in self.extra_fields:
            f = Field(manifold, dtype=Float3)
            setattr(self, name, f)
        self.t = state.t

    # -- helpers ---------------------------------------------------------------
    @staticmethod
    def _assign(dst, src):
        """dst[:] = src for whole fields (values and nothing else)"""
        dst.t.copy_(src.t)

    def _sum_species(self, fn):
        """zero the total sources, run fn(ions) for every species and add up their
        sources (horowitz.py:46-51, 97-103)"""
        self.sources.t.zero_()
        self.sources.boundaries_set = False
        for ions in self.state.species:
            fn(ions)
            self.sources.t.add_(ions.sources.t)
        self.sources.boundaries_set = True

    def calculate_diff(self, f, g):
        """sqrt(sum over ranks of sum_dim mean((f-g)^2 over active cells) / size)
        (horowitz.py:112-121)"""
        from ..comm import SUM
        m = self.manifold
        d = (f.t - g.t)[m.lby:m.uby, m.lbx:m.ubx]
        diff2 = float((d*d).mean(dim=(0, 1)).sum().item())
        comm = m.comm
        return math.sqrt(comm.allreduce(diff2, op=SUM)/comm.size)

    def step(self, dt, update):
        """Half-step scheme shared by both steppers (predictor_corrector.py:85-115):
        B2,E2 <- B,E; Faraday dt/2; push_and_deposit; Faraday dt/2; Ohm."""
        self._assign(self.B2, self.B)
        self._assign(self.E2, self.E)
        self.B2.boundaries_set = self.B.boundaries_set
        self.E2.boundaries_set = self.E.boundaries_set
        # Evolve magnetic field by a half step to n (n+1)
        self.faraday(self.E2, self.B2, dt/2, set_boundaries=True)
        # Push particle positions to n+1 (n+2) and kick velocities to n+1/2
        # (n+3/2); deposit at n+1/2 (n+3/2); only update particles if update=True
        self._sum_species(lambda ions: ions.push_and_deposit(self.E2, self.B2, dt, update))
        # Evolve magnetic field by a half step to n+1/2 (n+3/2)
        self.faraday(self.E2, self.B2, dt/2, set_boundaries=True)
        # Electric field at n+1/2 (n+3/2)
        self.ohm(self.sources, self.B2, self.E2, set_boundaries=True)

    def _prepare_common(self, dt, tol, maxiter, finish):
        # Deposit sources
        self._sum_species(lambda ions: ions.deposit(set_boundaries=True))
        # Calculate electric field (Solve Ohm's law)
        self.ohm(self.sources, self.B, self.E, set_boundaries=True)
        # Drift particle positions by a half time step
        for ions in self.state.species:
            ions.drift(dt/2)
        # Iterate to find true electric field at time 0
        for it in range(maxiter):
            # Compute electric field at time 1/2
            self.step_noupdate(dt)
            # Average to get electric field at time 0
            self.E3.t.copy_(0.5*(self.E.t + self.E2.t))
            # Compute difference to previous iteration
            diff = self.calculate_diff(self.E3, self.E)
            if self.manifold.comm.rank == 0:
                print("Difference to previous iteration: {}".format(diff))
            # Update electric field
            self._assign(self.E, self.E3)
            # Return if difference is sufficiently small
            if diff < tol:
                finish()
                return
        raise RuntimeError("Exceeded maxiter={} iterations!".format(maxiter))
