"""Horowitz iterative time stepper (reference
skeletor/time_steppers/horowitz.py:1-171): one push_and_deposit sweep per step and
up to `maxiter` Faraday/Ohm iterations for the fields at n+1."""
import torch

from .common import StepperBase


class TimeStepper(StepperBase):

    extra_fields = ("E2", "E3", "B2", "B3", "E4")

    def __init__(self, state, ohm, manifold):
        super().__init__(state, ohm, manifold)
        self.E2.copy_guards()
        self.E3.copy_guards()
        self.B2.copy_guards()
        self.B3.copy_guards()
        self._assign(self.B2, state.B)

    def step_noupdate(self, dt):
        # "Step method from predictor-corrector but update is always false"
        self.step(dt, update=False)

    def prepare(self, dt, tol=1.48e-8, maxiter=100):
        # on convergence: evolve magnetic field by a half step to t=t0
        self._prepare_common(
            dt, tol, maxiter,
            finish=lambda: self.faraday(self.E, self.B, dt/2, set_boundaries=True))

    def iterate(self, dt, tol=1.48e-8, maxiter=12):
        """Update fields and particles using Horowitz method"""
        # Push and deposit the particles, depositing the sources at n+1/2
        self._sum_species(lambda ions: ions.push_and_deposit(self.E, self.B, dt, True))
        # Start iteration by assuming E^(n+1) = E^n
        self._assign(self.E3, self.E)
        for it in range(maxiter):
            self._assign(self.E4, self.E3)
            # Average electric field to estimate it at n + 1/2
            self.E2.t.copy_(0.5*(self.E3.t + self.E.t))
            self.E2.boundaries_set = True
            # Estimate magnetic field at n+1
            self._assign(self.B3, self.B)
            self.faraday(self.E2, self.B3, dt, set_boundaries=True)
            # Estimate magnetic field at n+1/2
            self.B2.t.copy_(0.5*(self.B3.t + self.B.t))
            self.B2.boundaries_set = True
            # Estimate electric field at n+1
            self.ohm(self.sources, self.B2, self.E2, set_boundaries=True)
            # New estimate for E^(n+1)
            self.E3.t.copy_(torch.neg(self.E.t) + 2.0*self.E2.t)
            diff = self.calculate_diff(self.E3, self.E4)
            # Update E and B if difference is sufficiently small
            if diff < tol:
                self._assign(self.E, self.E3)
                self._assign(self.B, self.B3)
                self.t += dt
                self.state.t = self.t
                return
        raise RuntimeError("Exceeded maxiter={} iterations!".format(maxiter))
