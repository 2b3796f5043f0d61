"""Predictor-corrector time stepper (reference
skeletor/time_steppers/predictor_corrector.py:1-134): two push_and_deposit sweeps
per step, the second with update=False."""
from .common import StepperBase


class TimeStepper(StepperBase):

    extra_fields = ("E2", "E3", "B2")

    def __init__(self, state, ohm, manifold):
        super().__init__(state, ohm, manifold)
        self.E2.copy_guards()
        self._assign(self.B2, state.B)

    def step_noupdate(self, dt):
        self.step(dt, update=False)

    def prepare(self, dt, tol=1.48e-8, maxiter=100):
        self._prepare_common(dt, tol, maxiter, finish=lambda: None)

    def step(self, dt, update):
        super().step(dt, update)
        if update:
            self._assign(self.B, self.B2)
            self._assign(self.E, self.E2)
            self.B.boundaries_set = self.B2.boundaries_set
            self.E.boundaries_set = self.E2.boundaries_set
            self.t += dt
            self.state.t = self.t

    def iterate(self, dt):
        # Predictor step: electric field at n+1/2
        self.step(dt, update=True)
        self._assign(self.E3, self.E2)
        # Predict electric field at n+1
        self.E.t.copy_(2.0*self.E3.t - self.E.t)
        # Corrector step: electric field at n+3/2
        self.step(dt, update=False)
        # Predict electric field at n+1
        self.E.t.copy_(0.5*(self.E3.t + self.E2.t))
