"""Horowitz iterative time stepper (reference
skeletor/time_steppers/horowitz.py:1-171): one push_and_deposit sweep per step and
up to `maxiter` Faraday/Ohm iterations for the fields at n+1."""
import torch

from .. import _lib
from ..field import _stream
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
        self._state = None      # device-side loop control of iterate(): [done, iterations]
        self._acc = None

    def step_noupdate(self, dt):
        # "Step method from predictor-corrector but update is always false"
        self.step(dt, update=False)

    def prepare(self, dt, tol=1.48e-8, maxiter=100):
        # on convergence: evolve magnetic field by a half step to t=t0
        self._prepare_common(
            dt, tol, maxiter,
            finish=lambda: self.faraday(self.E, self.B, dt/2, set_boundaries=True))

    def iterate(self, dt, tol=1.48e-8, maxiter=12, check_every=3):
        """Update fields and particles using Horowitz method (horowitz.py:123-171).

        The field iteration runs on the device: per iteration two fused averages
        (skb_field_combine), Faraday from B into B3 (skb_faraday_to), Ohm (skb_ohm_if),
        the E3 update with the residual reduced in the same pass (skb_horowitz_update) and
        the convergence test (skb_converged) that raises a device flag every kernel of the
        following iterations checks.  The host only looks at that flag every `check_every`
        iterations, so a step that converges in 3 iterations costs ONE synchronisation
        instead of three; the values are those of the reference's loop (same operations
        in the same order; iterations past the converged one are skipped on the device)."""
        m = self.manifold
        comm = m.comm
        st = _stream()
        # Push and deposit the particles, depositing the sources at n+1/2
        self._sum_species(lambda ions: ions.push_and_deposit(self.E, self.B, dt, True))
        # Start iteration by assuming E^(n+1) = E^n
        self._assign(self.E3, self.E)
        if self._state is None:
            dev = self.E.t.device
            self._state = torch.zeros(2, dtype=torch.int32, device=dev)
            self._acc = torch.zeros(1, dtype=torch.float64, device=dev)
        state, acc = self._state, self._acc
        state.zero_()
        skip = state.data_ptr()
        n3 = self.E.t.numel()
        scale = 1.0/(m.nx*m.nyp*comm.size)
        ohm = self.ohm
        done = False
        for it in range(maxiter):
            # Average electric field to estimate it at n + 1/2: E2 = 0.5*(E3 + E)
            _lib.call("skb_field_combine", self.E2.ptr, self.E3.ptr, self.E.ptr, n3, 0.5, 0,
                      skip, st)
            # Estimate magnetic field at n+1: B3 = B - dt curl(E2)
            _lib.call("skb_faraday_to", self.E2.ptr, self.B.ptr, self.B3.ptr,
                      self.faraday.dB.ptr, m.c, float(dt), skip, st)
            self.B3.boundaries_set = False
            self.B3.copy_guards()
            # Estimate magnetic field at n+1/2: B2 = 0.5*(B3 + B)
            _lib.call("skb_field_combine", self.B2.ptr, self.B3.ptr, self.B.ptr, n3, 0.5, 0,
                      skip, st)
            self.B2.boundaries_set = True
            # Estimate electric field at n+1/2 (Ohm's law)
            _lib.call("skb_ohm_if", self.sources.ptr, self.B2.ptr, self.E2.ptr, ohm.Je.ptr,
                      ohm.B.ptr, m.c, float(ohm.alpha), float(ohm.eta), skip, st)
            self.E2.boundaries_set = False
            self.E2.copy_guards()
            # New estimate for E^(n+1) = -E + 2 E2 and its distance to the previous one
            acc.zero_()
            _lib.call("skb_horowitz_update", self.E3.ptr, self.E.ptr, self.E2.ptr, m.c,
                      acc.data_ptr(), skip, st)
            comm.allreduce_tensor_(acc)
            _lib.call("skb_converged", acc.data_ptr(), scale, float(tol), it,
                      state.data_ptr(), st)
            if (it + 1) % check_every == 0 or it == maxiter - 1:
                done = bool(state[0].item())
                if done:
                    break
        if not done:
            raise RuntimeError("Exceeded maxiter={} iterations!".format(maxiter))
        # Update E and B
        self._assign(self.E, self.E3)
        self._assign(self.B, self.B3)
        self.E.boundaries_set = self.B.boundaries_set = True
        self.t += dt
        self.state.t = self.t
