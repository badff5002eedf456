"""The reference's integrators (vmc_fluids/stepper.py:6-91,94-145) on torch tensors.

Host glue: the right-hand side `f` these classes drive is the GPU path; the step-size arithmetic, the order and `intStep`
tags of the right-hand-side calls and the quirks are the drop-in contract and are reproduced operation by
operation, so that a run steps through the same times as the reference:

  * `step(t, f, y, normFunction=..., **rhsArgs) -> (y_new, dt_used, info)`; `f(y, t, **rhsArgs, intStep=k) -> (dy/dt, info)`;
  * FixedStepper enlarges dt BEFORE the step (stepper.py:131) and returns the info of its last right-hand side;
  * AdaptiveHeun compares one Heun step with two half steps; its error measure is `normFunction(diff, f.SExp)` as it
    comes (main.py:24-26 passes a quadratic form, no square root), the step factor is 0.9 (tol/err)^0.33333 clipped to
    [0.2, 2] (stepper.py:71-86), and the half-step point is reached as (y + dt k0) - (dt/2) k0 (stepper.py:57), which
    differs from y + (dt/2) k0 in the last bit -- kept.
"""
import numpy as np
import torch


def _default_norm(v, *args):
    return torch.linalg.norm(v)


class AdaptiveHeun:
    """stepper.py:6-91: Heun with step doubling; `dt` carries over to the next call, capped by `maxStep`."""

    def __init__(self, timeStep=1e-3, tol=1e-8, maxStep=1):
        self.dt = timeStep
        self.tolerance = tol
        self.maxStep = maxStep

    @staticmethod
    def _factor(fe):
        """stepper.py:74-82."""
        return min(max(0.2, 0.9 * fe**0.33333), 2.)

    def _attempt(self, f, t, y0, dt, rhsArgs):
        """Five right-hand sides (intStep 0..4) from y0: the full step, then the two half steps (stepper.py:50-68).
        Returns (increment of the full step, increment of the two half steps, info of the first right-hand side)."""
        k0, info = f(y0.clone(), t, **rhsArgs, intStep=0)
        y = y0 + dt * k0
        k1, _ = f(y, t + dt, **rhsArgs, intStep=1)
        coarse = 0.5 * dt * (k0 + k1)
        y = y - 0.5 * dt * k0
        k10, _ = f(y, t + 0.5 * dt, **rhsArgs, intStep=2)
        fine = 0.25 * dt * (k0 + k10)
        y = y0 + fine
        k01, _ = f(y, t + 0.5 * dt, **rhsArgs, intStep=3)
        y = y + 0.5 * dt * k01
        k11, _ = f(y, t + dt, **rhsArgs, intStep=4)
        fine = fine + 0.25 * dt * (k01 + k11)
        return coarse, fine, info

    def step(self, t, f, y, normFunction=_default_norm, **rhsArgs):
        y0 = y.clone()
        dt = self.dt
        while True:
            coarse, fine, info = self._attempt(f, t, y0, dt, rhsArgs)
            fe = self.tolerance / float(normFunction(fine - coarse, f.SExp))   # f.SExp of the fifth right-hand side
            used = dt
            dt = min(dt * self._factor(fe), self.maxStep)
            if not fe < 1.:           # accepted; a rejected attempt is repeated from y0 with the reduced dt
                break
        self.dt = dt
        return y0 + fine, used, info


class FixedStepper:
    """stepper.py:94-145: explicit Heun or Euler with dt <- min(dt * increase_fac, maxStep) before every step."""

    def __init__(self, timeStep=1e-3, maxStep=1e-2, increase_fac=1.3, mode='Heun'):
        self.dt = timeStep
        self.maxStep = maxStep
        self.mode = mode
        self.increase_fac = increase_fac

    def step(self, t, f, y, normFunction=_default_norm, **rhsArgs):
        y0 = y.clone()
        self.dt = np.min([self.dt * self.increase_fac, self.maxStep])
        dt = self.dt
        if self.mode == 'Heun':
            k0, _ = f(y0.clone(), t, **rhsArgs, intStep=0)
            k1, info = f(y0 + dt * k0, t + dt, **rhsArgs, intStep=1)
            return y0 + 0.5 * dt * (k0 + k1), dt, info
        if self.mode == 'Euler':
            k0, info = f(y0.clone(), t, **rhsArgs, intStep=0)
            return y0 + dt * k0, dt, info
        # any other mode: the reference falls through and returns None (stepper.py:129-145); so does this
