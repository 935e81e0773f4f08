"""`'solver backend': 'b200'` behind autopdex.dae.TimeSteppingManager.

The reference's manager (dae.py:1734-2240) advances dict dofs with per-field time integrators; for PDE problems
('user residual' domains built by models.mixed_reference_domain_residual_time) every stage assembles a sparse tangent
(`_assemble_sparse_tangent_domain`, dae.py:1826-1876) and hands it to the linear-solver backend inside a Newton
iteration (`_multi_stage_step`, dae.py:1878-2085; backend dispatch :1923-1942).  This module is that slot for the
b200 backend, for the integrators whose single implicit stage has the affine form

    q_t = a q + b                      (BackwardEuler dae.py:288-318, BackwardDiffFormula dae.py:537-587)

so that a tagged transient integrand (models.heat_conduction_time: c theta_t dtheta + k grad theta . grad dtheta -
f dtheta) maps onto the device's steady + capacity kernels with `time increment` = -1/a and `dofs n` = -b/a: the
capacity kernel computes -c/dt_eff N_i N_j (theta - theta_n,eff).  Assembly, the Newton loop (dae.newton_solver
semantics: residual tolerance `atol`, at most `max_iter` updates) and the Krylov solve run on the device; the plan
(pattern, index maps, multigrid hierarchy if asked for) is built once and reused by every step.
Multi-stage / explicit integrators, step-size controllers other than the constant one and user-written integrands
are rejected (ValueError), never routed to a host path.
"""
from dataclasses import dataclass, field as _field
from typing import Any

import numpy as np

from . import solver as _solver


# ---- time integrators (one implicit stage, affine in the new value) ---------------------------------------------------
class TimeIntegrator:
    num_stages, num_derivs = 1, 1

    def rule_coefficients(self, dt):
        """(a, weights): q_t = a q + sum_j weights[j] q_n[j]."""
        raise NotImplementedError


class BackwardEuler(TimeIntegrator):
    """dae.BackwardEuler (dae.py:288-318): q_t = (q - q_n[0]) / dt.  Order 1, one step."""
    name, num_steps, order = "backward_euler", 1, 1

    def rule_coefficients(self, dt):
        return 1.0 / dt, np.array([-1.0 / dt])


class BackwardDiffFormula(TimeIntegrator):
    """dae.BackwardDiffFormula (dae.py:537-587): q_t = (c0 q + sum_j c_{j+1} q_n[j]) / dt with the BDF coefficients of
    :566-573.  The history starts as copies of the initial value (dae.py:1804), as in the reference."""
    name = "backward_diff_formula"
    _COEFFS = {1: [1, -1], 2: [3 / 2, -2, 1 / 2], 3: [11 / 6, -3, 3 / 2, -1 / 3], 4: [25 / 12, -4, 3, -4 / 3, 1 / 4],
               5: [137 / 60, -5, 5, -10 / 3, 5 / 4, -1 / 5], 6: [49 / 20, -6, 15 / 2, -20 / 3, 15 / 4, -6 / 5, 1 / 6]}

    def __init__(self, num_steps):
        if num_steps not in self._COEFFS:
            raise ValueError("Order of BDF method not supported. Supported orders: 1 to 6. From 7 on the BDF method is not stable.")
        self.num_steps = self.order = num_steps

    def rule_coefficients(self, dt):
        c = np.asarray(self._COEFFS[self.num_steps], dtype=np.float64)
        return c[0] / dt, c[1:] / dt


class ConstantStepSizeController:
    """dae.ConstantStepSizeController (dae.py:1474-1497): every converged step is accepted, dt never changes."""


class SaveAllPolicy:
    """dae.SaveAllPolicy (dae.py:1278-1311): history of (t, q) for every accepted step, the initial state included."""

    def __init__(self):
        self.t, self.q = [], []

    def save(self, t, q):
        self.t.append(float(t))
        self.q.append({k: np.array(v) for k, v in q.items()})


@dataclass
class TimeSteppingManagerState:
    """Return value of run(), fields as dae.TimeSteppingManagerState (dae.py:1715-1731)."""
    q: Any
    settings: Any
    history: Any
    num_steps: int
    num_accepted: int
    num_rejected: int
    newton_iterations: list = _field(default_factory=list)


class TimeSteppingManager:
    """autopdex.dae.TimeSteppingManager with `static_settings['solver backend'] == 'b200'`.

    static_settings: 'time integrators' {field: integrator}, 'assembling mode' ('user residual', ...), 'model'
    (models.mixed_reference_domain_residual_time(...) for the transient domain; steady 'user residual' / surface
    'user element' domains may accompany it), 'solver' 'cg' | 'bicgstab', 'type of preconditioner', 'verbose'.
    Newton keyword arguments as dae.newton_solver: atol (1e-8), max_iter (20); Krylov: tol, krylov_maxiter."""

    def __init__(self, static_settings, settings=None, root_solver=None, save_policy=None,
                 step_size_controller=None, postprocessing_fun=None, pre_step_updates=None, post_step_updates=None,
                 atol=1e-8, max_iter=20, tol=1e-10, krylov_maxiter=0):
        if static_settings.get("solver backend") != "b200":
            raise ValueError("autopdex_b200.dae handles 'solver backend': 'b200' only")
        if root_solver is not None:
            raise ValueError("b200 backend: the Newton iteration runs on the device; root_solver cannot be replaced")
        if step_size_controller is not None and not isinstance(step_size_controller, ConstantStepSizeController):
            raise ValueError("b200 backend: only ConstantStepSizeController is supported")
        self.integrators = dict(static_settings["time integrators"])
        if len(self.integrators) != 1:
            raise ValueError("b200 backend: one field (one time integrator) is supported, got %d" % len(self.integrators))
        for key, integ in self.integrators.items():
            if not isinstance(integ, TimeIntegrator):
                raise ValueError("b200 backend: time integrator %r of field %r is not supported (BackwardEuler, "
                                 "BackwardDiffFormula: one implicit stage, first derivative)" % (integ, key))
        self.static_settings = dict(static_settings)
        self.static_settings.setdefault("solver type", "newton")
        if "solution structure" not in self.static_settings:
            self.static_settings["solution structure"] = ("nodal imposition",) * len(static_settings["assembling mode"])
        self.save_policy = save_policy
        self.postprocessing_fun = postprocessing_fun
        if pre_step_updates is None:
            def pre_step_updates(t, settings):                 # dae.py:1768-1772
                settings["current time"] = t
                return settings
        self.pre_step_updates = pre_step_updates
        self.post_step_updates = post_step_updates
        self.verbose = static_settings.get("verbose", 0)
        self.atol, self.max_iter, self.tol, self.krylov_maxiter = atol, max_iter, tol, krylov_maxiter
        self.cfg = _solver._Config(self.static_settings)       # settings-read-time rejection
        if not self.cfg.transient:
            raise ValueError("b200 backend: TimeSteppingManager needs a time-dependent 'user residual' domain "
                             "(models.mixed_reference_domain_residual_time)")

    def run(self, dofs, dt0, t_max, num_time_steps, settings=None):
        """dae.TimeSteppingManager.run (dae.py:2087-2240) with a constant step size."""
        settings = dict(settings if settings is not None else {"current time": 0.0})
        (key, integ), = self.integrators.items()
        if list(dofs.keys()) != [key]:
            raise ValueError("b200 backend: dofs must be a dict with the field %r of the time integrator" % key)
        q = np.array(dofs[key], dtype=np.float64)
        q_n = np.repeat(q[None, ...], integ.num_steps, axis=0)            # dae.py:1804
        t, t_n, dt = 0.0, 0.0, float(dt0)
        if self.save_policy is not None:
            self.save_policy.save(t, {key: q})
        accepted = rejected = steps = 0
        newton_its = []
        while steps < num_time_steps and t_n < t_max * (1 - 1e-14):
            t = min(t_n + dt, t_max)
            dt_step = t - t_n
            settings = self.pre_step_updates(t, settings)
            a, w = integ.rule_coefficients(dt_step)
            b = np.einsum("j,j...->...", w, q_n)
            step_settings = dict(settings)
            step_settings["time increment"] = -1.0 / a             # capacity kernel: -c/dt_eff (theta - theta_n,eff)
            step_settings["dofs n"] = -b / a
            sol, (its, res, div) = _solver.solver({key: q_n[0]}, step_settings, self.static_settings, newton_tol=self.atol,
                                                  maxiter=self.max_iter, tol=self.tol, krylov_maxiter=self.krylov_maxiter)
            steps += 1
            converged = (not div) and res < self.atol
            newton_its.append(int(its))
            if self.verbose >= 1:
                print("Time %.6e: Newton iterations %d, residual norm %.3e, converged %s" % (t, its, res, converged))
            if not converged:                                       # constant controller: interrupt (dae.py:1486-1497)
                rejected += 1
                break
            accepted += 1
            q = np.asarray(sol[key], dtype=np.float64)
            q_n = np.roll(q_n, 1, axis=0)
            q_n[0] = q
            t_n = t
            if self.post_step_updates is not None:
                settings = self.post_step_updates(lambda tt: {key: q}, t, settings)
            if self.save_policy is not None:
                self.save_policy.save(t, {key: q})
        return TimeSteppingManagerState({key: q}, settings, self.save_policy, steps, accepted, rejected, newton_its)


__all__ = ["TimeSteppingManager", "TimeSteppingManagerState", "BackwardEuler", "BackwardDiffFormula",
           "ConstantStepSizeController", "SaveAllPolicy"]
