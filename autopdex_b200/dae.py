"""`'solver backend': 'b200'` behind autopdex.dae.TimeSteppingManager.

The reference's manager (dae.py:1734-2240) advances dict dofs with per-field time integrators; for PDE problems
('user residual' domains built by models.mixed_reference_domain_residual_time) every stage assembles a sparse tangent
(`_assemble_sparse_tangent_domain`, dae.py:1826-1876) and hands it to the linear-solver backend inside a Newton
iteration (`_multi_stage_step`, dae.py:1878-2085; backend dispatch :1923-1942).  This module is that slot for the
b200 backend, for the integrators whose implicit stages are solved one after the other and have the affine form

    value = x + d,   q_t = a value + b        (x: the unknown of the stage, dae.py `_rule` of the integrator)

-- BackwardEuler (dae.py:288-318), BackwardDiffFormula (:537-587), AdamsMoulton (:420-481; history of derivatives) with
d = 0, and DiagonallyImplicitRungeKutta (:707-766; the reference solves for x_i = q_n + dt a_ii K_i and evaluates the
residual at q_n + dt sum_j a_ij K_j, i.e. d = dt sum_{j<i} a_ij K_j) -- so that a tagged transient integrand
(models.heat_conduction_time: c theta_t dtheta + k grad theta . grad dtheta - f dtheta) maps onto the device's steady +
capacity kernels with `time increment` = -1/a and `dofs n` = -b/a: the capacity kernel computes
-c/dt_eff N_i N_j (theta - theta_n,eff).  Assembly, the Newton loop (dae.newton_solver semantics: residual tolerance
`atol`, at most `max_iter` updates) and the Krylov solve run on the device; the plan (pattern, index maps, multigrid
hierarchy if asked for) is built once and reused by every stage of every step.
One difference in the Newton loop: the device runs solver.damped_newton's loop (solver.py:837-948), which also declares a
step diverged when the residual norm grows by more than 10x after the second iteration; dae.newton_solver (dae.py:1580-1708)
has no such rule and would keep iterating up to max_iter.  Both report the step as not converged unless the residual
norm falls below `atol`; the iteration counts of converging steps are equal (linear problems: 1).
Step sizes: ConstantStepSizeController (dae.py:1474-1497) or RootIterationController (:1509-1573: proportional control on
the Newton iteration count, failed steps rejected and repeated with half the step), with the accept / reject / interrupt
logic of TimeSteppingManager.run (:2150-2249).
Coupled-stage (GaussLegendreRungeKutta) / explicit / embedded (Kvaerno, DormandPrince: their explicit first stage needs
the conduction term at a state other than the unknown) integrators, Newmark (second derivatives), the PID controller
(needs an embedded error estimate) and user-written integrands are rejected (ValueError), never routed to a host path.
"""
from dataclasses import dataclass, field as _field
from typing import Any

import numpy as np

from . import solver as _solver


# ---- time integrators (implicit stages solved one after the other, each affine in its unknown) ---------------------------
class TimeIntegrator:
    """Mirror of dae.TimeIntegrator (dae.py:208-283) for the b200 backend.  q_n [num_steps, ...]: values of the last steps
    (q_n[0] = time n), q_t_n [num_steps, 1, ...]: their first derivatives, q_stages [num_stages, ...]: the stage unknowns
    (stages not solved yet hold q_n[0], dae.py:1887)."""
    num_stages, num_steps, num_derivs = 1, 1, 1
    stage_positions = (1.0,)

    def rule_coefficients(self, dt):
        """One-stage rules without a derivative history: (a, weights) with q_t = a q + sum_j weights[j] q_n[j]."""
        raise NotImplementedError

    def stage_rule(self, s, dt, q_stages, q_n, q_t_n):
        """(a, b, d) of stage s: the residual is evaluated at value = x + d and q_t = a value + b (x = the stage unknown)."""
        a, w = self.rule_coefficients(dt)
        return a, np.einsum("j,j...->...", w, q_n), 0.0

    def update(self, q_stages, q_n, q_t_n, dt):
        """(q_{n+1}, q_t_{n+1}) from the solved stages (the integrators' `_update`)."""
        a, b, _ = self.stage_rule(0, dt, q_stages, q_n, q_t_n)
        return q_stages[0], a * q_stages[0] + b


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


class AdamsMoulton(TimeIntegrator):
    """dae.AdamsMoulton (dae.py:420-481): q_t = ((q - q_n[0]) / dt - sum_{j>=1} c_j q_t_n[j-1]) / c_0 with the
    Adams-Moulton coefficients of :450-466 (num_steps 1 = trapezoidal rule).  The derivative history starts at zero
    (dae.py:1805), as in the reference."""
    name = "adams_moulton"
    _COEFFS = {1: [1 / 2, 1 / 2], 2: [5 / 12, 8 / 12, -1 / 12], 3: [9 / 24, 19 / 24, -5 / 24, 1 / 24],
               4: [251 / 720, 646 / 720, -264 / 720, 106 / 720, -19 / 720],
               5: [475 / 1440, 1427 / 1440, -798 / 1440, 482 / 1440, -173 / 1440, 27 / 1440],
               6: [19087 / 60480, 65112 / 60480, -46461 / 60480, 37504 / 60480, -20211 / 60480, 6312 / 60480, -863 / 60480]}

    def __init__(self, num_steps):
        if num_steps not in self._COEFFS:
            raise ValueError("num_steps=%r is not supported. Supported: %s" % (num_steps, list(self._COEFFS)))
        self.num_steps, self.order = num_steps, num_steps + 1

    def stage_rule(self, s, dt, q_stages, q_n, q_t_n):
        c = np.asarray(self._COEFFS[self.num_steps], dtype=np.float64)
        a = 1.0 / (dt * c[0])
        b = -a * q_n[0] - np.einsum("j,j...->...", c[1:], q_t_n[:, 0]) / c[0]
        return a, b, 0.0


class DiagonallyImplicitRungeKutta(TimeIntegrator):
    """dae.DiagonallyImplicitRungeKutta (dae.py:707-766; tableaus of Butcher 2008: implicit midpoint, Crouzeix's 3rd- and
    4th-order methods).  As in the reference (invert_butcher_with_order, dae.py:147-205: the Butcher matrix is inverted
    block by block = stage by stage) the unknown of stage i is x_i = q_n + dt a_ii K_i, the residual is evaluated at the
    stage value q_n + dt sum_{j<=i} a_ij K_j = x_i + d and K_i = (x_i - q_n) / (dt a_ii);
    q_{n+1} = q_n + dt sum_i b_i K_i, q_t_{n+1} = sum_i b_i K_i."""
    name, num_steps = "diagonally_implicit_runge_kutta", 1

    def __init__(self, num_stages):
        if num_stages == 1:
            c, b, A, self.order = [1 / 2], [1.0], [[1 / 2]], 2
        elif num_stages == 2:
            r = np.sqrt(3.0)
            c, b, A, self.order = [1 / 2 + r / 6, 1 / 2 - r / 6], [1 / 2, 1 / 2], [[1 / 2 + r / 6, 0.0], [-r / 3, 1 / 2 + r / 6]], 3
        elif num_stages == 3:
            al = 2 * np.cos(np.pi / 18) / np.sqrt(3.0)
            c = [(1 + al) / 2, 1 / 2, (1 - al) / 2]
            b = [1 / (6 * al ** 2), 1 - 1 / (3 * al ** 2), 1 / (6 * al ** 2)]
            A = [[(1 + al) / 2, 0.0, 0.0], [-al / 2, (1 + al) / 2, 0.0], [1 + al, -(1 + 2 * al), (1 + al) / 2]]
            self.order = 4
        else:
            raise ValueError("num_stages not supported for DiagonallyImplicitRungeKutta. Supported: 1, 2, 3")
        self.num_stages = num_stages
        self.stage_positions = tuple(float(v) for v in c)
        self.butcher_A, self.butcher_b, self.butcher_c = np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64), np.asarray(c)

    def _slopes(self, upto, dt, q_stages, q_n):
        return [(q_stages[j] - q_n[0]) / (dt * self.butcher_A[j, j]) for j in range(upto)]

    def stage_rule(self, s, dt, q_stages, q_n, q_t_n):
        K = self._slopes(s, dt, q_stages, q_n)
        d = dt * sum((self.butcher_A[s, j] * K[j] for j in range(s)), np.zeros_like(q_n[0]))
        a = 1.0 / (dt * self.butcher_A[s, s])
        return a, -a * (d + q_n[0]), d

    def update(self, q_stages, q_n, q_t_n, dt):
        K = self._slopes(self.num_stages, dt, q_stages, q_n)
        q_t = sum((self.butcher_b[j] * K[j] for j in range(self.num_stages)), np.zeros_like(q_n[0]))
        return q_n[0] + dt * q_t, q_t


class ConstantStepSizeController:
    """dae.ConstantStepSizeController (dae.py:1474-1497): every converged step is accepted, dt never changes; a step whose
    root solve did not converge cannot be repeated with a smaller step, so the run is interrupted."""

    def initialize(self):
        return {"step_scaler": 1.0, "dt": 1.0, "accept": True, "interrupt": False}

    def compute_scaler(self, state, converged, num_iterations, dt):
        return state

    def check_accept(self, state, converged, verbose):
        if not converged and not state["interrupt"] and verbose >= 0:
            print("Root solver did not converge, but stepsize controller can not reduce step size!")
        return dict(state, accept=bool(converged), interrupt=state["interrupt"] or not converged)


class RootIterationController(ConstantStepSizeController):
    """dae.RootIterationController (dae.py:1509-1573): proportional control of the step size on the number of Newton
    iterations of the last step, dt *= 1 + gamma (target - iterations) / target, halved when the root solve did not
    converge, clipped to [min_step_size, max_step_size]; a failed step is rejected and repeated with the new step, the
    run is interrupted when that fails at the minimum step size."""

    def __init__(self, target_niters=6, gamma=0.5, max_step_size=1e20, min_step_size=1e-6):
        self.target_niters, self.gamma = target_niters, gamma
        self.max_step_size, self.min_step_size = max_step_size, min_step_size

    def compute_scaler(self, state, converged, num_iterations, dt):
        correction = 1 + self.gamma * (self.target_niters - num_iterations) / self.target_niters if converged else 0.5
        dt_new = float(np.clip(dt * correction, self.min_step_size, self.max_step_size))
        return dict(state, step_scaler=dt_new / dt, dt=dt_new)

    def check_accept(self, state, converged, verbose):
        warn = (not converged) and bool(np.isclose(state["dt"], self.min_step_size)) and not state["interrupt"]
        if warn and verbose >= 0:
            print("Root solver did not converge, but minimum step_size is reached!")
        return dict(state, accept=bool(converged), interrupt=state["interrupt"] or warn)


@dataclass
class HistoryState:
    """dae.HistoryState (dae.py:1099-1110): t [n + 1], q {field: [n + 1, ...]}, user {name: [n + 1, ...]}; rows that were
    never written hold NaN, as in the reference's pre-allocated arrays."""
    t: Any
    q: Any
    user: Any


class SavePolicy:
    """Interface of dae.SavePolicy (dae.py:1112-1157): initialize -> state, save_step(state, t, q, user_data) -> state,
    finalize(state) -> HistoryState."""

    def initialize(self, q, t_max, max_steps, user_data=None):
        return None

    def save_step(self, state, t, q, user_data=None):
        return state

    def finalize(self, state):
        return None


class SaveNothingPolicy(SavePolicy):
    """dae.SaveNothingPolicy (dae.py:1160-1170)."""


class _ArrayHistory(SavePolicy):
    def _allocate(self, n, q, user_data):
        user_data = user_data or {}
        return {"n": int(n), "idx": 0, "t": np.full(n + 1, np.nan),
                "q": {k: np.full((n + 1,) + np.shape(v), np.nan) for k, v in q.items()},
                "user": {k: np.full((n + 1,) + np.shape(v), np.nan) for k, v in user_data.items()}}

    def _write(self, state, t, q, user_data):
        i = state["idx"]
        state["t"][i] = t
        for k in state["q"]:
            state["q"][k][i] = q[k]
        for k in state["user"]:
            state["user"][k][i] = (user_data or {})[k]
        state["idx"] = min(i + 1, state["n"])                      # clipped like the reference's index
        return state

    def finalize(self, state):
        return HistoryState(state["t"], state["q"], state["user"])


class SaveAllPolicy(_ArrayHistory):
    """dae.SaveAllPolicy (dae.py:1278-1311): (t, q, user data) of the initial state and of every accepted step, in arrays
    of max_steps + 1 rows.  For convenience the policy object also keeps the saved steps as Python lists `t` / `q`."""

    def __init__(self):
        self.t, self.q = [], []

    def initialize(self, q, t_max, max_steps, user_data=None):
        self.t, self.q = [], []
        return self._allocate(max_steps, q, user_data)

    def save_step(self, state, t, q, user_data=None):
        self.t.append(float(t))
        self.q.append({k: np.array(v) for k, v in q.items()})
        return self._write(state, t, q, user_data)


class SaveEquidistantPolicy(_ArrayHistory):
    """dae.SaveEquidistantPolicy (dae.py:1186-1264): saves the first accepted step at or after each of the num_points + 1
    equidistant target times linspace(0, t_max, num_points + 1) (num_points defaults to the maximum number of steps)."""

    def __init__(self, num_points=None, tol=1e-6):
        self.tol, self.num_points = tol, num_points

    def initialize(self, q, t_max, max_steps, user_data=None):
        n = self.num_points if self.num_points is not None else max_steps
        state = self._allocate(n, q, user_data)
        state["targets"] = np.linspace(0.0, t_max, n + 1)
        return state

    def save_step(self, state, t, q, user_data=None):
        if t >= state["targets"][state["idx"]] - self.tol:
            state = self._write(state, t, q, user_data)
        return state


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
        if static_settings.get("dae", "call pde") != "call pde":
            # the reference's ODE / DAE mode: a user-written residual function (dae.py:1975-1978) -- never routed to a host path
            raise ValueError("b200 backend: static_settings['dae'] must be 'call pde' (the assembled PDE residual); a user-written "
                             "'dae' function is not supported")
        if step_size_controller is None:
            step_size_controller = ConstantStepSizeController()
        if not isinstance(step_size_controller, ConstantStepSizeController):
            raise ValueError("b200 backend: step-size controller %r is not supported (ConstantStepSizeController, "
                             "RootIterationController; the PID controller needs an embedded error estimate)"
                             % (step_size_controller,))
        self.step_size_controller = step_size_controller
        self.integrators = dict(static_settings["time integrators"])
        if len(self.integrators) != 1:
            raise ValueError("b200 backend: one field (one time integrator) is supported, got %d" % len(self.integrators))
        for key, integ in self.integrators.items():
            if not isinstance(integ, TimeIntegrator):
                raise ValueError("b200 backend: time integrator %r of field %r is not supported (BackwardEuler, "
                                 "BackwardDiffFormula, AdamsMoulton, DiagonallyImplicitRungeKutta: implicit stages solved "
                                 "one after the other, first derivative)" % (integ, key))
        self.static_settings = dict(static_settings)
        self.static_settings.setdefault("solver type", "newton")
        if "solution structure" not in self.static_settings:
            self.static_settings["solution structure"] = ("nodal imposition",) * len(static_settings["assembling mode"])
        self.save_policy = save_policy
        self.postprocessing_fun = postprocessing_fun if postprocessing_fun is not None else (lambda q_fun, t, settings: {})
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
        """dae.TimeSteppingManager.run (dae.py:2087-2268): at most `num_time_steps` attempted steps (accepted + rejected) up to
        t_max; returns the final state, the updated settings, the save policy's history and the step statistics."""
        settings = dict(settings if settings is not None else {"current time": 0.0})
        (key, integ), = self.integrators.items()
        if list(dofs.keys()) != [key]:
            raise ValueError("b200 backend: dofs must be a dict with the field %r of the time integrator" % key)
        q = np.array(dofs[key], dtype=np.float64)
        q_n = np.repeat(q[None, ...], integ.num_steps, axis=0)            # dae.py:1804
        q_t_n = np.zeros((integ.num_steps, 1) + q.shape)                  # dae.py:1805
        dd = settings.get("dirichlet dofs")
        mask = None if dd is None else np.asarray(dd[key] if isinstance(dd, dict) else dd, dtype=bool).reshape(q.shape)
        t, t_n, dt = 0.0, 0.0, float(dt0)
        q_fun = lambda tt: {key: q}                                    # the discrete value (dae.py:2137-2138)
        history = None
        if self.save_policy is not None:                             # dae.py:2140-2144
            user_data = self.postprocessing_fun(q_fun, t, settings)
            history = self.save_policy.initialize({key: q}, t_max, int(num_time_steps), user_data)
            history = self.save_policy.save_step(history, t, {key: q}, user_data)
        accepted = rejected = 0
        newton_its = []
        ctrl = self.step_size_controller.initialize()
        for _ in range(int(num_time_steps)):                            # fori_loop over num_time_steps attempts (dae.py:2266)
            if not (t < t_max * (1 - 1e-14)) or ctrl["interrupt"]:      # dae.py:2243
                break
            t = min(t_n + dt, t_max)
            dt_step = t - t_n
            settings = self.pre_step_updates(t, settings)                # dae.py:2156
            q_stages = np.repeat(q_n[0][None, ...], integ.num_stages, axis=0)    # dae.py:1887
            its, converged, res = 0, True, 0.0
            for s in range(integ.num_stages):                            # stage blocks in order (dae.py:1897-2069)
                settings = self.pre_step_updates(t_n + dt_step * integ.stage_positions[s], settings)
                a, b, d = integ.stage_rule(s, dt_step, q_stages, q_n, q_t_n)
                step_settings = dict(settings)
                step_settings["time increment"] = -1.0 / a         # capacity kernel: -c/dt_eff (theta - theta_n,eff)
                step_settings["dofs n"] = -b / a
                shifted = isinstance(d, np.ndarray)
                if shifted and mask is not None and "dirichlet conditions" in settings:
                    # the reference constrains the stage UNKNOWN x; the device solves for the stage VALUE x + d
                    dc = settings["dirichlet conditions"]
                    vals = np.asarray(dc[key] if isinstance(dc, dict) else dc, dtype=np.float64).reshape(q.shape) + d
                    step_settings["dirichlet conditions"] = {key: vals} if isinstance(dc, dict) else vals
                guess = q_n[0] + d if shifted else q_n[0]            # initial guess: x = q (dae.py:1989)
                sol, (it_s, res, div) = _solver.solver({key: guess}, step_settings, self.static_settings, newton_tol=self.atol,
                                                       maxiter=self.max_iter, tol=self.tol, krylov_maxiter=self.krylov_maxiter)
                value = np.asarray(sol[key], dtype=np.float64)
                q_stages[s] = value - d if shifted else value
                its = max(its, int(it_s))                          # dae.py:2051
                converged = converged and (not div) and res < self.atol
            newton_its.append(its)
            if self.verbose >= 1:
                print("Time %.6e: Newton iterations %d, residual norm %.3e, converged %s" % (t, its, res, converged))
            ctrl = self.step_size_controller.compute_scaler(ctrl, converged, its, dt_step)      # dae.py:2162-2166
            ctrl = self.step_size_controller.check_accept(ctrl, converged, self.verbose)
            if ctrl["accept"] and not ctrl["interrupt"]:            # do_accept (dae.py:2175-2196)
                accepted += 1
                q, q_t = integ.update(q_stages, q_n, q_t_n, dt_step)
                q = np.asarray(q, dtype=np.float64)
                q_n = np.roll(q_n, 1, axis=0)
                q_n[0] = q
                q_t_n = np.roll(q_t_n, 1, axis=0)
                q_t_n[0, 0] = q_t
                t_n = t
                user_data = self.postprocessing_fun(q_fun, t, settings)       # dae.py:2188-2191
                if self.post_step_updates is not None:
                    settings = self.post_step_updates(q_fun, t, settings)
                if self.save_policy is not None:
                    history = self.save_policy.save_step(history, t, {key: q}, user_data)
            else:                                                   # do_reject (dae.py:2198-2203): back to t_n, new step size
                rejected += 1
                t = t_n
            dt = ctrl["step_scaler"] * dt_step
        if ctrl["interrupt"]:                                       # dae.py:2245-2249
            q = np.full_like(q, np.nan)
        steps = accepted + rejected
        history = self.save_policy.finalize(history) if self.save_policy is not None else None
        return TimeSteppingManagerState({key: q}, settings, history, steps, accepted, rejected, newton_its)


__all__ = ["TimeSteppingManager", "TimeSteppingManagerState", "BackwardEuler", "BackwardDiffFormula", "AdamsMoulton",
           "DiagonallyImplicitRungeKutta", "RootIterationController",
           "ConstantStepSizeController", "SaveAllPolicy", "SaveNothingPolicy", "SaveEquidistantPolicy", "SavePolicy",
           "HistoryState"]
