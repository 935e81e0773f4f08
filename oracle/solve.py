"""Oracle Newton loop, linear step and adaptive load stepping (test infrastructure).

Restates solver.solve_linear (solver.py:544-659), solver.damped_newton (:837-948),
solver.linear_solve_scipy (:1493-1539) and solver.adaptive_load_stepping (:296-379, :455-457)
with NumPy control flow instead of lax.while_loop.
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import assemble as asm


class Problem:
    """sets / coords / settings bundle; dofs are (n_nodes, nf) arrays.
    dirichlet_mask (n_nodes, nf) bool, dirichlet_values same shape float."""

    def __init__(self, sets, coords, dirichlet_mask, dirichlet_values, settings=None):
        self.sets = sets
        self.coords = np.asarray(coords, dtype=np.float64)
        self.mask = np.asarray(dirichlet_mask, dtype=bool)
        self.values = np.asarray(dirichlet_values, dtype=np.float64)
        self.settings = dict(settings or {})
        self._coo = None
        self.threads = 1        # bench.py's CPU baseline sets this (chunked element loops on a thread pool)

    def coo(self):
        if self._coo is None:
            self._coo = asm.coo_indices(self.sets)
        return self._coo

    def residual(self, dofs):
        return asm.assemble(self.sets, self.coords, dofs, self.settings, want_tangent=False, threads=self.threads)[0]


def linear_solve_scipy(prob, data, rhs, free, solver="lapack", krylov_tol=None):
    """solver.py:1493-1539: reduced CSR, spsolve, scatter into zeros(n).
    solver 'cg'/'bicgstab' = SciPy Krylov with Jacobi M (CPU analogue of solver.py:1093-1126)."""
    rows, cols = prob.coo()
    n = rhs.shape[0]
    csr = asm.scipy_assembling(data, rows, cols, n, free)
    b = rhs[free] if free is not None else rhs
    if solver in ("lapack", "umfpack"):
        x = spla.spsolve(csr, b)
    else:
        d = csr.diagonal()
        M = spla.LinearOperator(csr.shape, matvec=lambda v: v / d)
        fun = spla.cg if solver == "cg" else spla.bicgstab
        x, _ = fun(csr, b, M=M, rtol=krylov_tol or 1e-12, atol=0.0, maxiter=200000)
    sol = np.zeros(n)
    if free is not None:
        sol[free] = x
    else:
        sol[:] = x
    return sol


def solve_linear(prob, dofs, solver="lapack", krylov_tol=None, nodal_imposition=True):
    """solver.py:586-659.  Returns the MIXED vector: free entries = delta, Dirichlet
    entries = imposed values (with nodal imposition); plain delta otherwise."""
    dofs = np.array(dofs, dtype=np.float64)
    free = None
    if nodal_imposition:
        dofs[prob.mask] = prob.values[prob.mask]                        # :586-604
        free = ~prob.mask.ravel()
    R, data = asm.assemble(prob.sets, prob.coords, dofs, prob.settings)
    sol = linear_solve_scipy(prob, data, -R, free, solver, krylov_tol)  # :610-647
    if nodal_imposition:
        out = dofs.ravel().copy()
        out[free] = sol[free]                                           # :648-656
        return out.reshape(dofs.shape)
    return sol.reshape(dofs.shape)


def damped_newton(prob, dofs0, newton_tol=1e-8, maxiter=30, damping=1.0, solver="lapack",
                  krylov_tol=None, nodal_imposition=True, history=None, lin_solve_fun=None, residual_fun=None,
                  free=None):
    """solver.py:872-948.  Returns (dofs, (n_steps, res_norm, diverged)).
    lin_solve_fun / residual_fun / free override the problem's own (used to pin the loop semantics)."""
    dofs = np.array(dofs0, dtype=np.float64)
    if free is None:
        free = ~prob.mask if nodal_imposition else np.ones(dofs.shape, dtype=bool)
    itt, not_stop, res_norm, diverged = 0, True, 0.0, False
    while not_stop:
        res_old = res_norm
        delta = (lin_solve_fun(dofs) if lin_solve_fun is not None
                 else solve_linear(prob, dofs, solver, krylov_tol, nodal_imposition))
        dofs = np.where(free, dofs + damping * delta, delta)           # :879-892
        R = (residual_fun(dofs) if residual_fun is not None else prob.residual(dofs)).reshape(dofs.shape)
        rf = np.where(free, R, 0.0).ravel()                             # :900 (mask_select zero-fills)
        res_norm = float(np.linalg.norm(rf))
        not_stop = res_norm > newton_tol                                # :904
        if itt < maxiter:                                               # :930
            with np.errstate(divide="ignore", invalid="ignore"):
                div = bool((np.float64(res_norm) / np.float64(res_old) > 10) and itt > 1)  # :915-917
            if np.any(np.isnan(rf) | np.isinf(rf)):
                div = True
            next_step, diverged = (not div), div
        else:
            next_step, diverged = False, True                           # :926-928
        if history is not None:
            history.append(res_norm)
        itt += 1
        not_stop = not_stop and next_step
    return dofs, (itt, res_norm, diverged)


def adaptive_load_stepping(prob, dofs0, multiplier_settings, max_multiplier=1.0, min_increment=0.01,
                           max_increment=1.0, init_increment=0.2, target_num_newton_iter=7,
                           newton_tol=1e-10, solver="lapack", krylov_tol=None, trace=None):
    """solver.py:296-379,455-457 (no implicit differentiation)."""
    dofs = np.array(dofs0, dtype=np.float64)
    m, inc, res = 0.0, init_increment, 0.0
    while m < max_multiplier and inc > min_increment:                   # :298-300
        m += inc
        multiplier_settings(prob, m)                                    # :323
        new, (steps, res, div) = damped_newton(prob, dofs, newton_tol=newton_tol, solver=solver,
                                               krylov_tol=krylov_tol)
        if trace is not None:
            trace.append((m, steps, div))
        if div:                                                         # :356-371
            m -= inc
            inc *= 0.5
        else:
            inc *= 1 + 0.5 * (target_num_newton_iter - steps) / target_num_newton_iter
            dofs = new
        inc = min(inc, max_increment)                                   # :363
        if m + inc > max_multiplier:                                    # :374-378
            inc = max_multiplier - m
    return dofs, (m, inc, res)
