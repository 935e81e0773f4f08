"""jax.ffi binding of the b200 hot path (BASELINE.json north_star: "Host code stays in Python/JAX and calls a C-ABI
.so through jax.ffi custom calls").

NOT exercised in the build image: jax / jaxlib are not installed there and cannot be (no network), so the product
path this repository tests is the ctypes binding (autopdex_b200/solver.py).  This module is what a JAX installation
uses instead; importing it without JAX raises ImportError, nothing falls back to a CPU path.

    make -C autopdex_b200/csrc xla XLA_FFI_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")

builds autopdex_b200/lib/libapdx_b200_xla.so from csrc/xla/apdx_b200_xla.cc (four handlers in front of the C ABI of
include/apdx_b200.h).  The plan (pattern, index maps, sliced-ELL storage: data-dependent sizes) is created outside of
jit with backend.Plan and enters the traced function as the static integer `plan.h`; only dof-shaped FP64 arrays cross
into XLA, so the calls can sit inside jit / lax.while_loop / lax.fori_loop (load stepping, time loops):

    newton(plan, dofs, dirichlet_values)        -> solver.damped_newton           autopdex/solver.py:837-948
    linear_step(plan, dofs, dirichlet_values)   -> solver.solve_linear            autopdex/solver.py:586-659
    residual(plan, dofs)                        -> assembler.assemble_residual    autopdex/assembler.py:587-637
    tangent_solve(plan, dofs, rhs)              -> solve_fun(mat, rhs, free_dofs) autopdex/implicit_diff.py:139-183
"""
import ctypes
import os

import numpy as np

try:
    import jax
    import jax.ffi
except ImportError as exc:    # pragma: no cover - jax is not installable in the build image
    raise ImportError("autopdex_b200.jax_ffi needs jax/jaxlib (jax.ffi); use autopdex_b200.solver (ctypes binding of "
                      "the same C ABI) where JAX is not installed") from exc

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
XLA_LIB_PATH = os.environ.get("APDX_XLA_LIB") or os.path.join(_HERE, "lib", "libapdx_b200_xla.so")
_TARGETS = ("apdx_newton", "apdx_linear_step", "apdx_residual", "apdx_tangent_solve")
_registered = False


def register():
    """Register the four handlers as XLA custom-call targets on the CUDA platform (once)."""
    global _registered
    if _registered:
        return
    if not os.path.exists(XLA_LIB_PATH):
        raise ImportError("autopdex_b200.jax_ffi: %s not found; build it with `make -C autopdex_b200/csrc xla "
                          "XLA_FFI_INCLUDE=<jax.ffi.include_dir()>`" % XLA_LIB_PATH)
    _lib.load()                                            # libapdx_b200.so first: the shim links against it
    shim = ctypes.CDLL(XLA_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name in _TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(shim, name + "_ffi")), platform="CUDA")
    _registered = True


def _krylov_attrs(opts):
    c = opts.c
    return dict(method=np.int32(c.method), jacobi=np.int32(c.jacobi), rtol=np.float64(c.rtol), atol=np.float64(c.atol),
                krylov_maxiter=np.int32(c.maxiter))


def _plan_id(plan):
    return np.int64(plan.h.value)


def newton(plan, opts, dofs, dirichlet_values, newton_tol=1e-8, maxiter=30, damping=1.0):
    """(dofs, (n_steps, res_norm, diverged)) as solver.solve_newton returns them (solver.py:948)."""
    register()
    flat = dofs.reshape(-1)
    out_types = (jax.ShapeDtypeStruct(flat.shape, flat.dtype), jax.ShapeDtypeStruct((3,), flat.dtype))
    sol, infos = jax.ffi.ffi_call("apdx_newton", out_types)(
        flat, dirichlet_values.reshape(-1), plan_id=_plan_id(plan), **_krylov_attrs(opts),
        newton_tol=np.float64(newton_tol), maxiter=np.int32(maxiter), damping=np.float64(damping))
    return sol.reshape(dofs.shape), (infos[0].astype(int), infos[1], infos[2].astype(bool))


def linear_step(plan, opts, dofs, dirichlet_values):
    """The MIXED vector of solver.solve_linear (increment on free dofs, imposed values on Dirichlet dofs)."""
    register()
    flat = dofs.reshape(-1)
    delta = jax.ffi.ffi_call("apdx_linear_step", jax.ShapeDtypeStruct(flat.shape, flat.dtype))(
        flat, dirichlet_values.reshape(-1), plan_id=_plan_id(plan), **_krylov_attrs(opts))
    return delta.reshape(dofs.shape)


def residual(plan, dofs):
    register()
    flat = dofs.reshape(-1)
    res = jax.ffi.ffi_call("apdx_residual", jax.ShapeDtypeStruct(flat.shape, flat.dtype))(flat, plan_id=_plan_id(plan))
    return res.reshape(dofs.shape)


def tangent_solve(plan, opts, dofs, rhs, transpose=False):
    """K(dofs)[free][:, free]^-1 rhs[free] scattered into a dof-shaped vector, zeros on Dirichlet dofs."""
    register()
    flat = rhs.reshape(-1)
    out = jax.ffi.ffi_call("apdx_tangent_solve", jax.ShapeDtypeStruct(flat.shape, flat.dtype))(
        dofs.reshape(-1), flat, plan_id=_plan_id(plan), **_krylov_attrs(opts), transpose=np.int32(bool(transpose)))
    return out.reshape(rhs.shape)
