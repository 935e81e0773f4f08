// libapdx_b200_xla.so -- XLA FFI (jax.ffi) handlers in front of the C ABI of libapdx_b200.so.
//
// BASELINE.json's north_star asks for "Python/JAX host code calling a C-ABI .so through jax.ffi custom calls".  jaxlib
// and its headers (jax.ffi.include_dir() -> xla/ffi/api/ffi.h) are not installable in the build image, so this file is
// NOT part of the default build and has never been run; `make xla XLA_FFI_INCLUDE=$(python -c "import jax.ffi;
// print(jax.ffi.include_dir())")` builds it where JAX exists (tests/stubs/xla/ffi/api/ffi.h is a syntax-only stand-in
// used by tests/test_capi_and_host.py to keep this file compiling).  autopdex_b200/jax_ffi.py registers the handlers.
//
// Contract: CSR storage has data-dependent size, so it lives in the C-side plan (created from Python with
// apdx_plan_create, keyed by `plan_id` = the plan pointer); only dof-shaped FP64 arrays cross into XLA.  Every handler
// hands XLA's stream to the plan (apdx_plan_set_stream), so the plan's work is ordered with XLA's without a host wait in
// front of it: the residual handler is fully asynchronous (apdx_assemble_async); the Newton / linear-step / sensitivity
// handlers return after their loops have finished (the convergence scalars are read by the host inside them).
//   apdx_newton_ffi        solver.damped_newton        autopdex/solver.py:837-948
//   apdx_linear_step_ffi   solver.solve_linear         autopdex/solver.py:586-659
//   apdx_residual_ffi      assembler.assemble_residual autopdex/assembler.py:587-637
//   apdx_tangent_solve_ffi solve_fun(mat, rhs, free)   autopdex/implicit_diff.py:139-183, 274-304
#include <cuda_runtime_api.h>

#include <cstdint>

#include "apdx_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64In = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;

namespace {

apdx_plan *plan_of(int64_t plan_id) { return reinterpret_cast<apdx_plan *>(static_cast<intptr_t>(plan_id)); }

apdx_krylov_opts krylov_opts(int32_t method, double rtol, double atol, int32_t maxiter, int32_t jacobi) {
  apdx_krylov_opts o;
  o.method = method;
  o.maxiter = maxiter;
  o.rtol = rtol;
  o.atol = atol;
  o.jacobi = jacobi;
  o.check_every = 0;
  return o;
}

ffi::Error check_size(const apdx_plan *plan, size_t elements, const char *what) {
  int64_t q[8];
  if (apdx_plan_query(plan, q) != APDX_OK) return ffi::Error::Internal(apdx_last_error());
  if (static_cast<int64_t>(elements) != q[0])   // q[0] = n_dofs
    return ffi::Error::InvalidArgument(what);
  return ffi::Error::Success();
}

// The plan enqueues on XLA's stream: inputs produced by earlier XLA operations are ready when the plan's kernels run.
ffi::Error wait_inputs(apdx_plan *plan, cudaStream_t stream) {
  if (apdx_plan_set_stream(plan, stream) != APDX_OK) return ffi::Error::Internal(apdx_last_error());
  return ffi::Error::Success();
}

ffi::Error NewtonImpl(cudaStream_t stream, int64_t plan_id, int32_t method, int32_t jacobi, double rtol, double atol,
                      int32_t krylov_maxiter, double newton_tol, int32_t maxiter, double damping, F64In dofs,
                      F64In dirichlet_values, F64Out dofs_out, F64Out infos) {
  apdx_plan *plan = plan_of(plan_id);
  if (ffi::Error e = check_size(plan, dofs.element_count(), "dofs must have n_dofs entries"); e.failure()) return e;
  if (infos->element_count() != 3) return ffi::Error::InvalidArgument("infos must have 3 entries");
  cudaMemcpyAsync(dofs_out->typed_data(), dofs.typed_data(), dofs.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  if (ffi::Error e = wait_inputs(plan, stream); e.failure()) return e;
  const apdx_krylov_opts o = krylov_opts(method, rtol, atol, krylov_maxiter, jacobi);
  int32_t it = 0, div = 0;
  double rn = 0.0;
  if (apdx_newton(plan, &o, dofs_out->typed_data(), dirichlet_values.typed_data(), newton_tol, maxiter, damping, &it,
                  &rn, &div) != APDX_OK)
    return ffi::Error::Internal(apdx_last_error());
  const double h[3] = {static_cast<double>(it), rn, static_cast<double>(div)};   // (n_steps, res_norm, diverged)
  if (cudaMemcpyAsync(infos->typed_data(), h, sizeof(h), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess)   // `h` lives on this frame
    return ffi::Error::Internal("copy of the Newton infos failed");
  return ffi::Error::Success();
}

ffi::Error LinearStepImpl(cudaStream_t stream, int64_t plan_id, int32_t method, int32_t jacobi, double rtol,
                          double atol, int32_t krylov_maxiter, F64In dofs, F64In dirichlet_values, F64Out delta) {
  apdx_plan *plan = plan_of(plan_id);
  if (ffi::Error e = check_size(plan, dofs.element_count(), "dofs must have n_dofs entries"); e.failure()) return e;
  if (ffi::Error e = wait_inputs(plan, stream); e.failure()) return e;
  const apdx_krylov_opts o = krylov_opts(method, rtol, atol, krylov_maxiter, jacobi);
  int32_t kit = 0;
  if (apdx_linear_step(plan, &o, dofs.typed_data(), dirichlet_values.typed_data(), delta->typed_data(), &kit) != APDX_OK)
    return ffi::Error::Internal(apdx_last_error());
  if (apdx_synchronize() != APDX_OK) return ffi::Error::Internal(apdx_last_error());
  return ffi::Error::Success();
}

ffi::Error ResidualImpl(cudaStream_t stream, int64_t plan_id, F64In dofs, F64Out residual) {
  apdx_plan *plan = plan_of(plan_id);
  if (ffi::Error e = check_size(plan, dofs.element_count(), "dofs must have n_dofs entries"); e.failure()) return e;
  if (ffi::Error e = wait_inputs(plan, stream); e.failure()) return e;
  if (apdx_assemble_async(plan, dofs.typed_data(), /*want_tangent=*/0, residual->typed_data()) != APDX_OK)
    return ffi::Error::Internal(apdx_last_error());
  return ffi::Error::Success();   // asynchronous on XLA's stream
}

ffi::Error TangentSolveImpl(cudaStream_t stream, int64_t plan_id, int32_t method, int32_t jacobi, double rtol,
                            double atol, int32_t krylov_maxiter, int32_t transpose, F64In dofs, F64In rhs, F64Out out) {
  apdx_plan *plan = plan_of(plan_id);
  if (ffi::Error e = check_size(plan, rhs.element_count(), "rhs must have n_dofs entries"); e.failure()) return e;
  if (ffi::Error e = wait_inputs(plan, stream); e.failure()) return e;
  const apdx_krylov_opts o = krylov_opts(method, rtol, atol, krylov_maxiter, jacobi);
  int32_t kit = 0;
  if (apdx_tangent_solve(plan, &o, dofs.typed_data(), rhs.typed_data(), transpose, out->typed_data(), &kit) != APDX_OK)
    return ffi::Error::Internal(apdx_last_error());
  if (apdx_synchronize() != APDX_OK) return ffi::Error::Internal(apdx_last_error());
  return ffi::Error::Success();
}

}  // namespace

#define APDX_KRYLOV_ATTRS()                                                                                     \
  .Attr<int32_t>("method").Attr<int32_t>("jacobi").Attr<double>("rtol").Attr<double>("atol").Attr<int32_t>(     \
      "krylov_maxiter")

XLA_FFI_DEFINE_HANDLER_SYMBOL(apdx_newton_ffi, NewtonImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan_id") APDX_KRYLOV_ATTRS()
                                  .Attr<double>("newton_tol")
                                  .Attr<int32_t>("maxiter")
                                  .Attr<double>("damping")
                                  .Arg<F64In>()     // dofs
                                  .Arg<F64In>()     // dirichlet values
                                  .Ret<F64In>()     // dofs_out
                                  .Ret<F64In>());   // infos[3]

XLA_FFI_DEFINE_HANDLER_SYMBOL(apdx_linear_step_ffi, LinearStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan_id") APDX_KRYLOV_ATTRS()
                                  .Arg<F64In>()
                                  .Arg<F64In>()
                                  .Ret<F64In>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(apdx_residual_ffi, ResidualImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan_id")
                                  .Arg<F64In>()
                                  .Ret<F64In>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(apdx_tangent_solve_ffi, TangentSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan_id") APDX_KRYLOV_ATTRS()
                                  .Attr<int32_t>("transpose")
                                  .Arg<F64In>()
                                  .Arg<F64In>()
                                  .Ret<F64In>());
