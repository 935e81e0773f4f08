// Geometric multigrid V-cycle as the preconditioner of CG (row N4 of SURVEY.md 8f).
//
// The reference reaches beyond Jacobi through host libraries: pyamg (solver.py:1399-1491: smoothed aggregation /
// Ruge-Stueben hierarchies used as a solver or as the preconditioner of a SciPy Krylov method) and PETSc's pc types
// (solver.py:1224-1333).  The device analogue built here keeps everything on the GPU:
//   * the hierarchy is a chain of plans -- the same model discretised on coarser meshes (linked by apdx_plan_set_coarse),
//     so that every level's operator is assembled by the same element kernels, at the state injected from the level above
//     (re-discretisation: exact for linear problems, the usual choice for Newton tangents);
//   * transfer: P = (multi-)linear interpolation between the node sets, reduced to the free dofs and built by the host
//     once per mesh (CSR), R = P^T;
//   * smoother: Chebyshev-accelerated Jacobi of fixed degree on [lmax / ratio, 1.1 lmax], lmax(D^-1 A) from a power
//     iteration after every assembly; the same polynomial before and after the coarse correction, so the cycle is a
//     symmetric operator and plain PCG applies;
//   * coarsest level: a Chebyshev polynomial of higher degree over a wider interval.
// Every level's products run through the sliced-ELL SpMV of krylov.cu.
//
// Multi-GPU (slab partitions): every level is partitioned like the finest one -- the rank that owns fine
// node plane 2I owns coarse plane I, one ghost plane per side -- so a level's vectors are [ghost_lo | owned | ghost_hi] in
// its reduced numbering, every kernel works on the owned range [f0, f1), every product is preceded by the level's halo
// exchange, the restriction by a halo exchange of the fine residual (R = P^T needs the ghost plane's rows), the
// prolongation by one of the coarse correction, and the dot products of the PCG loop / the power iteration are
// all-reduced.  The iterates are those of the single-GPU cycle up to the summation order of the dot products.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "krylov.cuh"

namespace apdx {

// ---- small vector kernels -------------------------------------------------------------------------------------
// Chebyshev / Jacobi step:  r = b - y (y == nullptr: r = b) ; d = cd d + cr minv r ; x = (zero_x ? 0 : x) + d
__global__ void __launch_bounds__(VEC_BLOCK) k_cheb_step(const double *__restrict__ b, const double *__restrict__ y,
                                                         const double *__restrict__ minv, double *__restrict__ d,
                                                         double *__restrict__ x, double cd, double cr, int zero_x, int64_t i0,
                                                         int64_t i1) {
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double r = y ? b[i] - y[i] : b[i];
    const double di = (cd != 0.0 ? cd * d[i] : 0.0) + cr * minv[i] * r;
    d[i] = di;
    x[i] = zero_x ? di : x[i] + di;
  }
}
// r = b - y (r may alias y)
__global__ void __launch_bounds__(VEC_BLOCK) k_residual(const double *__restrict__ b, const double *y, double *r, int64_t i0,
                                                        int64_t i1) {
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x)
    r[i] = b[i] - y[i];
}
// y (+)= M x for the rows [r0, r1) of a CSR matrix, one thread per row (transfer operators: 1..8 nf entries per row of
// P, <= 27 of R)
template <bool ADD>
__global__ void __launch_bounds__(256) k_csr_spmv(const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                  const double *__restrict__ val, const double *__restrict__ x,
                                                  double *__restrict__ y, int64_t r0, int64_t r1) {
  const int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  double acc = 0.0;
  for (int32_t j = ptr[i]; j < ptr[i + 1]; ++j) acc += val[j] * __ldg(x + idx[j]);
  y[i] = ADD ? y[i] + acc : acc;
}
__global__ void k_inject(const double *__restrict__ fine, const int32_t *__restrict__ inject, int64_t n,
                         double *__restrict__ coarse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) coarse[i] = fine[inject[i]];
}
// deterministic start vector of the power iteration
__global__ void k_power_start(double *__restrict__ x, int64_t i0, int64_t i1, uint64_t seed) {
  const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  uint64_t h = ((uint64_t)i + seed) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  x[i] = 0.5 + (double)(h & 0xfffff) / 1048576.0;
}
// x <- v = scale * minv * y (y = A x); sums (x.x, v.v): |v| / |x| -> scale * lambda_max(D^-1 A)
__global__ void __launch_bounds__(VEC_BLOCK) k_power_step(const double *__restrict__ y, const double *__restrict__ minv,
                                                          double *__restrict__ x, int64_t i0, int64_t i1, double scale,
                                                          double *partial, unsigned int *ticket, double *sc, int32_t *fl) {
  double acc[2] = {0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = scale * minv[i] * y[i], xi = x[i];
    acc[0] += xi * xi;
    acc[1] += v * v;
    x[i] = v;
  }
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_NONE, 1);
}
// the sliced-ELL matrix holds the owned rows [row0, row0 + n) of the level (all rows on one GPU)
__global__ void k_jacobi_inv_mg(const double *__restrict__ sell_val, const int64_t *__restrict__ valptr,
                                const int32_t *__restrict__ diag, int64_t row0, int64_t n, int nf, double *__restrict__ minv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t slice = (i / (64 * nf)) * nf + (i % nf);
  minv[row0 + i] = diag[i] >= 0 ? 1.0 / sell_val[valptr[slice] + diag[i]] : 1.0;
}

// ---- multigrid-PCG vector kernels -------------------------------------------------------------------------------
// r = b - q (q == nullptr: r = b) ; sums (r.r, b.b)
__global__ void __launch_bounds__(VEC_BLOCK) k_mg_init(const double *__restrict__ b, const double *__restrict__ q,
                                                       double *__restrict__ r, int64_t i0, int64_t i1, double *partial,
                                                       unsigned int *ticket, double *sc, int32_t *fl, int fused) {
  double acc[2] = {0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i], ri = q ? bi - q[i] : bi;
    r[i] = ri;
    acc[0] += ri * ri; acc[1] += bi * bi;
  }
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_MG_INIT, fused);
}
// sums (r.z); first != 0: p = z as well
__global__ void __launch_bounds__(VEC_BLOCK) k_mg_rz(const double *__restrict__ r, const double *__restrict__ z,
                                                     double *__restrict__ p, int first, int64_t i0, int64_t i1, double *partial,
                                                     unsigned int *ticket, double *sc, int32_t *fl, int fused) {
  if (!first && fl[F_DONE]) return;
  double acc[1] = {0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double zi = z[i];
    acc[0] += r[i] * zi;
    if (first) p[i] = zi;
  }
  reduce_finalize<1>(acc, partial, ticket, sc, fl, first ? ST_MG_RZ0 : ST_MG_RZ, fused);
}
// x += alpha p ; r -= alpha q ; sums (r.r)
__global__ void __launch_bounds__(VEC_BLOCK) k_mg_update(const double *__restrict__ p, const double *__restrict__ q,
                                                         double *__restrict__ x, double *__restrict__ r, int64_t i0, int64_t i1,
                                                         double *partial, unsigned int *ticket, double *sc, int32_t *fl, int fused) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA];
  double acc[1] = {0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    acc[0] += ri * ri;
  }
  reduce_finalize<1>(acc, partial, ticket, sc, fl, ST_MG_RR, fused);
}
// p = z + beta p
__global__ void __launch_bounds__(VEC_BLOCK) k_mg_p(const double *__restrict__ z, double *__restrict__ p, int64_t i0, int64_t i1,
                                                    const double *sc, const int32_t *fl) {
  if (fl[F_DONE]) return;
  const double beta = sc[S_BETA];
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = z[i] + beta * p[i];
}

// ---- host side ------------------------------------------------------------------------------------------------------
static unsigned vgrid(int64_t n) {
  const int64_t nb = (n + VEC_BLOCK - 1) / VEC_BLOCK;
  const int64_t cap = vec_grid();
  return (unsigned)(nb < 1 ? 1 : (nb < cap ? nb : cap));
}

// y = A x on the level's owned rows; partitioned levels refresh the ghost entries of x first
static int mg_spmv(apdx_plan *pl, double *x, double *y) {
  if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, x, pl->stream));
  return spmv_launch(pl, x, y, nullptr, 0, ST_NONE, 0);
}

static int mg_alloc(apdx_plan *pl) {
  MgLevel &m = pl->mg;
  const int64_t n = pl->n_free;
  if (!m.minv.p) {
    APDX_CHECK(m.minv.alloc(n));
    APDX_CHECK(m.d.alloc(n));
    APDX_CHECK(m.r.alloc(n));
    APDX_CHECK(m.x.alloc(n));
    APDX_CHECK(m.b.alloc(n));
    APDX_CHECK(m.ev.alloc(n));
    if (comm_active()) {   // ghost entries travel through NCCL before they are ever written: keep them finite
      for (DevBuf<double> *v : {&m.d, &m.r, &m.x, &m.b, &m.ev, &m.minv})
        APDX_CUDA(cudaMemsetAsync(v->p, 0, (size_t)n * sizeof(double), pl->stream));
    }
  }
  return krylov_alloc(pl);
}

// minv and lambda_max(D^-1 A) of the level's current matrix
int mg_level_setup(apdx_plan *pl) {
  APDX_REQUIRE(pl->have_sell_values, APDX_ERR_STATE, "multigrid level without an assembled tangent");
  APDX_CHECK(mg_alloc(pl));
  MgLevel &m = pl->mg;
  KrylovWork &k = pl->kw;
  cudaStream_t s = pl->stream;
  const int64_t i0 = pl->f0, i1 = pl->f1, n = i1 - i0;
  APDX_REQUIRE(pl->sell.n_rows == n, APDX_ERR_STATE, "multigrid level: the sliced-ELL matrix does not hold the owned rows");
  k_jacobi_inv_mg<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pl->sell.val.p, pl->sell.valptr.p, pl->sell.diag.p, i0, n,
                                                              pl->sell.nf, m.minv.p);
  // power iteration on D^-1 A; the last step's |D^-1 A x| / |x| is the estimate.  The first set-up of a plan runs 12
  // steps from a pseudo-random vector; later tangents (Newton steps, load steps: the scaled spectrum hardly moves)
  // refine the kept vector with 3 steps.  The vector is renormalised by the host-free trick of dividing by a power
  // of two of the running estimate, so that it neither over- nor underflows over many Newton steps.
  const bool first = m.lmax == 0.0;
  if (first) k_power_start<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(m.ev.p, i0, i1, 0ull);
  const int steps = first ? 12 : 3;
  for (int it = 0; it < steps; ++it) {
    APDX_CHECK(mg_spmv(pl, m.ev.p, m.r.p));
    k_power_step<<<vgrid(n), VEC_BLOCK, 0, s>>>(m.r.p, m.minv.p, m.ev.p, i0, i1, first ? 1.0 : 1.0 / m.lmax, k.partial.p,
                                                k.ticket.p, k.scal.p, k.flags.p);
    pl->stats.kernel_launches += 1;
  }
  if (comm_active()) APDX_CHECK(comm_allreduce_sum(k.scal.p + S_PEND, 2, s));   // (x.x, v.v) over all ranks
  double sc_h[2] = {0, 0};
  APDX_CUDA(cudaMemcpyAsync(sc_h, k.scal.p + S_PEND, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  double lam = (sc_h[0] > 0.0 && sc_h[1] > 0.0) ? sqrt(sc_h[1] / sc_h[0]) : 0.0;
  if (!first) lam *= m.lmax;   // the steps above stored v / lmax_old
  if (!(lam > 0.0) || lam != lam) lam = 2.0;   // degenerate level (one dof, zero matrix): any positive bound works
  m.lmax = lam;
  m.ready = true;
  return APDX_OK;
}

struct Cheb {
  double theta, delta, sigma, rho;
  Cheb(double lmax, double ratio) {
    const double b = 1.1 * lmax, a = lmax / ratio;
    theta = 0.5 * (b + a); delta = 0.5 * (b - a); sigma = theta / delta; rho = 1.0 / sigma;
  }
};

// x <- x + p_k(D^-1 A) D^-1 (b - A x): k Chebyshev steps (zero_init: x starts from zero and the first product is skipped)
static int mg_smooth(apdx_plan *pl, const double *b, double *x, bool zero_init, int degree, double ratio) {
  MgLevel &m = pl->mg;
  cudaStream_t s = pl->stream;
  const int64_t i0 = pl->f0, i1 = pl->f1;
  Cheb c(m.lmax, ratio);
  for (int j = 0; j < degree; ++j) {
    const double *y = nullptr;
    if (!(zero_init && j == 0)) {
      APDX_CHECK(mg_spmv(pl, x, m.r.p));
      y = m.r.p;
    }
    double cd, cr;
    if (j == 0) { cd = 0.0; cr = 1.0 / c.theta; }
    else {
      const double rho_n = 1.0 / (2.0 * c.sigma - c.rho);
      cd = rho_n * c.rho; cr = 2.0 * rho_n / c.delta;
      c.rho = rho_n;
    }
    k_cheb_step<<<vgrid(i1 - i0), VEC_BLOCK, 0, s>>>(b, y, m.minv.p, m.d.p, x, cd, cr, (zero_init && j == 0) ? 1 : 0, i0, i1);
    pl->stats.kernel_launches += 1;
  }
  return APDX_OK;
}

// x = V(b), from zero
static int mg_vcycle(apdx_plan *pl, const double *b, double *x, apdx_plan *top) {
  MgLevel &m = pl->mg;
  cudaStream_t s = pl->stream;
  const int64_t i0 = pl->f0, i1 = pl->f1;
  if (!m.coarse) {
    APDX_CHECK(mg_smooth(pl, b, x, true, m.coarsest, m.coarsest_ratio));
  } else {
    apdx_plan *c = m.coarse;
    const int64_t c0 = c->f0, c1 = c->f1;
    APDX_CHECK(mg_smooth(pl, b, x, true, m.pre, m.ratio));
    APDX_CHECK(mg_spmv(pl, x, m.r.p));
    k_residual<<<vgrid(i1 - i0), VEC_BLOCK, 0, s>>>(b, m.r.p, m.r.p, i0, i1);
    // R = P^T: an owned coarse row also collects from the fine ghost plane next to it
    if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, m.r.p, s));
    k_csr_spmv<false><<<(unsigned)((c1 - c0 + 255) / 256), 256, 0, s>>>(m.R.ptr.p, m.R.idx.p, m.R.val.p, m.r.p, c->mg.b.p, c0, c1);
    APDX_CHECK(mg_vcycle(c, c->mg.b.p, c->mg.x.p, top));
    // an owned fine row next to the interface interpolates from the coarse ghost plane
    if (comm_active()) APDX_CHECK(comm_halo_exchange(c, c->mg.x.p, s));
    k_csr_spmv<true><<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, s>>>(m.P.ptr.p, m.P.idx.p, m.P.val.p, c->mg.x.p, x, i0, i1);
    top->stats.kernel_launches += 3;
    APDX_CHECK(mg_smooth(pl, b, x, false, m.post, m.ratio));
  }
  if (pl != top) {   // launches of the coarse levels are counted on the plan the caller reads
    top->stats.kernel_launches += pl->stats.kernel_launches;
    top->stats.spmv_launches += pl->stats.spmv_launches;
    pl->stats.kernel_launches = 0;
    pl->stats.spmv_launches = 0;
  }
  return APDX_OK;
}

int mg_pcg_solve(apdx_plan *pl, const apdx_krylov_opts *o, const double *rhs, double *x, int32_t *iters, double *relres) {
  APDX_REQUIRE(o->method == APDX_KRYLOV_CG, APDX_ERR_UNSUPPORTED,
               "the multigrid preconditioner is used with 'solver': 'cg' (symmetric V-cycle)");
  const bool multi = comm_active();
  for (apdx_plan *l = pl; l; l = l->mg.coarse) {
    APDX_REQUIRE(l->mg.ready, APDX_ERR_STATE, "multigrid hierarchy not set up for the current tangent");
    APDX_REQUIRE(!multi || !l->hl.active, APDX_ERR_UNSUPPORTED,
                 "the partitioned multigrid hierarchy needs slab partitions (apdx_plan_set_partition) on every level");
  }
  KrylovWork &k = pl->kw;
  cudaStream_t s = pl->stream;
  const int64_t i0 = pl->f0, i1 = pl->f1, n = pl->n_free;
  const unsigned VG = vgrid(i1 - i0);
  const int fused = multi ? 0 : 1;
  const int64_t n_glob_hint = n * (int64_t)(multi ? comm_size() : 1);
  const int maxiter = o->maxiter > 0 ? o->maxiter : (int)std::min<int64_t>(10 * n_glob_hint, 2000000000ll);
  nvtx_push("apdx:krylov_multigrid");
  double sc_h[S_COUNT] = {0};
  sc_h[S_TOL2] = o->rtol * o->rtol;
  sc_h[S_SS] = o->atol * o->atol;
  int32_t fl_h[F_COUNT] = {0, 0, 0, maxiter, 0};
  APDX_CUDA(cudaMemcpyAsync(k.scal.p, sc_h, sizeof(sc_h), cudaMemcpyHostToDevice, s));
  APDX_CUDA(cudaMemcpyAsync(k.flags.p, fl_h, sizeof(fl_h), cudaMemcpyHostToDevice, s));
  double *z = k.minv.p;   // the Jacobi vector of the plain loops is free here: z = V(r)
  if (pl->x0_is_zero) {
    pl->x0_is_zero = false;
    k_mg_init<<<VG, VEC_BLOCK, 0, s>>>(rhs, nullptr, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused);
  } else {
    APDX_CHECK(mg_spmv(pl, x, k.q.p));
    k_mg_init<<<VG, VEC_BLOCK, 0, s>>>(rhs, k.q.p, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused);
  }
  APDX_CHECK(krylov_finish_stage(pl, ST_MG_INIT, 2));
  APDX_CHECK(mg_vcycle(pl, k.r.p, z, pl));
  k_mg_rz<<<VG, VEC_BLOCK, 0, s>>>(k.r.p, z, k.p.p, 1, i0, i1, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused);
  APDX_CHECK(krylov_finish_stage(pl, ST_MG_RZ0, 1));
  pl->stats.kernel_launches += 2;
  int32_t *fl_pin = reinterpret_cast<int32_t *>(pl->pinned);
  double *sc_pin = pl->pinned + 8;
  const int chunk = o->check_every > 0 ? o->check_every : 1;
  int launched = 0;
  while (true) {
    APDX_CUDA(cudaMemcpyAsync(fl_pin, k.flags.p, sizeof(int32_t) * F_COUNT, cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaMemcpyAsync(sc_pin, k.scal.p, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaStreamSynchronize(s));
    if (fl_pin[F_DONE] || launched >= maxiter) break;
    const int todo = maxiter - launched < chunk ? maxiter - launched : chunk;
    for (int it = 0; it < todo; ++it) {
      if (multi) APDX_CHECK(comm_halo_exchange(pl, k.p.p, s));
      APDX_CHECK(spmv_launch(pl, k.p.p, k.q.p, k.p.p, 1, ST_CG_PQ, 1));
      APDX_CHECK(krylov_finish_stage(pl, ST_CG_PQ, 1));
      k_mg_update<<<VG, VEC_BLOCK, 0, s>>>(k.p.p, k.q.p, x, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused);
      APDX_CHECK(krylov_finish_stage(pl, ST_MG_RR, 1));
      APDX_CHECK(mg_vcycle(pl, k.r.p, z, pl));
      k_mg_rz<<<VG, VEC_BLOCK, 0, s>>>(k.r.p, z, k.p.p, 0, i0, i1, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused);
      APDX_CHECK(krylov_finish_stage(pl, ST_MG_RZ, 1));
      k_mg_p<<<VG, VEC_BLOCK, 0, s>>>(z, k.p.p, i0, i1, k.scal.p, k.flags.p);
      pl->stats.kernel_launches += 3;
    }
    launched += todo;
  }
  nvtx_pop();
  APDX_CUDA(cudaGetLastError());
  if (multi && pl->p2p.mbox) {
    int err = 0;
    APDX_CUDA(cudaMemcpy(&err, pl->p2p.err_d, sizeof(int), cudaMemcpyDeviceToHost));
    APDX_REQUIRE(err == 0, APDX_ERR_NCCL, "peer-to-peer wait timed out (a rank stopped posting reductions)");
  }
  const double rr = sc_pin[S_BB] > 0 ? sqrt(sc_pin[S_RR] / sc_pin[S_BB]) : sqrt(sc_pin[S_RR]);
  if (iters) *iters = fl_pin[F_ITERS];
  if (relres) *relres = rr;
  pl->stats.krylov_iters += fl_pin[F_ITERS];
  pl->stats.krylov_relres = rr;
  pl->stats.krylov_converged = (sc_pin[S_RR] <= sc_pin[S_TOL2] && !fl_pin[F_BREAKDOWN]) ? 1.0 : 0.0;
  if (fl_pin[F_BREAKDOWN]) set_error("multigrid-PCG breakdown (NaN) after %d iterations", fl_pin[F_ITERS]);
  return APDX_OK;
}

// state transfer for the coarse assembly: coarse dofs = fine dofs at the coinciding nodes
int mg_inject(apdx_plan *fine, const double *fine_dofs, double *coarse_dofs) {
  apdx_plan *c = fine->mg.coarse;
  k_inject<<<(unsigned)((c->n_dofs + 255) / 256), 256, 0, fine->stream>>>(fine_dofs, fine->mg.inject.p, c->n_dofs, coarse_dofs);
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

static int upload_csr(CsrDev &M, int64_t n_rows, const int32_t *ptr_h, const int32_t *idx_h, const double *val_h) {
  M.n_rows = n_rows;
  M.nnz = ptr_h[n_rows];
  APDX_CHECK(M.ptr.alloc(n_rows + 1));
  APDX_CHECK(M.idx.alloc(M.nnz > 0 ? M.nnz : 1));
  APDX_CHECK(M.val.alloc(M.nnz > 0 ? M.nnz : 1));
  APDX_CUDA(cudaMemcpy(M.ptr.p, ptr_h, (n_rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (M.nnz > 0) {
    APDX_CUDA(cudaMemcpy(M.idx.p, idx_h, M.nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
    APDX_CUDA(cudaMemcpy(M.val.p, val_h, M.nnz * sizeof(double), cudaMemcpyHostToDevice));
  }
  return APDX_OK;
}

// the coarse plan runs on the fine plan's stream from now on
static int mg_link_finish(apdx_plan *fine, apdx_plan *coarse);

int mg_link(apdx_plan *fine, apdx_plan *coarse, const int32_t *p_ptr, const int32_t *p_idx, const double *p_val,
            const int32_t *r_ptr, const int32_t *r_idx, const double *r_val, const int64_t *inject_h) {
  APDX_REQUIRE(fine->nf == coarse->nf && fine->dim == coarse->dim, APDX_ERR_INVALID, "multigrid levels differ in dim / dofs per node");
  APDX_REQUIRE(!coarse->mg.stream_borrowed, APDX_ERR_STATE, "this plan already is the coarse level of another plan");
  MgLevel &m = fine->mg;
  APDX_CHECK(upload_csr(m.P, fine->n_free, p_ptr, p_idx, p_val));
  APDX_CHECK(upload_csr(m.R, coarse->n_free, r_ptr, r_idx, r_val));
  std::vector<int32_t> inj((size_t)coarse->n_dofs);
  for (int64_t i = 0; i < coarse->n_dofs; ++i) {
    APDX_REQUIRE(inject_h[i] >= 0 && inject_h[i] < fine->n_dofs, APDX_ERR_INVALID, "inject[%lld] = %lld outside the fine dofs",
                 (long long)i, (long long)inject_h[i]);
    inj[(size_t)i] = (int32_t)inject_h[i];
  }
  APDX_CHECK(m.inject.alloc(coarse->n_dofs));
  APDX_CUDA(cudaMemcpy(m.inject.p, inj.data(), inj.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return mg_link_finish(fine, coarse);
}

// ---- transfer operators of a structured hierarchy, built on the device -------------------------------------------------
// Coarse node (I, J, K) of a structured mesh coincides with fine node (2I, 2J, 2K); P is (multi-)linear interpolation
// reduced to the free dofs of the two plans (their own free_id / free_list), R = P^T.  Both are written row by row in
// ascending column order without a sort: a fine node has 1 or 2 coarse neighbours per direction (visited coarse index
// first = lexicographic = ascending node id), a coarse node the fine nodes 2I + {-1, 0, 1} per direction.  Weights are
// products of 1 and 1/2, so the result equals the host construction (multigrid.prolongation) to the last bit.
// Slab partitions: df[0] / dc[0] are the LOCAL plane counts, off_f / off_c the global index of local plane 0; a neighbour
// outside the local planes is dropped on both sides (it never is for an owned fine plane).
struct XferGeom {
  int32_t nf;
  int64_t df[3], dc[3];   // node counts per direction, slowest first (2-D meshes: the last one is 1)
  int64_t off_f, off_c;
};
// entries of row `r` of P (fine reduced dof): returns their number, columns ascending
__device__ __forceinline__ int prolong_row(const XferGeom &g, const int32_t *__restrict__ free_list_f,
                                           const int32_t *__restrict__ free_id_c, int64_t r, int32_t *col, double *w) {
  const int64_t dof = free_list_f[r], node = dof / g.nf;
  const int comp = (int)(dof - node * g.nf);
  const int64_t a2 = node % g.df[2], a1 = (node / g.df[2]) % g.df[1], A0 = node / (g.df[2] * g.df[1]) + g.off_f;
  const int o0 = (int)(A0 & 1), o1 = (int)(a1 & 1), o2 = (int)(a2 & 1);
  const double wt = (o0 ? 0.5 : 1.0) * (o1 ? 0.5 : 1.0) * (o2 ? 0.5 : 1.0);
  int n = 0;
  for (int b0 = 0; b0 <= o0; ++b0) {
    const int64_t c0 = (A0 >> 1) + b0 - g.off_c;
    if (c0 < 0 || c0 >= g.dc[0]) continue;
    for (int b1 = 0; b1 <= o1; ++b1)
      for (int b2 = 0; b2 <= o2; ++b2) {
        const int64_t cnode = (c0 * g.dc[1] + (a1 >> 1) + b1) * g.dc[2] + (a2 >> 1) + b2;
        const int32_t q = free_id_c[cnode * g.nf + comp];
        if (q < 0) continue;
        col[n] = q; w[n] = wt; ++n;
      }
  }
  return n;
}
// entries of row `r` of R = P^T (coarse reduced dof), columns ascending
__device__ __forceinline__ int restrict_row(const XferGeom &g, const int32_t *__restrict__ free_list_c,
                                            const int32_t *__restrict__ free_id_f, int64_t r, int32_t *col, double *w) {
  const int64_t dof = free_list_c[r], node = dof / g.nf;
  const int comp = (int)(dof - node * g.nf);
  const int64_t i2 = node % g.dc[2], i1 = (node / g.dc[2]) % g.dc[1], I0 = node / (g.dc[2] * g.dc[1]) + g.off_c;
  int n = 0;
  for (int d0 = -1; d0 <= 1; ++d0) {
    const int64_t a0 = 2 * I0 + d0 - g.off_f;
    if (a0 < 0 || a0 >= g.df[0]) continue;
    for (int d1 = -1; d1 <= 1; ++d1) {
      const int64_t a1 = 2 * i1 + d1;
      if (a1 < 0 || a1 >= g.df[1]) continue;
      for (int d2 = -1; d2 <= 1; ++d2) {
        const int64_t a2 = 2 * i2 + d2;
        if (a2 < 0 || a2 >= g.df[2]) continue;
        const int32_t q = free_id_f[((a0 * g.df[1] + a1) * g.df[2] + a2) * g.nf + comp];
        if (q < 0) continue;
        col[n] = q; w[n] = (d0 ? 0.5 : 1.0) * (d1 ? 0.5 : 1.0) * (d2 ? 0.5 : 1.0); ++n;
      }
    }
  }
  return n;
}
// RESTRICT = false: rows of P; true: rows of R.  FILL = false: cnt[r] = entries of row r (cnt[n_rows] = 0 for the scan);
// FILL = true: the row's entries at ptr[r]
template <bool RESTRICT, bool FILL>
__global__ void __launch_bounds__(256) k_transfer_rows(XferGeom g, const int32_t *__restrict__ free_list_rows,
                                                       const int32_t *__restrict__ free_id_cols, int64_t n_rows,
                                                       int32_t *__restrict__ cnt, const int32_t *__restrict__ ptr,
                                                       int32_t *__restrict__ idx, double *__restrict__ val) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  if (r == n_rows) { if (!FILL) cnt[r] = 0; return; }
  int32_t col[RESTRICT ? 27 : 8];
  double w[RESTRICT ? 27 : 8];
  const int n = RESTRICT ? restrict_row(g, free_list_rows, free_id_cols, r, col, w)
                         : prolong_row(g, free_list_rows, free_id_cols, r, col, w);
  if (!FILL) { cnt[r] = n; return; }
  const int32_t at = ptr[r];
  for (int k = 0; k < n; ++k) { idx[at + k] = col[k]; val[at + k] = w[k]; }
}
// inject[coarse full dof] = fine full dof of the coinciding node; a coarse ghost plane whose fine plane lies outside the
// local planes is pointed at the nearest local one (the caller exchanges those values: multigrid.fine_node_ids_slab)
__global__ void k_inject_map(XferGeom g, int64_t n_dofs_c, int32_t *__restrict__ inject) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dofs_c) return;
  const int64_t node = i / g.nf, comp = i - node * g.nf;
  const int64_t i2 = node % g.dc[2], i1 = (node / g.dc[2]) % g.dc[1], I0 = node / (g.dc[2] * g.dc[1]) + g.off_c;
  int64_t a0 = 2 * I0 - g.off_f;
  a0 = a0 < 0 ? 0 : (a0 > g.df[0] - 1 ? g.df[0] - 1 : a0);
  inject[i] = (int32_t)((((a0 * g.df[1]) + 2 * i1) * g.df[2] + 2 * i2) * g.nf + comp);
}

template <bool RESTRICT>
static int build_transfer(CsrDev &M, const XferGeom &g, const apdx_plan *rows, const apdx_plan *cols, cudaStream_t s) {
  const int64_t n = rows->n_free;
  const unsigned grid = (unsigned)((n + 1 + 255) / 256);
  DevBuf<int32_t> cnt;
  APDX_CHECK(cnt.alloc(n + 1));
  k_transfer_rows<RESTRICT, false><<<grid, 256, 0, s>>>(g, rows->free_list.p, cols->free_id.p, n, cnt.p, nullptr, nullptr, nullptr);
  APDX_CUDA(cudaGetLastError());
  M.n_rows = n;
  APDX_CHECK(M.ptr.alloc(n + 1));
  APDX_CHECK(scan_exclusive_i32(cnt.p, M.ptr.p, n + 1, s));   // synchronises the stream
  cnt.release();
  int32_t nnz = 0;
  APDX_CUDA(cudaMemcpy(&nnz, M.ptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost));
  APDX_REQUIRE(nnz >= 0, APDX_ERR_UNSUPPORTED, "transfer operator exceeds the 32-bit index range");
  M.nnz = nnz;
  APDX_CHECK(M.idx.alloc(nnz > 0 ? nnz : 1));
  APDX_CHECK(M.val.alloc(nnz > 0 ? nnz : 1));
  k_transfer_rows<RESTRICT, true><<<grid, 256, 0, s>>>(g, rows->free_list.p, cols->free_id.p, n, nullptr, M.ptr.p, M.idx.p, M.val.p);
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

int mg_link_structured(apdx_plan *fine, apdx_plan *coarse, int dim, const int64_t *dims_f, const int64_t *dims_c,
                       int64_t plane_off_f, int64_t plane_off_c) {
  APDX_REQUIRE(fine->nf == coarse->nf && fine->dim == coarse->dim, APDX_ERR_INVALID, "multigrid levels differ in dim / dofs per node");
  APDX_REQUIRE(!coarse->mg.stream_borrowed, APDX_ERR_STATE, "this plan already is the coarse level of another plan");
  APDX_REQUIRE(dim == 2 || dim == 3, APDX_ERR_INVALID, "structured transfer operators: dim = %d (2 or 3)", dim);
  XferGeom g{};
  g.nf = fine->nf;
  g.off_f = plane_off_f; g.off_c = plane_off_c;
  int64_t nf_nodes = 1, nc_nodes = 1;
  for (int d = 0; d < 3; ++d) {
    g.df[d] = d < dim ? dims_f[d] : 1;
    g.dc[d] = d < dim ? dims_c[d] : 1;
    APDX_REQUIRE(g.df[d] >= 1 && g.dc[d] >= 1, APDX_ERR_INVALID, "structured transfer operators: empty direction %d", d);
    // directions that are not split between ranks are halved exactly; the slowest one may carry ghost planes
    if (d > 0) APDX_REQUIRE(g.df[d] == 2 * (g.dc[d] - 1) + 1, APDX_ERR_INVALID,
                            "structured transfer operators: %lld fine and %lld coarse nodes in direction %d", (long long)g.df[d],
                            (long long)g.dc[d], d);
    nf_nodes *= g.df[d]; nc_nodes *= g.dc[d];
  }
  APDX_REQUIRE(nf_nodes == fine->n_nodes && nc_nodes == coarse->n_nodes, APDX_ERR_INVALID,
               "structured transfer operators: the node counts (%lld, %lld) do not match the plans (%lld, %lld)", (long long)nf_nodes,
               (long long)nc_nodes, (long long)fine->n_nodes, (long long)coarse->n_nodes);
  APDX_REQUIRE(plane_off_f >= 0 && plane_off_c >= 0 && 2 * plane_off_c + 1 >= plane_off_f &&
                   2 * (plane_off_c + g.dc[0] - 1) <= plane_off_f + g.df[0], APDX_ERR_INVALID,
               "structured transfer operators: coarse planes [%lld, %lld) do not lie over the fine planes [%lld, %lld)",
               (long long)plane_off_c, (long long)(plane_off_c + g.dc[0]), (long long)plane_off_f, (long long)(plane_off_f + g.df[0]));
  MgLevel &m = fine->mg;
  cudaStream_t s = fine->stream;
  APDX_CHECK(build_transfer<false>(m.P, g, fine, coarse, s));
  APDX_CHECK(build_transfer<true>(m.R, g, coarse, fine, s));
  APDX_REQUIRE(m.P.nnz == m.R.nnz, APDX_ERR_STATE, "P and R = P^T disagree on the number of entries (%lld, %lld)",
               (long long)m.P.nnz, (long long)m.R.nnz);
  APDX_CHECK(m.inject.alloc(coarse->n_dofs));
  k_inject_map<<<(unsigned)((coarse->n_dofs + 255) / 256), 256, 0, s>>>(g, coarse->n_dofs, m.inject.p);
  APDX_CUDA(cudaGetLastError());
  APDX_CUDA(cudaStreamSynchronize(s));
  return mg_link_finish(fine, coarse);
}

static int mg_link_finish(apdx_plan *fine, apdx_plan *coarse) {
  MgLevel &m = fine->mg;
  APDX_CHECK(coarse->mg.dofs.alloc(coarse->n_dofs));
  // the coarse plan runs on the fine plan's stream from now on
  if (coarse->own_stream) {
    cudaStreamSynchronize(coarse->own_stream);
    cudaStreamDestroy(coarse->own_stream);
    coarse->own_stream = nullptr;
  }
  coarse->stream = fine->stream;
  coarse->mg.stream_borrowed = true;
  for (apdx_plan *l = coarse->mg.coarse; l; l = l->mg.coarse) l->stream = fine->stream;
  m.coarse = coarse;
  m.ready = false;
  return APDX_OK;
}

}  // namespace apdx
