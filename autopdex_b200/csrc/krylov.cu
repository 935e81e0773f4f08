// Jacobi-preconditioned CG / BiCGSTAB on the reduced CSR system, all scalars on device.
//
// Device analogue of solver.linear_solve_jax (solver.py:1093-1126: M = 1/diag,
// jax.scipy.sparse.linalg.cg / bicgstab) operating on the assembled reduced matrix that the
// reference hands to SciPy (solver.py:1207-1217).  The stopping rule is the one of
// jax.scipy.sparse.linalg: ||r||_2^2 <= max(rtol^2 ||b||^2, atol^2).  No positivity checks:
// the reference's tangents are negative definite on mesher quad meshes.
//
// SpMV: sliced-ELL with per-slice compressed column indices (sell.cu), 128-bit value loads, the dot
// product(s) that follow the SpMV fused into the same kernel (warp-shuffle + fixed-order block reduction).  Vector updates are
// fused axpy+dot kernels.  Dot products are reduced per block, then by the last block in a fixed
// order (deterministic for a given grid), then -- multi-GPU -- by ncclAllReduce.
#include "common.cuh"

namespace apdx {

// device scalar slots
enum {
  S_RZ = 0, S_PQ, S_ALPHA, S_BETA, S_RR, S_TOL2, S_BB, S_RHO, S_OMEGA, S_TS, S_TT, S_R0V, S_SS,
  S_PEND = 16,  // pending (locally reduced) sums of the running stage, up to 4
  S_COUNT = 24
};
enum { F_DONE = 0, F_ITERS = 1, F_BREAKDOWN = 2, F_MAXITER = 3, F_COUNT = 4 };
// stages of the scalar recurrences
enum { ST_CG_INIT = 0, ST_CG_PQ, ST_CG_UPDATE, ST_BI_INIT, ST_BI_R0V, ST_BI_S, ST_BI_T, ST_BI_X };

constexpr int VEC_BLOCK = 256;
constexpr int VEC_GRID = 148 * 8;

__device__ __forceinline__ void apply_stage(int stage, double *sc, int32_t *fl) {
  const double *pd = sc + S_PEND;
  switch (stage) {
    case ST_CG_INIT:  // pend = (r.z, r.r, b.b)
      sc[S_RZ] = pd[0]; sc[S_RR] = pd[1]; sc[S_BB] = pd[2];
      {
        double t = sc[S_TOL2] /*rtol^2*/ * pd[2];
        double a2 = sc[S_SS] /*atol^2 parked here by the host*/;
        sc[S_TOL2] = t > a2 ? t : a2;
      }
      if (!(sc[S_RR] > sc[S_TOL2])) fl[F_DONE] = 1;
      break;
    case ST_CG_PQ:  // pend = (p.q)
      sc[S_PQ] = pd[0];
      sc[S_ALPHA] = sc[S_RZ] / pd[0];
      break;
    case ST_CG_UPDATE:  // pend = (r.z, r.r) after the update
      sc[S_BETA] = pd[0] / sc[S_RZ];
      sc[S_RZ] = pd[0];
      sc[S_RR] = pd[1];
      fl[F_ITERS] += 1;
      if (!(pd[1] > sc[S_TOL2]) || fl[F_ITERS] >= fl[F_MAXITER]) fl[F_DONE] = 1;
      if (pd[1] != pd[1]) { fl[F_DONE] = 1; fl[F_BREAKDOWN] = 1; }
      break;
    case ST_BI_INIT:  // pend = (r0.r, r.r, b.b)
      sc[S_RHO] = pd[0]; sc[S_RR] = pd[1]; sc[S_BB] = pd[2];
      {
        double t = sc[S_TOL2] * pd[2];
        double a2 = sc[S_SS];
        sc[S_TOL2] = t > a2 ? t : a2;
      }
      sc[S_ALPHA] = 1.0; sc[S_OMEGA] = 1.0; sc[S_BETA] = 0.0;
      if (!(sc[S_RR] > sc[S_TOL2])) fl[F_DONE] = 1;
      break;
    case ST_BI_R0V:  // pend = (r0.v)
      sc[S_ALPHA] = sc[S_RHO] / pd[0];
      break;
    case ST_BI_S:  // pend = (s.s)
      sc[S_SS] = pd[0];
      break;
    case ST_BI_T:  // pend = (t.s, t.t); early exit of jax's bicgstab: s already converged -> omega = 0
      sc[S_OMEGA] = (sc[S_SS] > sc[S_TOL2]) ? pd[0] / pd[1] : 0.0;
      break;
    case ST_BI_X:  // pend = (r0.r, r.r) of the new residual
      sc[S_BETA] = (pd[0] / sc[S_RHO]) * (sc[S_ALPHA] / sc[S_OMEGA]);
      sc[S_RHO] = pd[0];
      sc[S_RR] = pd[1];
      fl[F_ITERS] += 1;
      if (!(pd[1] > sc[S_TOL2]) || fl[F_ITERS] >= fl[F_MAXITER]) fl[F_DONE] = 1;
      if (pd[1] != pd[1] || pd[0] == 0.0) { fl[F_DONE] = 1; fl[F_BREAKDOWN] = (pd[1] > sc[S_TOL2]) ? 1 : 0; }
      break;
  }
}

__global__ void k_apply_stage(int stage, double *sc, int32_t *fl) {
  if (fl[F_DONE] && stage != ST_CG_INIT && stage != ST_BI_INIT) return;
  apply_stage(stage, sc, fl);
}

// spin until *flag >= epoch (system-scope visibility); bounded so that a lost peer cannot hang the GPU
__device__ __forceinline__ void p2p_wait(const int *flag, int epoch, int *err) {
  const long long t0 = clock64();
  while (*(volatile const int *)flag < epoch) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s
      *err = 1;
      break;
    }
  }
  __threadfence_system();
}

// all-reduce by mailboxes: wait for every rank's post of this epoch, sum in rank order (identical bits on all
// ranks), then run the scalar recurrence
__global__ void k_apply_stage_p2p(int stage, int nv, double *sc, int32_t *fl, const P2PDev *pd, int epoch) {
  if (fl[F_DONE] && stage != ST_CG_INIT && stage != ST_BI_INIT) return;
  const int slot = epoch & 1;
  if ((int)threadIdx.x < pd->nranks) p2p_wait(pd->mflag[pd->rank] + slot * P2P_MAX_RANKS + threadIdx.x, epoch, pd->err);
  __syncthreads();
  if (threadIdx.x == 0) {
    const volatile double *mb = pd->mbox[pd->rank] + (size_t)slot * P2P_MAX_RANKS * 4;
    for (int i = 0; i < nv; ++i) {
      double s = 0.0;
      for (int r = 0; r < pd->nranks; ++r) s += mb[r * 4 + i];
      sc[S_PEND + i] = s;
    }
    apply_stage(stage, sc, fl);
  }
}

// copy the owned boundary entries of a heap vector into the neighbours' ghost ranges (peer stores over
// NVLink), then raise their halo flags for this epoch
__global__ void __launch_bounds__(VEC_BLOCK) k_halo_push(const double *__restrict__ v, int64_t f0, int64_t f1,
                                                         int64_t send_lo, int64_t send_hi, double *peer_lo_dst,
                                                         double *peer_hi_dst, int *peer_lo_flag, int *peer_hi_flag,
                                                         int epoch, unsigned int *ticket, const int32_t *fl) {
  if (fl[F_DONE]) return;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (peer_lo_dst)
    for (int64_t i = tid; i < send_lo; i += nth) peer_lo_dst[i] = v[f0 + i];
  if (peer_hi_dst)
    for (int64_t i = tid; i < send_hi; i += nth) peer_hi_dst[i] = v[f1 - send_hi + i];
  __threadfence_system();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence_system();
    if (peer_lo_flag) *(volatile int *)peer_lo_flag = epoch;
    if (peer_hi_flag) *(volatile int *)peer_hi_flag = epoch;
  }
}

// Block-level reduction of NV running sums, then cross-block reduction by the last block.
// fused != 0: the last block also applies the scalar stage (single-GPU path).
template <int NV>
__device__ __forceinline__ void reduce_finalize(double (&v)[NV], double *partial, unsigned int *ticket,
                                                double *sc, int32_t *fl, int stage, int fused, const P2PDev *pd,
                                                int epoch) {
  __shared__ double sh[NV][VEC_BLOCK / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[i][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += sh[i][w];
      partial[(size_t)i * gridDim.x + blockIdx.x] = x;
    }
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) x += __ldcg(&partial[(size_t)i * gridDim.x + b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) sh[i][wid] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double y = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) y += sh[i][w];
      sc[S_PEND + i] = y;
    }
  }
  if (threadIdx.x == 0) {
    *ticket = 0u;
    if (fused) apply_stage(stage, sc, fl);
  }
  if (pd) {
    // multi-GPU without NCCL: post the local sums into every rank's mailbox over NVLink (peer stores), then
    // raise that rank's flag for this epoch; k_apply_stage_p2p on each rank sums the mailboxes in rank order.
    __syncthreads();
    if ((int)threadIdx.x < pd->nranks) {
      const int r = threadIdx.x, slot = epoch & 1;
      double *mb = pd->mbox[r] + ((size_t)slot * P2P_MAX_RANKS + pd->rank) * 4;
#pragma unroll
      for (int i = 0; i < NV; ++i) mb[i] = sc[S_PEND + i];
      __threadfence_system();
      *(volatile int *)(pd->mflag[r] + slot * P2P_MAX_RANKS + pd->rank) = epoch;
    }
  }
}

// Sliced-ELL SpMV (layout: sell.cu).  One warp per slice of 64 rows, two rows per lane, values read as
// 128-bit double2 (512 contiguous bytes per warp instruction).  Offset mode: column = row + off[j]
// (one broadcast int per slice column, x gathers coalesced); explicit mode: int2 column pairs.
// n_cols: length of x (clamp target of the padded offsets, whose values are exact zeros).
template <int NDOT>
__global__ void __launch_bounds__(VEC_BLOCK) k_spmv_sell(const int32_t *__restrict__ sl_w, const int64_t *__restrict__ valptr,
                                                         const int64_t *__restrict__ idxptr, const double *__restrict__ val,
                                                         const int32_t *__restrict__ idx, const double *__restrict__ x,
                                                         double *__restrict__ y, const double *__restrict__ w,
                                                         int64_t row0, int64_t row1, int64_t n_slices, int32_t n_cols,
                                                         double *partial, unsigned int *ticket, double *sc, int32_t *fl,
                                                         int stage, int fused, int check_done, const P2PDev *pd, int epoch,
                                                         int halo_epoch) {
  if (check_done && fl[F_DONE]) return;
  if (pd && halo_epoch > 0) {
    // ghost entries of x are written by the neighbours' k_halo_push over NVLink: wait for this epoch's flags
    if (threadIdx.x == 0) {
      if (pd->has_lo) p2p_wait(pd->hflag_self + 0, halo_epoch, pd->err);
      if (pd->has_hi) p2p_wait(pd->hflag_self + 1, halo_epoch, pd->err);
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
  for (int i = 0; i < (NDOT > 0 ? NDOT : 1); ++i) acc[i] = 0.0;
  for (int64_t s = warp0; s < n_slices; s += nwarps) {
    const int32_t wenc = sl_w[s];
    const int32_t W = wenc & 0x7fffffff;
    const double2 *vp = reinterpret_cast<const double2 *>(val + valptr[s]) + lane;
    const int64_t r0 = row0 + s * 64 + 2 * lane;
    double a0 = 0.0, a1 = 0.0;
    if (wenc < 0) {
      const int32_t *op = idx + idxptr[s];
      const int32_t rr = (int32_t)r0;
#pragma unroll 4
      for (int32_t j = 0; j < W; ++j) {
        const int32_t off = __ldg(op + j);
        const double2 v = __ldcs(vp + (size_t)j * 32);
        int32_t c0 = min(max(rr + off, 0), n_cols - 1);
        int32_t c1 = min(max(rr + 1 + off, 0), n_cols - 1);
        a0 += v.x * __ldg(x + c0);
        a1 += v.y * __ldg(x + c1);
      }
    } else {
      const int2 *cp = reinterpret_cast<const int2 *>(idx + idxptr[s]) + lane;
#pragma unroll 4
      for (int32_t j = 0; j < W; ++j) {
        const int2 c = __ldcs(cp + (size_t)j * 32);
        const double2 v = __ldcs(vp + (size_t)j * 32);
        a0 += v.x * __ldg(x + c.x);
        a1 += v.y * __ldg(x + c.y);
      }
    }
    if (r0 < row1) {
      y[r0] = a0;
      if (NDOT >= 1) acc[0] += w[r0] * a0;
      if (NDOT >= 2) acc[1] += a0 * a0;
    }
    if (r0 + 1 < row1) {
      y[r0 + 1] = a1;
      if (NDOT >= 1) acc[0] += w[r0 + 1] * a1;
      if (NDOT >= 2) acc[1] += a1 * a1;
    }
  }
  if constexpr (NDOT > 0) reduce_finalize<NDOT>(acc, partial, ticket, sc, fl, stage, fused, pd, epoch);
}

// ---- CG vector kernels -----------------------------------------------------------------------
// r = b - q ; z = minv r ; p = z ; sums (r.z, r.r, b.b)
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_init(const double *__restrict__ b, const double *__restrict__ q,
                                                       const double *__restrict__ minv, double *__restrict__ r,
                                                       double *__restrict__ p, int64_t i0, int64_t i1,
                                                       double *partial, unsigned int *ticket, double *sc, int32_t *fl, int stage, int fused, const P2PDev *pd, int epoch) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double bi = b[i];
    double ri = bi - q[i];
    double zi = minv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    acc[0] += ri * zi; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  reduce_finalize<3>(acc, partial, ticket, sc, fl, stage, fused, pd, epoch);
}
// x += alpha p ; r -= alpha q ; sums (r.(minv r), r.r)
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_update(const double *__restrict__ p, const double *__restrict__ q,
                                                         const double *__restrict__ minv, double *__restrict__ x,
                                                         double *__restrict__ r, int64_t i0, int64_t i1,
                                                         double *partial, unsigned int *ticket, double *sc,
                                                         int32_t *fl, int fused, const P2PDev *pd, int epoch) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA];
  double acc[2] = {0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    double ri = r[i] - alpha * q[i];
    r[i] = ri;
    acc[0] += ri * (minv[i] * ri);
    acc[1] += ri * ri;
  }
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_CG_UPDATE, fused, pd, epoch);
}
// p = minv r + beta p
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_p(const double *__restrict__ r, const double *__restrict__ minv,
                                                    double *__restrict__ p, int64_t i0, int64_t i1, const double *sc,
                                                    const int32_t *fl) {
  if (fl[F_DONE]) return;
  const double beta = sc[S_BETA];
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = minv[i] * r[i] + beta * p[i];
}

// ---- BiCGSTAB vector kernels --------------------------------------------------------------------
// r = b - q ; r0 = r ; p = 0 ; v = 0 ; sums (r0.r, r.r, b.b)
__global__ void __launch_bounds__(VEC_BLOCK) k_bi_init(const double *__restrict__ b, const double *__restrict__ q,
                                                       double *__restrict__ r, double *__restrict__ r0,
                                                       double *__restrict__ p, double *__restrict__ v, int64_t i0,
                                                       int64_t i1, double *partial, unsigned int *ticket, double *sc,
                                                       int32_t *fl, int fused, const P2PDev *pd, int epoch) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double bi = b[i];
    double ri = bi - q[i];
    r[i] = ri; r0[i] = ri; p[i] = 0.0; v[i] = 0.0;
    acc[0] += ri * ri; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  reduce_finalize<3>(acc, partial, ticket, sc, fl, ST_BI_INIT, fused, pd, epoch);
}
// p = r + beta (p - omega v) ; phat = minv p
__global__ void __launch_bounds__(VEC_BLOCK) k_bi_p(const double *__restrict__ r, const double *__restrict__ v,
                                                    const double *__restrict__ minv, double *__restrict__ p,
                                                    double *__restrict__ phat, int64_t i0, int64_t i1,
                                                    const double *sc, const int32_t *fl) {
  if (fl[F_DONE]) return;
  const double beta = sc[S_BETA], omega = sc[S_OMEGA];
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double pi = r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pi;
    phat[i] = minv[i] * pi;
  }
}
// s = r - alpha v ; shat = minv s ; sums (s.s)
__global__ void __launch_bounds__(VEC_BLOCK) k_bi_s(const double *__restrict__ r, const double *__restrict__ v,
                                                    const double *__restrict__ minv, double *__restrict__ s,
                                                    double *__restrict__ shat, int64_t i0, int64_t i1,
                                                    double *partial, unsigned int *ticket, double *sc, int32_t *fl, int fused, const P2PDev *pd, int epoch) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA];
  double acc[1] = {0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double si = r[i] - alpha * v[i];
    s[i] = si;
    shat[i] = minv[i] * si;
    acc[0] += si * si;
  }
  reduce_finalize<1>(acc, partial, ticket, sc, fl, ST_BI_S, fused, pd, epoch);
}
// x += alpha phat + omega shat ; r = s - omega t ; sums (r0.r, r.r)
__global__ void __launch_bounds__(VEC_BLOCK) k_bi_x(const double *__restrict__ phat, const double *__restrict__ shat,
                                                    const double *__restrict__ s, const double *__restrict__ t,
                                                    const double *__restrict__ r0, double *__restrict__ x,
                                                    double *__restrict__ r, int64_t i0, int64_t i1, double *partial,
                                                    unsigned int *ticket, double *sc, int32_t *fl, int fused, const P2PDev *pd, int epoch) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA], omega = sc[S_OMEGA];
  double acc[2] = {0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * phat[i] + omega * shat[i];
    double ri = s[i] - omega * t[i];
    r[i] = ri;
    acc[0] += r0[i] * ri;
    acc[1] += ri * ri;
  }
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_BI_X, fused, pd, epoch);
}

// minv over the owned rows [row0, row0+n): 1/diag read from the sliced-ELL values (solver.py:1095)
__global__ void k_jacobi_inv(const double *__restrict__ sell_val, const int64_t *__restrict__ valptr,
                             const int32_t *__restrict__ diag, int64_t row0, int64_t n, int jacobi,
                             double *__restrict__ minv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  minv[row0 + i] = (jacobi && diag[i] >= 0) ? 1.0 / sell_val[valptr[i >> 6] + diag[i]] : 1.0;
}

// ---- host side ------------------------------------------------------------------------------------
int krylov_alloc(apdx_plan *pl) {
  KrylovWork &k = pl->kw;
  if (k.r.p) return APDX_OK;
  const int64_t n = pl->n_free;
  APDX_CHECK(k.r.alloc(n));
  if (pl->p2p.enabled) k.p.adopt(pl->p2p.vec_base + 0 * pl->p2p.stride, n); else APDX_CHECK(k.p.alloc(n));
  APDX_CHECK(k.q.alloc(n));
  APDX_CHECK(k.minv.alloc(n));
  APDX_CHECK(k.partial.alloc(4 * (size_t)(VEC_GRID > 148 * 32 ? VEC_GRID : 148 * 32)));
  APDX_CHECK(k.scal.alloc(S_COUNT));
  APDX_CHECK(k.ticket.alloc(2));
  APDX_CHECK(k.flags.alloc(F_COUNT));
  APDX_CUDA(cudaMemsetAsync(k.ticket.p, 0, 2 * sizeof(unsigned int), pl->stream));
  // halo entries outside the owned range must read as finite numbers
  APDX_CUDA(cudaMemsetAsync(k.p.p, 0, n * sizeof(double), pl->stream));
  return APDX_OK;
}
static int krylov_alloc_bicgstab(apdx_plan *pl) {
  KrylovWork &k = pl->kw;
  if (k.s.p) return APDX_OK;
  const int64_t n = pl->n_free;
  APDX_CHECK(k.s.alloc(n));
  APDX_CHECK(k.t.alloc(n));
  if (pl->p2p.enabled) k.phat.adopt(pl->p2p.vec_base + 1 * pl->p2p.stride, n); else APDX_CHECK(k.phat.alloc(n));
  if (pl->p2p.enabled) k.shat.adopt(pl->p2p.vec_base + 2 * pl->p2p.stride, n); else APDX_CHECK(k.shat.alloc(n));
  APDX_CHECK(k.r0.alloc(n));
  APDX_CUDA(cudaMemsetAsync(k.phat.p, 0, n * sizeof(double), pl->stream));
  APDX_CUDA(cudaMemsetAsync(k.shat.p, 0, n * sizeof(double), pl->stream));
  return APDX_OK;
}

// ---- launch helpers ---------------------------------------------------------------------------------------
// comm modes: single GPU (dot products finalised inside the kernel), NCCL (ncclSend/Recv halo + ncclAllReduce
// + one-thread scalar kernel), P2P (peer stores over NVLink: k_halo_push + mailbox all-reduce, no NCCL in the loop)
struct Comm {
  bool multi, p2p;
  const P2PDev *pd;
  int fused;
};
static Comm comm_of(apdx_plan *pl) {
  Comm c;
  c.multi = comm_active();
  c.p2p = c.multi && pl->p2p.enabled;
  c.pd = c.p2p ? pl->p2p.dev : nullptr;
  c.fused = c.multi ? 0 : 1;
  return c;
}

template <int NDOT>
static int launch_spmv(apdx_plan *pl, const double *x, double *y, const double *w, int stage, int check_done,
                       int halo_epoch = 0) {
  KrylovWork &k = pl->kw;
  Sell &S = pl->sell;
  const Comm c = comm_of(pl);
  const int64_t warps_per_block = VEC_BLOCK / 32;
  int64_t nb = (S.n_slices + warps_per_block - 1) / warps_per_block;
  const int64_t cap = 148ll * 32;
  unsigned grid = (unsigned)(nb < cap ? (nb > 0 ? nb : 1) : cap);
  const int epoch = (NDOT > 0 && c.p2p && stage >= 0) ? ++pl->p2p.red_epoch : 0;
  k_spmv_sell<NDOT><<<grid, VEC_BLOCK, 0, pl->stream>>>(S.sl_w.p, S.valptr.p, S.idxptr.p, S.val.p, S.idx.p, x, y, w,
                                                        pl->f0, pl->f1, S.n_slices, (int32_t)pl->n_free, k.partial.p,
                                                        k.ticket.p, k.scal.p, k.flags.p, stage, c.fused, check_done,
                                                        (NDOT > 0 && stage >= 0) ? c.pd : (halo_epoch ? c.pd : nullptr),
                                                        epoch, halo_epoch);
  pl->stats.spmv_launches += 1;
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

// make the ghost entries of a Krylov vector current; returns the halo epoch the consuming SpMV must wait for (P2P)
static int exchange_halo(apdx_plan *pl, double *v, int *halo_epoch) {
  *halo_epoch = 0;
  const Comm c = comm_of(pl);
  if (!c.multi) return APDX_OK;
  if (c.p2p && p2p_is_heap_vector(pl, v)) {
    P2P &P = pl->p2p;
    const int e = ++P.halo_epoch;
    const int64_t off = v - P.vec_base;  // same offset inside every rank's heap
    double *lo = pl->rank_lo >= 0 ? P.peer_vec[0] + off + P.peer_lo_f1 : nullptr;
    double *hi = pl->rank_hi >= 0 ? P.peer_vec[1] + off : nullptr;
    k_halo_push<<<32, VEC_BLOCK, 0, pl->stream>>>(v, pl->f0, pl->f1, pl->send_lo, pl->send_hi, lo, hi, P.peer_hflag[0],
                                                  P.peer_hflag[1], e, pl->kw.ticket.p + 1, pl->kw.flags.p);
    pl->stats.kernel_launches += 1;
    *halo_epoch = e;
    return APDX_OK;
  }
  return comm_halo_exchange(pl, v, pl->stream);
}

// after a dot-product kernel: multi-GPU reduction of the pending sums + scalar stage
static int finish_stage(apdx_plan *pl, int stage, int nv) {
  const Comm c = comm_of(pl);
  if (!c.multi) return APDX_OK;
  KrylovWork &k = pl->kw;
  if (c.p2p) {
    k_apply_stage_p2p<<<1, 32, 0, pl->stream>>>(stage, nv, k.scal.p, k.flags.p, c.pd, pl->p2p.red_epoch);
  } else {
    APDX_CHECK(comm_allreduce_sum(k.scal.p + S_PEND, nv, pl->stream));
    k_apply_stage<<<1, 1, 0, pl->stream>>>(stage, k.scal.p, k.flags.p);
  }
  pl->stats.kernel_launches += 1;
  return APDX_OK;
}
// epoch argument of a vector kernel that ends in a reduction
static int next_red_epoch(apdx_plan *pl) { return comm_of(pl).p2p ? ++pl->p2p.red_epoch : 0; }

int spmv_reduced(apdx_plan *pl, const double *x, double *y) {
  APDX_REQUIRE(pl->have_sell_values, APDX_ERR_STATE, "no assembled tangent: call apdx_assemble first");
  APDX_CHECK(krylov_alloc(pl));
  if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, const_cast<double *>(x), pl->stream));
  return launch_spmv<0>(pl, x, y, nullptr, 0, 0);
}

int time_spmv(apdx_plan *pl, int reps, double *ms_avg) {
  APDX_REQUIRE(pl->have_sell_values, APDX_ERR_STATE, "no assembled tangent: call apdx_assemble first");
  APDX_CHECK(krylov_alloc(pl));
  KrylovWork &k = pl->kw;
  cudaStream_t s = pl->stream;
  // a smooth non-trivial input vector: p = 1/diag on the owned rows
  k_jacobi_inv<<<(unsigned)((pl->sell.n_rows + 255) / 256), 256, 0, s>>>(pl->sell.val.p, pl->sell.valptr.p, pl->sell.diag.p,
                                                                      pl->f0, pl->sell.n_rows, 1, k.p.p);
  for (int i = 0; i < 3; ++i) APDX_CHECK(launch_spmv<1>(pl, k.p.p, k.q.p, k.p.p, -1, 0));
  APDX_CUDA(cudaEventRecord(pl->ev[0], s));
  for (int i = 0; i < reps; ++i) APDX_CHECK(launch_spmv<1>(pl, k.p.p, k.q.p, k.p.p, -1, 0));
  APDX_CUDA(cudaEventRecord(pl->ev[1], s));
  APDX_CUDA(cudaEventSynchronize(pl->ev[1]));
  float ms = 0.f;
  APDX_CUDA(cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[1]));
  *ms_avg = (double)ms / reps;
  return APDX_OK;
}

int krylov_solve(apdx_plan *pl, const apdx_krylov_opts *o, const double *rhs, double *x, int32_t *iters,
                 double *relres) {
  APDX_REQUIRE(pl->have_sell_values, APDX_ERR_STATE, "no assembled tangent: call apdx_assemble first");
  APDX_REQUIRE(o->method == APDX_KRYLOV_CG || o->method == APDX_KRYLOV_BICGSTAB, APDX_ERR_UNSUPPORTED,
               "Krylov method %d not supported (cg, bicgstab)", o->method);
  APDX_CHECK(krylov_alloc(pl));
  const bool bi = o->method == APDX_KRYLOV_BICGSTAB;
  if (bi) APDX_CHECK(krylov_alloc_bicgstab(pl));
  KrylovWork &k = pl->kw;
  cudaStream_t s = pl->stream;
  const Comm c = comm_of(pl);
  const int64_t i0 = pl->f0, i1 = pl->f1, n = pl->n_free;
  const int fused = c.fused;
  const P2PDev *pd = c.pd;
  const int maxiter = o->maxiter > 0 ? o->maxiter : 10 * (int)(n < 100000 ? n : 100000);
  const int chunk = o->check_every > 0 ? o->check_every : 32;

  k_jacobi_inv<<<(unsigned)((pl->sell.n_rows + 255) / 256), 256, 0, s>>>(pl->sell.val.p, pl->sell.valptr.p, pl->sell.diag.p,
                                                                      pl->f0, pl->sell.n_rows, o->jacobi, k.minv.p);
  double sc_h[S_COUNT] = {0};
  sc_h[S_TOL2] = o->rtol * o->rtol;
  sc_h[S_SS] = o->atol * o->atol;
  int32_t fl_h[F_COUNT] = {0, 0, 0, maxiter};
  APDX_CUDA(cudaMemcpyAsync(k.scal.p, sc_h, sizeof(sc_h), cudaMemcpyHostToDevice, s));
  APDX_CUDA(cudaMemcpyAsync(k.flags.p, fl_h, sizeof(fl_h), cudaMemcpyHostToDevice, s));
  pl->stats.kernel_launches += 1;

  // q = A x0 (x is the caller's buffer: its ghost entries travel with NCCL)
  if (c.multi) APDX_CHECK(comm_halo_exchange(pl, x, s));
  APDX_CHECK(launch_spmv<0>(pl, x, bi ? k.t.p : k.q.p, nullptr, 0, 0));
  if (!bi) {
    k_cg_init<<<VEC_GRID, VEC_BLOCK, 0, s>>>(rhs, k.q.p, k.minv.p, k.r.p, k.p.p, i0, i1, k.partial.p, k.ticket.p,
                                             k.scal.p, k.flags.p, ST_CG_INIT, fused, pd, next_red_epoch(pl));
    pl->stats.kernel_launches += 1;
    APDX_CHECK(finish_stage(pl, ST_CG_INIT, 3));
  } else {
    k_bi_init<<<VEC_GRID, VEC_BLOCK, 0, s>>>(rhs, k.t.p, k.r.p, k.r0.p, k.p.p, k.q.p, i0, i1, k.partial.p,
                                             k.ticket.p, k.scal.p, k.flags.p, fused, pd, next_red_epoch(pl));
    pl->stats.kernel_launches += 1;
    APDX_CHECK(finish_stage(pl, ST_BI_INIT, 3));
  }

  int32_t *fl_pin = reinterpret_cast<int32_t *>(pl->pinned);
  double *sc_pin = pl->pinned + 8;
  int launched = 0, he = 0;
  while (true) {
    APDX_CUDA(cudaMemcpyAsync(fl_pin, k.flags.p, sizeof(int32_t) * F_COUNT, cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaMemcpyAsync(sc_pin, k.scal.p, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaStreamSynchronize(s));
    if (fl_pin[F_DONE] || launched >= maxiter) break;
    int todo = maxiter - launched < chunk ? maxiter - launched : chunk;
    for (int it = 0; it < todo; ++it) {
      if (!bi) {
        APDX_CHECK(exchange_halo(pl, k.p.p, &he));
        APDX_CHECK(launch_spmv<1>(pl, k.p.p, k.q.p, k.p.p, ST_CG_PQ, 1, he));
        APDX_CHECK(finish_stage(pl, ST_CG_PQ, 1));
        k_cg_update<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.p.p, k.q.p, k.minv.p, x, k.r.p, i0, i1, k.partial.p,
                                                   k.ticket.p, k.scal.p, k.flags.p, fused, pd, next_red_epoch(pl));
        APDX_CHECK(finish_stage(pl, ST_CG_UPDATE, 2));
        k_cg_p<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.minv.p, k.p.p, i0, i1, k.scal.p, k.flags.p);
        pl->stats.kernel_launches += 2;
      } else {
        k_bi_p<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.q.p, k.minv.p, k.p.p, k.phat.p, i0, i1, k.scal.p, k.flags.p);
        APDX_CHECK(exchange_halo(pl, k.phat.p, &he));
        APDX_CHECK(launch_spmv<1>(pl, k.phat.p, k.q.p, k.r0.p, ST_BI_R0V, 1, he));
        APDX_CHECK(finish_stage(pl, ST_BI_R0V, 1));
        k_bi_s<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.q.p, k.minv.p, k.s.p, k.shat.p, i0, i1, k.partial.p,
                                              k.ticket.p, k.scal.p, k.flags.p, fused, pd, next_red_epoch(pl));
        APDX_CHECK(finish_stage(pl, ST_BI_S, 1));
        APDX_CHECK(exchange_halo(pl, k.shat.p, &he));
        APDX_CHECK(launch_spmv<2>(pl, k.shat.p, k.t.p, k.s.p, ST_BI_T, 1, he));
        APDX_CHECK(finish_stage(pl, ST_BI_T, 2));
        k_bi_x<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.phat.p, k.shat.p, k.s.p, k.t.p, k.r0.p, x, k.r.p, i0, i1,
                                              k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused, pd, next_red_epoch(pl));
        APDX_CHECK(finish_stage(pl, ST_BI_X, 2));
        pl->stats.kernel_launches += 3;
      }
    }
    launched += todo;
  }
  APDX_CUDA(cudaGetLastError());
  if (c.p2p) {
    int err = 0;
    APDX_CUDA(cudaMemcpy(&err, pl->p2p.err_d, sizeof(int), cudaMemcpyDeviceToHost));
    APDX_REQUIRE(err == 0, APDX_ERR_NCCL, "peer-to-peer wait timed out (a rank stopped posting halos / reductions)");
  }
  if (iters) *iters = fl_pin[F_ITERS];
  if (relres) *relres = sc_pin[S_BB] > 0 ? sqrt(sc_pin[S_RR] / sc_pin[S_BB]) : sqrt(sc_pin[S_RR]);
  pl->stats.krylov_iters += fl_pin[F_ITERS];
  if (fl_pin[F_BREAKDOWN]) {
    set_error("Krylov breakdown (NaN or zero inner product) after %d iterations", fl_pin[F_ITERS]);
  }
  return APDX_OK;
}

}  // namespace apdx
