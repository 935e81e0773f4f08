// Device-side pieces shared by the Krylov loops (krylov.cu: Jacobi-PCG / BiCGSTAB; multigrid.cu: multigrid-PCG):
// scalar slots and stages of the recurrences, the fixed-order dot-product reduction.
#pragma once
#include "common.cuh"

namespace apdx {

// device scalar slots
enum {
  S_RZ = 0, S_PQ, S_ALPHA, S_BETA, S_RR, S_TOL2, S_BB, S_RHO, S_OMEGA, S_TS, S_TT, S_R0V, S_SS,
  S_PEND = 16,  // pending (locally reduced) sums of the running stage, up to 4
  S_COUNT = 24
};
enum { F_DONE = 0, F_ITERS = 1, F_BREAKDOWN = 2, F_MAXITER = 3, F_FINAL = 4, F_COUNT = 5 };   // F_FINAL: see k_cg_p
// stages of the scalar recurrences
enum { ST_CG_INIT = 0, ST_CG_PQ, ST_CG_UPDATE, ST_BI_INIT, ST_BI_R0V, ST_BI_S, ST_BI_T, ST_BI_X,
       ST_MG_INIT, ST_MG_RZ0, ST_MG_RR, ST_MG_RZ, ST_NONE };

constexpr int VEC_BLOCK = 256;

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
    // ---- multigrid-preconditioned CG (multigrid.cu): z = V-cycle(r) sits between the r.r and the r.z reductions ----
    case ST_MG_INIT:  // pend = (r.r, b.b)
      sc[S_RR] = pd[0]; sc[S_BB] = pd[1];
      {
        double t = sc[S_TOL2] * pd[1];
        double a2 = sc[S_SS];
        sc[S_TOL2] = t > a2 ? t : a2;
      }
      if (!(sc[S_RR] > sc[S_TOL2])) fl[F_DONE] = 1;
      break;
    case ST_MG_RZ0:  // pend = (r.z) of the initial residual
      sc[S_RZ] = pd[0];
      break;
    case ST_MG_RR:  // pend = (r.r) after the update
      sc[S_RR] = pd[0];
      fl[F_ITERS] += 1;
      if (!(pd[0] > sc[S_TOL2]) || fl[F_ITERS] >= fl[F_MAXITER]) fl[F_DONE] = 1;
      if (pd[0] != pd[0]) { fl[F_DONE] = 1; fl[F_BREAKDOWN] = 1; }
      break;
    case ST_MG_RZ:  // pend = (r.z) of the new residual
      sc[S_BETA] = pd[0] / sc[S_RZ];
      sc[S_RZ] = pd[0];
      break;
    default:  // ST_NONE: the sums stay in sc[S_PEND ..] for the host
      break;
  }
}

// Block-level reduction of NV running sums, then cross-block reduction by the last block.
// fused != 0: the last block also applies the scalar stage (single-GPU path).
template <int NV>
__device__ __forceinline__ void reduce_finalize(double (&v)[NV], double *partial, unsigned int *ticket,
                                                double *sc, int32_t *fl, int stage, int fused) {
  const unsigned nblk = gridDim.x;
  __shared__ double sh[NV][32];   // up to 32 warps per block
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
      partial[(size_t)i * nblk + blockIdx.x] = x;
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
    for (unsigned b = threadIdx.x; b < nblk; b += blockDim.x) x += __ldcg(&partial[(size_t)i * nblk + b]);
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
}

unsigned vec_grid();   // grid of the streaming vector kernels (krylov.cu)
// y = A x on the plan's sliced-ELL matrix, optionally with fused dot products (w.y [, y.y]) reduced into sc[S_PEND ..]
// and the scalar stage applied by the last block; check_done: no-op once the done flag is set (krylov.cu)
int spmv_launch(apdx_plan *pl, const double *x, double *y, const double *w, int ndot, int stage, int check_done);
// after a kernel that left locally reduced sums in sc[S_PEND ..]: multi-GPU all-reduce of the nv sums (peer-memory
// mailboxes or ncclAllReduce) followed by the scalar stage; no-op on one GPU, where the producing kernel applied it
int krylov_finish_stage(apdx_plan *pl, int stage, int nv);

}  // namespace apdx
