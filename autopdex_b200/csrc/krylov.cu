// Jacobi-preconditioned CG / BiCGSTAB on the reduced CSR system, all scalars on device.
//
// Device analogue of solver.linear_solve_jax (solver.py:1093-1126: M = 1/diag,
// jax.scipy.sparse.linalg.cg / bicgstab) operating on the assembled reduced matrix that the
// reference hands to SciPy (solver.py:1207-1217).  The stopping rule is the one of
// jax.scipy.sparse.linalg: ||r||_2^2 <= max(rtol^2 ||b||^2, atol^2).  No positivity checks:
// the reference's tangents are negative definite on mesher quad meshes.
//
// SpMV: sliced-ELL with per-slice compressed column indices (sell.cu), warp-contiguous value loads, the dot
// product(s) that follow the SpMV fused into the same kernel (warp-shuffle + fixed-order block reduction).  Vector updates are
// fused axpy+dot kernels.  Dot products are reduced per block, then by the last block in a fixed
// order (deterministic for a given grid), then -- multi-GPU -- by ncclAllReduce.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

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
enum { ST_CG_INIT = 0, ST_CG_PQ, ST_CG_UPDATE, ST_BI_INIT, ST_BI_R0V, ST_BI_S, ST_BI_T, ST_BI_X, ST_CG2_INIT, ST_CG2_ITER };

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
    case ST_CG2_INIT:  // single-reduction CG (Chronopoulos-Gear): pend = (w.u, r.u, r.r, b.b)
      sc[S_RZ] = pd[1]; sc[S_RR] = pd[2]; sc[S_BB] = pd[3];
      {
        double t = sc[S_TOL2] * pd[3];
        double a2 = sc[S_SS];
        sc[S_TOL2] = t > a2 ? t : a2;
      }
      sc[S_BETA] = 0.0;
      sc[S_ALPHA] = pd[1] / pd[0];
      if (!(sc[S_RR] > sc[S_TOL2])) fl[F_DONE] = 1;
      break;
    case ST_CG2_ITER:  // pend = (w.u, r.u, r.r) after one more update of x and r
      {
        const double beta = pd[1] / sc[S_RZ];
        sc[S_ALPHA] = pd[1] / (pd[0] - beta * pd[1] / sc[S_ALPHA]);
        sc[S_BETA] = beta;
        sc[S_RZ] = pd[1];
        sc[S_RR] = pd[2];
        fl[F_ITERS] += 1;
        if (!(pd[2] > sc[S_TOL2]) || fl[F_ITERS] >= fl[F_MAXITER]) fl[F_DONE] = 1;
        if (pd[2] != pd[2]) { fl[F_DONE] = 1; fl[F_BREAKDOWN] = 1; }
      }
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
  if (fl[F_DONE] && stage != ST_CG_INIT && stage != ST_BI_INIT && stage != ST_CG2_INIT) return;
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

// mbox mode: the whole all-reduce of a dot-product stage in ONE small kernel -- post the local sums into every rank's
// mailbox (peer stores over NVLink), raise the flags, wait for every rank's post, sum in rank order (identical bits on
// all ranks), run the scalar recurrence.  The reduction counter lives on the device (every rank runs the same
// sequence of reductions), so the kernel has no per-launch argument and can be replayed inside the CUDA graph of a
// Krylov chunk.  Replaces ncclAllReduce (8-24 bytes, ~22 us) + k_apply_stage.  Mailboxes are double-buffered by the
// parity of the counter: a rank can be at most one reduction ahead of the slowest one.
__global__ void k_allreduce_mbox_apply(int stage, int nv, double *sc, int32_t *fl, const P2PDev *pd) {
  if (fl[F_DONE] && stage != ST_CG_INIT && stage != ST_BI_INIT && stage != ST_CG2_INIT) return;
  const int e = *pd->epoch_self + 1;
  const int slot = e & 1;
  const int r = threadIdx.x;
  if (r < pd->nranks) {
    double *mb = pd->mbox[r] + ((size_t)slot * P2P_MAX_RANKS + pd->rank) * 4;
    for (int i = 0; i < nv; ++i) mb[i] = sc[S_PEND + i];
    __threadfence_system();
    *(volatile int *)(pd->mflag[r] + slot * P2P_MAX_RANKS + pd->rank) = e;
    p2p_wait(pd->mflag[pd->rank] + slot * P2P_MAX_RANKS + r, e, pd->err);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const volatile double *mbs = pd->mbox[pd->rank] + (size_t)slot * P2P_MAX_RANKS * 4;
    for (int i = 0; i < nv; ++i) {
      double s = 0.0;
      for (int q = 0; q < pd->nranks; ++q) s += mbs[q * 4 + i];
      sc[S_PEND + i] = s;
    }
    apply_stage(stage, sc, fl);
    *pd->epoch_self = e;
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
                                                int epoch, int slot0 = 0, int blk0 = 0, int nblk_total = -1,
                                                int finalize = 1) {
  // blk0 / nblk_total / finalize: one reduction may be fed by two launches (interior and boundary SpMV): the first
  // only deposits its block partials, the second sums the partials of both
  const unsigned nblk = nblk_total < 0 ? gridDim.x : (unsigned)nblk_total;
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
      partial[(size_t)i * nblk + blk0 + blockIdx.x] = x;
    }
    if (finalize) {
      __threadfence();
      unsigned int t = atomicAdd(ticket, 1u);
      last = (t == gridDim.x - 1);
    } else {
      last = false;
    }
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
      sc[S_PEND + slot0 + i] = y;
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

// Sliced-ELL SpMV (layout: sell.cu).  One warp per slice of 64 rows; lane owns the local rows lane and lane + 32, so
// each load instruction of the warp covers 256 contiguous bytes.  Offset mode: column = row + off[j] (offsets held one
// per lane, x gathers coalesced); explicit mode: one int32 column per stored entry.
// n_cols: length of x (clamp target of the padded offsets, whose values are exact zeros).
struct SliceRange {
  int64_t a0, a1, b0, b1;
  int blk0, nblk_total, finalize;
};

// Loads are issued in batches of up to SPMV_U slice columns before the first FMA.  The warp issues in order, so a
// loop that alternates load / FMA keeps only one 512-byte value row in flight per warp (measured: 62 % of the DRAM
// peak, long-scoreboard bound; batching: 96 % of the measured copy bandwidth).  A slice's stored columns are taken in
// chunks of 32 (lane u holds the offset of column u of the chunk, handed out by shuffles: no dependent broadcast load
// per column) and every chunk in equal batches of at most SPMV_U columns.
//
// SYM: lower columns of offset-mode slices are read from their transposed position (sell.cu): the mirror table
// (offset, posA, posB, split) is padded to full batches of MB = 7 or 8 entries, read with warp-uniform 16-byte loads
// (its cache lines are prefetched into the L1 while the stored columns stream), and the value of local row k is
// val[(k < split ? posA : posB) + k].  The value stream then uses the default L2 policy instead of evict-first: the
// partner slices re-read it from the L2 a few MB later.
//
// Work distribution: a persistent grid of exactly the resident blocks (SMs x blocks per SM), slices strided over all
// its warps.  At any time the slices in flight form one narrow window (~2400 slices) that moves through the matrix:
// the x entries and the mirrored values a slice needs were touched moments ago by its neighbours and are L2 hits, and
// every warp has ~100 slices over which the metadata prefetch and the block-level reduction are amortised.
// (A larger grid-stride grid keeps several distant fronts alive -- 30 % of the mirrored reads then missed the L2 --
// and one block per 8..32 consecutive slices pays the start-up latency chain per block: 0.83 / 0.69 / 0.64 ms.)
constexpr int SPMV_U = 9;

// Fast paths for slices flagged SELL_FAST (every column of every row inside x: no index clamps) whose column count is
// a multiple of the batch B: compile-time trip counts, no predicates, offsets / table entries read with warp-uniform
// loads from cache lines prefetched when the slice's header arrived.  Half the instructions of the generic path.
// XG: the x operands are fetched in groups of XG columns AFTER all B value loads of the batch have been issued (x
// gathers are L1/L2 hits, the values come from HBM): with XG < B the registers hold twice as many value loads in
// flight (B = 14: a whole slice of the symmetric P256 matrix in one round trip) at the price of short, cached
// latencies in series.
template <int B, int XG, bool SYM>
__device__ __forceinline__ void spmv_stored_fast(int32_t offl, const double *__restrict__ vpc, int32_t nb,
                                                 const double *__restrict__ xr0, const double *__restrict__ xr1, double &a0,
                                                 double &a1) {
  static_assert(XG >= 1 && XG <= B, "x groups are taken out of the batch");   // the last group may be shorter
  for (int32_t jb = 0; jb < nb; jb += B) {
    double va[B], vb[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const double *q = vpc + (size_t)(jb + u) * 64;
      va[u] = SYM ? __ldg(q) : __ldcs(q);
      vb[u] = SYM ? __ldg(q + 32) : __ldcs(q + 32);
    }
#pragma unroll
    for (int g = 0; g < B; g += XG) {
      int32_t off[XG];
      double xa[XG], xb[XG];
#pragma unroll
      for (int u = 0; u < XG; ++u)
        if (g + u < B) off[u] = __shfl_sync(0xffffffffu, offl, jb + g + u);   // lane j holds column j's offset
#pragma unroll
      for (int u = 0; u < XG; ++u)
        if (g + u < B) { xa[u] = __ldg(xr0 + off[u]); xb[u] = __ldg(xr1 + off[u]); }
#pragma unroll
      for (int u = 0; u < XG; ++u)
        if (g + u < B) { a0 += va[g + u] * xa[u]; a1 += vb[g + u] * xb[u]; }
      if (XG < B) asm volatile("" ::: "memory");   // keep the next group's gathers behind this group's FMAs
    }
  }
}
// PIPE variant (APDX_SPMV_PIPE=1): the values of a slice's first B columns were requested while the warp was still
// in the mirrored phase of its previous slice (pa / pb), so that one HBM round trip of every slice overlaps the L2
// round trips of the slice before.  Same column order as spmv_stored_fast<B, XG>: bit-identical sums.
template <int B, int XG, bool SYM>
__device__ __forceinline__ void spmv_stored_fast_pre(const double (&pa)[B], const double (&pb)[B], int32_t offl,
                                                     const double *__restrict__ vpc, int32_t nb,
                                                     const double *__restrict__ xr0, const double *__restrict__ xr1,
                                                     double &a0, double &a1) {
  static_assert(XG >= 1 && XG <= B, "x groups are taken out of the batch");
#pragma unroll
  for (int g = 0; g < B; g += XG) {
    int32_t off[XG];
    double xa[XG], xb[XG];
#pragma unroll
    for (int u = 0; u < XG; ++u)
      if (g + u < B) off[u] = __shfl_sync(0xffffffffu, offl, g + u);
#pragma unroll
    for (int u = 0; u < XG; ++u)
      if (g + u < B) { xa[u] = __ldg(xr0 + off[u]); xb[u] = __ldg(xr1 + off[u]); }
#pragma unroll
    for (int u = 0; u < XG; ++u)
      if (g + u < B) { a0 += pa[g + u] * xa[u]; a1 += pb[g + u] * xb[u]; }
  }
  for (int32_t jb = B; jb < nb; jb += B) {
    double va[B], vb[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const double *q = vpc + (size_t)(jb + u) * 64;
      va[u] = SYM ? __ldg(q) : __ldcs(q);
      vb[u] = SYM ? __ldg(q + 32) : __ldcs(q + 32);
    }
#pragma unroll
    for (int g = 0; g < B; g += XG) {
      int32_t off[XG];
      double xa[XG], xb[XG];
#pragma unroll
      for (int u = 0; u < XG; ++u)
        if (g + u < B) off[u] = __shfl_sync(0xffffffffu, offl, jb + g + u);
#pragma unroll
      for (int u = 0; u < XG; ++u)
        if (g + u < B) { xa[u] = __ldg(xr0 + off[u]); xb[u] = __ldg(xr1 + off[u]); }
#pragma unroll
      for (int u = 0; u < XG; ++u)
        if (g + u < B) { a0 += va[g + u] * xa[u]; a1 += vb[g + u] * xb[u]; }
    }
  }
}
// chunks of 32 columns that the fast paths take in batches of 7 (not 9, not 8)
__device__ __forceinline__ bool spmv_takes_7(int32_t nb) { return nb > 0 && nb % 7 == 0 && nb % 9 != 0 && nb % 8 != 0; }

template <int MB>
__device__ __forceinline__ void spmv_mirrored_fast(const int4 *__restrict__ tab, int32_t M, const double *__restrict__ val,
                                                   const double *__restrict__ xr0, const double *__restrict__ xr1, int lane,
                                                   double &a0, double &a1) {
  const int k0 = lane, k1 = lane + 32;
  for (int32_t jb = 0; jb < M; jb += MB) {
    int4 t[MB];
#pragma unroll
    for (int u = 0; u < MB; ++u) t[u] = __ldg(tab + jb + u);
    double m0[MB], m1[MB], xa[MB], xb[MB];
#pragma unroll
    for (int u = 0; u < MB; ++u) {
      m0[u] = __ldg(val + ((k0 < t[u].w ? t[u].y : t[u].z) + k0));
      m1[u] = __ldg(val + ((k1 < t[u].w ? t[u].y : t[u].z) + k1));
      xa[u] = __ldg(xr0 + t[u].x);
      xb[u] = __ldg(xr1 + t[u].x);
    }
#pragma unroll
    for (int u = 0; u < MB; ++u) { a0 += m0[u] * xa[u]; a1 += m1[u] * xb[u]; }
  }
}

template <int MB>
__device__ __forceinline__ void spmv_mirrored(const int4 *__restrict__ tab, int32_t M, const double *__restrict__ val,
                                              const double *__restrict__ x, int lane, int32_t rr0, int32_t rr1,
                                              int32_t n_cols, double &a0, double &a1) {
  const int k0 = lane, k1 = lane + 32;
  for (int32_t jb = 0; jb < M; jb += MB) {
    int4 t[MB];
#pragma unroll
    for (int u = 0; u < MB; ++u) t[u] = __ldg(tab + jb + u);   // same address in every lane: one broadcast transaction
    double m0[MB], m1[MB], xa[MB], xb[MB];
#pragma unroll
    for (int u = 0; u < MB; ++u) {
      m0[u] = __ldg(val + ((k0 < t[u].w ? t[u].y : t[u].z) + k0));
      m1[u] = __ldg(val + ((k1 < t[u].w ? t[u].y : t[u].z) + k1));
      xa[u] = __ldg(x + min(max(rr0 + t[u].x, 0), n_cols - 1));
      xb[u] = __ldg(x + min(max(rr1 + t[u].x, 0), n_cols - 1));
    }
#pragma unroll
    for (int u = 0; u < MB; ++u) { a0 += m0[u] * xa[u]; a1 += m1[u] * xb[u]; }
  }
}

// Occupancy variants (OCC = blocks of 256 threads per SM the kernel is compiled for: 2 -> 128 registers, 3 -> 80,
// 4 -> 64; APDX_SPMV_BPS selects one).  The r01f capture shows 4 warps per scheduler that sit in long-scoreboard
// stalls (issue slots 34 % used): OCC 3 / 4 trade the depth of a warp's load batch for more warps.  The mirrored
// batches are then taken in sub-batches of SUB table entries, the generic paths in batches of fewer columns; the
// order of the FMAs into a0 / a1 is the same in every variant, so the results are bit-identical.
template <int N, bool CLAMP>
__device__ __forceinline__ void spmv_mirrored_chunk(const int4 *__restrict__ tab, const double *__restrict__ val,
                                                    const double *__restrict__ x, const double *__restrict__ xr0,
                                                    const double *__restrict__ xr1, int k0, int k1, int32_t rr0,
                                                    int32_t rr1, int32_t n_cols, double &a0, double &a1) {
  int4 t[N];
#pragma unroll
  for (int u = 0; u < N; ++u) t[u] = __ldg(tab + u);
  double m0[N], m1[N], xa[N], xb[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    m0[u] = __ldg(val + ((k0 < t[u].w ? t[u].y : t[u].z) + k0));
    m1[u] = __ldg(val + ((k1 < t[u].w ? t[u].y : t[u].z) + k1));
    if (CLAMP) {
      xa[u] = __ldg(x + min(max(rr0 + t[u].x, 0), n_cols - 1));
      xb[u] = __ldg(x + min(max(rr1 + t[u].x, 0), n_cols - 1));
    } else {
      xa[u] = __ldg(xr0 + t[u].x);
      xb[u] = __ldg(xr1 + t[u].x);
    }
  }
#pragma unroll
  for (int u = 0; u < N; ++u) { a0 += m0[u] * xa[u]; a1 += m1[u] * xb[u]; }
}
template <int MB, int SUB, int S0, bool CLAMP>
__device__ __forceinline__ void spmv_mirrored_steps(const int4 *__restrict__ tab, const double *__restrict__ val,
                                                    const double *__restrict__ x, const double *__restrict__ xr0,
                                                    const double *__restrict__ xr1, int k0, int k1, int32_t rr0,
                                                    int32_t rr1, int32_t n_cols, double &a0, double &a1) {
  if constexpr (S0 < MB) {
    constexpr int N = (MB - S0 < SUB) ? MB - S0 : SUB;
    spmv_mirrored_chunk<N, CLAMP>(tab + S0, val, x, xr0, xr1, k0, k1, rr0, rr1, n_cols, a0, a1);
    spmv_mirrored_steps<MB, SUB, S0 + N, CLAMP>(tab, val, x, xr0, xr1, k0, k1, rr0, rr1, n_cols, a0, a1);
  }
}
template <int MB, int SUB, bool CLAMP>
__device__ __forceinline__ void spmv_mirrored_sub(const int4 *__restrict__ tab, int32_t M, const double *__restrict__ val,
                                                  const double *__restrict__ x, const double *__restrict__ xr0,
                                                  const double *__restrict__ xr1, int lane, int32_t rr0, int32_t rr1,
                                                  int32_t n_cols, double &a0, double &a1) {
  for (int32_t jb = 0; jb < M; jb += MB)
    spmv_mirrored_steps<MB, SUB, 0, CLAMP>(tab + jb, val, x, xr0, xr1, lane, lane + 32, rr0, rr1, n_cols, a0, a1);
}

template <int NDOT, int NF, bool SYM, int OCCV>
__global__ void __launch_bounds__(VEC_BLOCK, OCCV == 5 ? 2 : OCCV)
    k_spmv_sell(const int32_t *__restrict__ sl_w, const int32_t *__restrict__ sl_m, const int64_t *__restrict__ valptr,
                const int64_t *__restrict__ idxptr, const double *__restrict__ val, const int32_t *__restrict__ idx,
                const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, int64_t row0,
                int64_t row1, int64_t n_slices, int32_t n_cols, double *partial, unsigned int *ticket, double *sc,
                int32_t *fl, int stage, int fused, int check_done, const P2PDev *pd, int epoch, int halo_epoch,
                SliceRange rg) {
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
  constexpr int WPB = VEC_BLOCK / 32;
  constexpr bool PIPE = OCCV == 5;          // 2 blocks per SM + first value batch of the next slice requested early
  constexpr int OCC = PIPE ? 2 : OCCV;
  // batch sizes of the occupancy variants (OCC == 2: the measured default)
  // (PIPE keeps 14 prefetched values alive through the stored phase: shallower batches there, like OCC 3)
  constexpr int U = PIPE ? 5 : (OCC == 2 ? SPMV_U : (OCC == 3 ? 5 : 3));          // generic paths: columns per batch
  constexpr bool DEEP = OCC == 2 && !PIPE;
  constexpr int XG9 = DEEP ? 9 : 3, XG8 = DEEP ? 8 : 4, XG7 = DEEP ? 7 : 4;       // (OCC 4 takes its own path)
  constexpr int MSUB = PIPE ? 4 : (OCC == 2 ? 8 : (OCC == 3 ? 4 : 2));   // mirrored table entries per sub-batch
  double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
  for (int i = 0; i < (NDOT > 0 ? NDOT : 1); ++i) acc[i] = 0.0;
  // rg: this launch covers slices [a0,a1) and [b0,b1) (interior launch: one range; boundary launch: the two ends)
  const int64_t n_mine = (rg.a1 - rg.a0) + (rg.b1 - rg.b0);
  const int64_t si_begin = (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
  const int64_t si_end = n_mine;
  const int64_t si_step = (int64_t)gridDim.x * WPB;
  // The per-slice metadata is a chain of dependent loads (header -> offsets -> x gathers): the header of the warp's
  // NEXT slice is requested before the current slice is processed, its first 32 offsets half-way through.
  struct Hdr { int32_t wenc, M; int64_t vp, ip; };
  auto load_hdr = [&](int64_t si, Hdr &h) {
    h.wenc = 0; h.M = 0; h.vp = 0; h.ip = 0;
    if (si < si_end) {
      const int64_t s = si < rg.a1 - rg.a0 ? rg.a0 + si : rg.b0 + (si - (rg.a1 - rg.a0));
      h.wenc = sl_w[s];
      h.M = sl_m[s];
      h.vp = valptr[s];
      h.ip = idxptr[s];
    }
  };
  auto load_offsets = [&](const Hdr &h) -> int32_t {
    const int32_t W = h.wenc & 0x7fffffff;
    if (h.wenc >= 0 || W == 0) return 0;
    // pull the slice's offsets and mirror table into the L1: the fast paths read them with warp-uniform loads
    const int32_t rec_ints = ((W + 3) & ~3) + 4 * (h.M & SELL_MMASK);
    if (lane * 32 < rec_ints) asm volatile("prefetch.global.L1 [%0];" ::"l"(idx + h.ip + lane * 32));
    return __ldg(idx + h.ip + min(lane, min(32, W) - 1));
  };
  Hdr cur;
  load_hdr(si_begin, cur);
  int32_t offl0 = load_offsets(cur);
  double pva[7], pvb[7];                    // PIPE: the first seven value columns of the coming slice
  auto prefetch_values = [&](const Hdr &h) {
    const bool take = (h.M & SELL_FAST) != 0 && spmv_takes_7(min(32, h.wenc & 0x7fffffff));   // warp-uniform
    if (take) {
      const double *q = val + h.vp + lane;
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        pva[u] = SYM ? __ldg(q + (size_t)u * 64) : __ldcs(q + (size_t)u * 64);
        pvb[u] = SYM ? __ldg(q + (size_t)u * 64 + 32) : __ldcs(q + (size_t)u * 64 + 32);
      }
    }
  };
  if constexpr (PIPE) {
#pragma unroll
    for (int u = 0; u < 7; ++u) pva[u] = pvb[u] = 0.0;
    prefetch_values(cur);
  }
  for (int64_t si = si_begin; si < si_end; si += si_step) {
    Hdr nxt;
    load_hdr(si + si_step, nxt);
    const int64_t s = si < rg.a1 - rg.a0 ? rg.a0 + si : rg.b0 + (si - (rg.a1 - rg.a0));
    const int32_t wenc = cur.wenc;
    const int32_t W = wenc & 0x7fffffff;
    const int32_t M = SYM ? (cur.M & SELL_MMASK) : 0;
    const bool fast = (cur.M & SELL_FAST) != 0;
    const double *vp = val + cur.vp + lane;
    // rows of the slice: one field component of 64 consecutive nodes (sell.cu); lane owns local rows k = lane and
    // lane + 32, so that every x gather, value load and y store of the warp covers one contiguous run of 32 entries
    // (half the L1 tag traffic of an interleaved (2 lane, 2 lane + 1) ownership, which bounded the kernel)
    const int64_t r0 = row0 + (s / NF) * (int64_t)64 * NF + (s % NF) + (int64_t)NF * lane;
    const int64_t r1 = r0 + (int64_t)32 * NF;
    const int32_t rr0 = (int32_t)r0, rr1 = (int32_t)r1;
    const double *xr0 = x + r0, *xr1 = x + r1;
    const int32_t *ip = idx + cur.ip;
    const int4 *tab = reinterpret_cast<const int4 *>(ip + ((W + 3) & ~3));
    if (SYM && lane * 8 < M)   // one lane per 128-byte line of the mirror table
      asm volatile("prefetch.global.L1 [%0];" ::"l"(tab + lane * 8));
    double a0 = 0.0, a1 = 0.0;
    for (int32_t jc = 0; jc < W; jc += 32) {
      const int32_t nb = min(32, W - jc);
      const int32_t nbt = (nb + U - 1) / U, bs = (nb + nbt - 1) / nbt;   // equal batches of <= U columns
      if (PIPE && jc == 0 && fast && spmv_takes_7(nb)) {
        spmv_stored_fast_pre<7, 4, SYM>(pva, pvb, offl0, vp, nb, xr0, xr1, a0, a1);
      } else if (OCC == 4 && fast && (nb % 7 == 0 || nb % 4 == 0 || nb % 3 == 0 || nb % 5 == 0)) {
        // 64 registers: value batches of at most 7 columns, the x operands of a 7-batch one column at a time
        const int32_t offl = jc == 0 ? offl0 : __ldg(ip + jc + min(lane, nb - 1));
        if (nb % 7 == 0) spmv_stored_fast<7, 1, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else if (nb % 4 == 0) spmv_stored_fast<4, 4, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else if (nb % 3 == 0) spmv_stored_fast<3, 3, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else spmv_stored_fast<5, 5, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
      } else if (OCC != 4 && fast && (nb % 9 == 0 || nb % 8 == 0 || nb % 7 == 0)) {
        const int32_t offl = jc == 0 ? offl0 : __ldg(ip + jc + min(lane, nb - 1));
        if (nb % 9 == 0) spmv_stored_fast<9, XG9, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else if (nb % 8 == 0) spmv_stored_fast<8, XG8, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else spmv_stored_fast<7, XG7, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
      } else if (wenc < 0) {
        const int32_t offl = jc == 0 ? offl0 : __ldg(ip + jc + min(lane, nb - 1));
        for (int32_t jb = 0; jb < nb; jb += bs) {
          double va[U], vb[U], xa[U], xb[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            va[u] = vb[u] = 0.0;
            if (u < bs && jb + u < nb) {
              const double *q = vp + (size_t)(jc + jb + u) * 64;
              va[u] = SYM ? __ldg(q) : __ldcs(q);
              vb[u] = SYM ? __ldg(q + 32) : __ldcs(q + 32);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int32_t off = __shfl_sync(0xffffffffu, offl, (jb + u) & 31);
            xa[u] = xb[u] = 0.0;
            if (u < bs && jb + u < nb) {
              xa[u] = __ldg(x + min(max(rr0 + off, 0), n_cols - 1));
              xb[u] = __ldg(x + min(max(rr1 + off, 0), n_cols - 1));
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) { a0 += va[u] * xa[u]; a1 += vb[u] * xb[u]; }
        }
      } else {
        const int32_t *cp = ip + lane;
        for (int32_t jb = 0; jb < nb; jb += bs) {
          int32_t ca[U], cb[U];
          double va[U], vb[U], xa[U], xb[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            ca[u] = cb[u] = 0;
            va[u] = vb[u] = 0.0;
            if (u < bs && jb + u < nb) {
              const size_t o = (size_t)(jc + jb + u) * 64;
              ca[u] = __ldcs(cp + o); cb[u] = __ldcs(cp + o + 32);
              va[u] = __ldcs(vp + o); vb[u] = __ldcs(vp + o + 32);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) { xa[u] = __ldg(x + ca[u]); xb[u] = __ldg(x + cb[u]); }
#pragma unroll
          for (int u = 0; u < U; ++u) { a0 += va[u] * xa[u]; a1 += vb[u] * xb[u]; }
        }
      }
    }
    // the next slice's header has arrived by now: request its offsets
    const int32_t offl_n = load_offsets(nxt);
    if constexpr (PIPE) prefetch_values(nxt);   // in flight during the mirrored phase below
    if (SYM && M > 0) {
      if constexpr (OCC == 2 && !PIPE) {
        if (fast) {
          if (cur.M & SELL_MB7) spmv_mirrored_fast<7>(tab, M, val, xr0, xr1, lane, a0, a1);
          else spmv_mirrored_fast<8>(tab, M, val, xr0, xr1, lane, a0, a1);
        } else {
          if (cur.M & SELL_MB7) spmv_mirrored<7>(tab, M, val, x, lane, rr0, rr1, n_cols, a0, a1);
          else spmv_mirrored<8>(tab, M, val, x, lane, rr0, rr1, n_cols, a0, a1);
        }
      } else {
        if (fast) {
          if (cur.M & SELL_MB7) spmv_mirrored_sub<7, MSUB, false>(tab, M, val, x, xr0, xr1, lane, rr0, rr1, n_cols, a0, a1);
          else spmv_mirrored_sub<8, MSUB, false>(tab, M, val, x, xr0, xr1, lane, rr0, rr1, n_cols, a0, a1);
        } else {
          if (cur.M & SELL_MB7) spmv_mirrored_sub<7, MSUB, true>(tab, M, val, x, xr0, xr1, lane, rr0, rr1, n_cols, a0, a1);
          else spmv_mirrored_sub<8, MSUB, true>(tab, M, val, x, xr0, xr1, lane, rr0, rr1, n_cols, a0, a1);
        }
      }
    }
    cur = nxt;
    offl0 = offl_n;
    if (r0 < row1) {
      y[r0] = a0;
      if (NDOT >= 1) acc[0] += w[r0] * a0;
      if (NDOT >= 2) acc[1] += a0 * a0;
    }
    if (r1 < row1) {
      y[r1] = a1;
      if (NDOT >= 1) acc[0] += w[r1] * a1;
      if (NDOT >= 2) acc[1] += a1 * a1;
    }
  }
  if constexpr (NDOT > 0)
    reduce_finalize<NDOT>(acc, partial, ticket, sc, fl, stage, fused, pd, epoch, 0, rg.blk0, rg.nblk_total, rg.finalize);
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
// The two vector kernels of a CG iteration.  x += alpha p rides with the p-update, where p is read anyway (10 vector
// streams per iteration instead of 11):
//   k_cg_update:  r -= alpha q ; sums (r.(minv r), r.r)                       reads q r minv, writes r
//   k_cg_p:       x += alpha p ; p = minv r + beta p                           reads p x r minv, writes x p
// The iteration that sets the done flag (in the scalar stage after k_cg_update) still owes x its update: k_cg_p applies
// it regardless of the flag and its last block then raises F_FINAL, which turns the k_cg_p launches of the remaining
// (no-op) iterations of a replayed chunk off.  VEC: 16-byte accesses over the even-aligned body of [i0, i1).
template <bool VEC>
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_update(const double *__restrict__ q, const double *__restrict__ minv,
                                                         double *__restrict__ r, int64_t i0, int64_t i1, double *partial,
                                                         unsigned int *ticket, double *sc, int32_t *fl, int fused,
                                                         const P2PDev *pd, int epoch) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA];
  double acc[2] = {0.0, 0.0};
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](int64_t i) {
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    acc[0] += ri * (minv[i] * ri);
    acc[1] += ri * ri;
  };
  if (VEC) {
    const int64_t a0 = (i0 + 1) & ~(int64_t)1;
    int64_t a1 = i1 & ~(int64_t)1;
    if (a1 < a0) a1 = a0;
    const double2 *q2 = reinterpret_cast<const double2 *>(q), *m2 = reinterpret_cast<const double2 *>(minv);
    double2 *r2 = reinterpret_cast<double2 *>(r);
    for (int64_t j = a0 / 2 + tid; j < a1 / 2; j += nth) {
      const double2 qv = q2[j], mv = m2[j];
      double2 rv = r2[j];
      rv.x -= alpha * qv.x;
      rv.y -= alpha * qv.y;
      r2[j] = rv;
      acc[0] += rv.x * (mv.x * rv.x) + rv.y * (mv.y * rv.y);
      acc[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if (tid == 0 && i0 < a0 && i0 < i1) one(i0);
    if (tid == 1 && a1 < i1 && a1 >= a0) one(a1);
  } else {
    for (int64_t i = i0 + tid; i < i1; i += nth) one(i);
  }
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_CG_UPDATE, fused, pd, epoch);
}
template <bool VEC>
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_p(const double *__restrict__ r, const double *__restrict__ minv,
                                                    double *__restrict__ p, double *__restrict__ x, int64_t i0, int64_t i1,
                                                    const double *sc, int32_t *fl, unsigned int *ticket) {
  if (fl[F_FINAL]) return;
  const bool done = fl[F_DONE] != 0;
  const double alpha = sc[S_ALPHA], beta = sc[S_BETA];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](int64_t i) {
    const double pi = p[i];
    x[i] += alpha * pi;
    if (!done) p[i] = minv[i] * r[i] + beta * pi;
  };
  if (VEC) {
    const int64_t a0 = (i0 + 1) & ~(int64_t)1;
    int64_t a1 = i1 & ~(int64_t)1;
    if (a1 < a0) a1 = a0;
    const double2 *r2 = reinterpret_cast<const double2 *>(r), *m2 = reinterpret_cast<const double2 *>(minv);
    double2 *p2 = reinterpret_cast<double2 *>(p), *x2 = reinterpret_cast<double2 *>(x);
    for (int64_t j = a0 / 2 + tid; j < a1 / 2; j += nth) {
      const double2 pv = p2[j];
      double2 xv = x2[j];
      xv.x += alpha * pv.x;
      xv.y += alpha * pv.y;
      x2[j] = xv;
      if (!done) {
        const double2 rv = r2[j], mv = m2[j];
        p2[j] = make_double2(mv.x * rv.x + beta * pv.x, mv.y * rv.y + beta * pv.y);
      }
    }
    if (tid == 0 && i0 < a0 && i0 < i1) one(i0);
    if (tid == 1 && a1 < i1 && a1 >= a0) one(a1);
  } else {
    for (int64_t i = i0 + tid; i < i1; i += nth) one(i);
  }
  if (done) {   // the last block to finish marks x as final (every block has read F_FINAL before taking its ticket)
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
      *ticket = 0u;
      fl[F_FINAL] = 1;
    }
  }
}

// ---- single-reduction CG (Chronopoulos & Gear) for the multi-GPU path: one vector kernel + one SpMV and ONE
// all-reduce of (w.u, r.u, r.r) per iteration instead of two.  Same Krylov space and, in exact arithmetic, the
// same iterates as the three-kernel CG above.
//   p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; sums (r.u, r.r) -> slots 1,2
__global__ void __launch_bounds__(VEC_BLOCK) k_cg2_vec(const double *__restrict__ w, const double *__restrict__ minv,
                                                       double *__restrict__ u, double *__restrict__ p, double *__restrict__ s,
                                                       double *__restrict__ x, double *__restrict__ r, int64_t i0, int64_t i1,
                                                       double *partial, unsigned int *ticket, double *sc, int32_t *fl) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA], beta = sc[S_BETA];
  double acc[2] = {0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double pi = u[i] + beta * p[i];
    const double si = w[i] + beta * s[i];
    p[i] = pi;
    s[i] = si;
    x[i] += alpha * pi;
    const double ri = r[i] - alpha * si;
    r[i] = ri;
    const double ui = minv[i] * ri;
    u[i] = ui;
    acc[0] += ri * ui;
    acc[1] += ri * ri;
  }
  reduce_finalize<2>(acc, partial, ticket, sc, fl, -1, 0, nullptr, 0, 1);
}
// r = b - q ; u = M^-1 r ; p = s = 0 ; sums (r.u, r.r, b.b) -> slots 1,2,3
__global__ void __launch_bounds__(VEC_BLOCK) k_cg2_init(const double *__restrict__ b, const double *__restrict__ q,
                                                        const double *__restrict__ minv, double *__restrict__ r,
                                                        double *__restrict__ u, double *__restrict__ p, double *__restrict__ s,
                                                        int64_t i0, int64_t i1, double *partial, unsigned int *ticket,
                                                        double *sc, int32_t *fl) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    const double ri = bi - q[i];
    const double ui = minv[i] * ri;
    r[i] = ri; u[i] = ui; p[i] = 0.0; s[i] = 0.0;
    acc[0] += ri * ui; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  reduce_finalize<3>(acc, partial, ticket, sc, fl, -1, 0, nullptr, 0, 1);
}

// ---- CG vector kernels, multi-GPU "p2p-fused" variants ---------------------------------------------------------
// The scalar stage that follows a distributed dot product is applied in the PROLOGUE of the next kernel by every
// block redundantly (wait for all ranks' mailbox posts, sum them in rank order, run the recurrence on a private
// copy of the scalar state); block 0 stores the new state into the other half of a double-buffered state array.
// The halo push of p is fused into the epilogue of the p-update.  Per iteration the GPU runs the same three
// kernels as on one GPU; NCCL is not involved.
struct StageCtx {
  const double *sc_in;
  double *sc_out;
  const int32_t *fl_in;
  int32_t *fl_out;
  const P2PDev *pd;
  int epoch, stage, nv;
};
__device__ __forceinline__ void stage_prologue(const StageCtx &c, double *s_sc, int32_t *s_fl) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < S_COUNT; ++i) s_sc[i] = c.sc_in[i];
    for (int i = 0; i < F_COUNT; ++i) s_fl[i] = c.fl_in[i];
    if (!s_fl[F_DONE]) {
      const int slot = c.epoch & 1;
      for (int r = 0; r < c.pd->nranks; ++r)
        p2p_wait(c.pd->mflag[c.pd->rank] + slot * P2P_MAX_RANKS + r, c.epoch, c.pd->err);
      const volatile double *mb = c.pd->mbox[c.pd->rank] + (size_t)slot * P2P_MAX_RANKS * 4;
      for (int i = 0; i < c.nv; ++i) {
        double s = 0.0;
        for (int r = 0; r < c.pd->nranks; ++r) s += mb[r * 4 + i];
        s_sc[S_PEND + i] = s;
      }
      apply_stage(c.stage, s_sc, s_fl);
    }
    if (blockIdx.x == 0) {
      for (int i = 0; i < S_COUNT; ++i) c.sc_out[i] = s_sc[i];
      for (int i = 0; i < F_COUNT; ++i) c.fl_out[i] = s_fl[i];
    }
  }
  __syncthreads();
}
// prologue: alpha = rz / (p.q);  body as k_cg_update;  the (r.z, r.r) sums are posted with `epoch`
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_update_pf(StageCtx st, const double *__restrict__ p,
                                                            const double *__restrict__ q, const double *__restrict__ minv,
                                                            double *__restrict__ x, double *__restrict__ r, int64_t i0,
                                                            int64_t i1, double *partial, unsigned int *ticket,
                                                            double *scratch, int epoch) {
  __shared__ double s_sc[S_COUNT];
  __shared__ int32_t s_fl[F_COUNT];
  stage_prologue(st, s_sc, s_fl);
  if (s_fl[F_DONE]) return;
  const double alpha = s_sc[S_ALPHA];
  double acc[2] = {0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    double ri = r[i] - alpha * q[i];
    r[i] = ri;
    acc[0] += ri * (minv[i] * ri);
    acc[1] += ri * ri;
  }
  reduce_finalize<2>(acc, partial, ticket, scratch, nullptr, ST_CG_UPDATE, 0, st.pd, epoch);
}
// prologue: beta, iteration count, convergence test;  body as k_cg_p;  epilogue: push the owned boundary entries of
// the new p into the neighbours' ghost ranges and raise their halo flags
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_p_pf(StageCtx st, const double *__restrict__ r,
                                                       const double *__restrict__ minv, double *__restrict__ p, int64_t i0,
                                                       int64_t i1, int64_t send_lo, int64_t send_hi, double *peer_lo_dst,
                                                       double *peer_hi_dst, int *peer_lo_flag, int *peer_hi_flag,
                                                       int halo_epoch, unsigned int *ticket) {
  __shared__ double s_sc[S_COUNT];
  __shared__ int32_t s_fl[F_COUNT];
  __shared__ bool last;
  stage_prologue(st, s_sc, s_fl);
  if (s_fl[F_DONE]) return;
  const double beta = s_sc[S_BETA];
  bool pushed = false;
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    const double pn = minv[i] * r[i] + beta * p[i];
    p[i] = pn;
    if (peer_lo_dst && i - i0 < send_lo) { peer_lo_dst[i - i0] = pn; pushed = true; }
    if (peer_hi_dst && i >= i1 - send_hi) { peer_hi_dst[i - (i1 - send_hi)] = pn; pushed = true; }
  }
  if (pushed) __threadfence_system();   // only the few threads that stored into peer memory pay the system fence
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence_system();
    if (peer_lo_flag) *(volatile int *)peer_lo_flag = halo_epoch;
    if (peer_hi_flag) *(volatile int *)peer_hi_flag = halo_epoch;
  }
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
                             const int32_t *__restrict__ diag, int64_t row0, int64_t n, int nf, int jacobi,
                             double *__restrict__ minv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t slice = (i / (64 * nf)) * nf + (i % nf);   // slices interleave the nf components (sell.cu)
  minv[row0 + i] = (jacobi && diag[i] >= 0) ? 1.0 / sell_val[valptr[slice] + diag[i]] : 1.0;
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
  if (comm_active()) {  // single-reduction CG needs u and s = A p as well
    APDX_CHECK(k.z.alloc(n));
    APDX_CHECK(k.s.alloc(n));
    APDX_CUDA(cudaMemsetAsync(k.z.p, 0, n * sizeof(double), pl->stream));
  }
  {  // per-block partial sums of the fused dot products: up to 4 sums, SpMV grids of the interior + boundary launches
    APDX_CHECK(k.partial.alloc(4 * 2 * std::max((size_t)VEC_GRID, (size_t)148 * 32)));
  }
  APDX_CHECK(k.st_sc.alloc(2 * S_COUNT));
  APDX_CHECK(k.st_fl.alloc(2 * F_COUNT));
  APDX_CHECK(k.scratch.alloc(S_COUNT));
  k.scal.adopt(k.st_sc.p, S_COUNT);     // half 0 of the double-buffered scalar state
  k.flags.adopt(k.st_fl.p, F_COUNT);
  APDX_CHECK(k.ticket.alloc(2));
  APDX_CUDA(cudaMemsetAsync(k.ticket.p, 0, 2 * sizeof(unsigned int), pl->stream));
  // halo entries outside the owned range must read as finite numbers
  APDX_CUDA(cudaMemsetAsync(k.p.p, 0, n * sizeof(double), pl->stream));
  return APDX_OK;
}
static int krylov_alloc_bicgstab(apdx_plan *pl) {
  KrylovWork &k = pl->kw;
  if (k.t.p) return APDX_OK;
  const int64_t n = pl->n_free;
  if (!k.s.p) APDX_CHECK(k.s.alloc(n));
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
  bool multi, p2p, mbox;
  const P2PDev *pd, *mpd;
  int fused;
};
static Comm comm_of(apdx_plan *pl) {
  Comm c;
  c.multi = comm_active();
  c.p2p = c.multi && pl->p2p.enabled;
  c.pd = c.p2p ? pl->p2p.dev : nullptr;
  c.mbox = c.multi && !c.p2p && pl->p2p.mbox;
  c.mpd = c.mbox ? pl->p2p.dev : nullptr;
  c.fused = c.multi ? 0 : 1;
  return c;
}

// persistent SpMV grid: the blocks that are resident at once (the kernels are compiled for 2 blocks of 256 threads per
// SM; APDX_SPMV_BPS overrides the blocks per SM for measurements)
static int spmv_bps() {
  static int bps = 0;
  if (!bps) {
    const char *e = getenv("APDX_SPMV_BPS");
    bps = e && atoi(e) > 0 ? atoi(e) : 2;
  }
  return bps;
}
// kernel variant compiled for that many resident blocks per SM (2 = default; 3 and 4: fewer registers, more warps)
// 5 = the default occupancy with the pipelined value prefetch (APDX_SPMV_PIPE=1)
static int spmv_occ() {
  static int occ = 0;
  if (!occ) {
    const int b = spmv_bps();
    const char *e = getenv("APDX_SPMV_PIPE");
    occ = b <= 2 ? ((e && atoi(e) == 1) ? 5 : 2) : (b == 3 ? 3 : 4);
  }
  return occ;
}
static unsigned spmv_grid(int64_t n_slices) {
  static int resident = 0;
  if (!resident) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident = sms * spmv_bps();
  }
  const int64_t nb = (n_slices + VEC_BLOCK / 32 - 1) / (VEC_BLOCK / 32);
  return (unsigned)(nb < resident ? (nb > 0 ? nb : 1) : resident);
}

// part: 0 = all slices in one launch; 1 = interior slices only (deposits its dot partials); 2 = the boundary slices
// at both ends of the owned range (they read ghost columns) + the final reduction over both launches
template <int NDOT>
static int launch_spmv(apdx_plan *pl, const double *x, double *y, const double *w, int stage, int check_done,
                       int halo_epoch = 0, int part = 0) {
  KrylovWork &k = pl->kw;
  Sell &S = pl->sell;
  const Comm c = comm_of(pl);
  SliceRange rg{0, S.n_slices, 0, 0, 0, -1, 1};
  unsigned grid = spmv_grid(S.n_slices);
  if (part != 0) {
    const unsigned g_int = spmv_grid(S.hi_begin - S.lo_end), g_bnd = spmv_grid(S.lo_end + (S.n_slices - S.hi_begin));
    if (part == 1) { rg = SliceRange{S.lo_end, S.hi_begin, 0, 0, 0, (int)(g_int + g_bnd), 0}; grid = g_int; }
    else { rg = SliceRange{0, S.lo_end, S.hi_begin, S.n_slices, (int)g_int, (int)(g_int + g_bnd), 1}; grid = g_bnd; }
  }
  const int epoch = (NDOT > 0 && c.p2p && stage >= 0) ? ++pl->p2p.red_epoch : 0;
#define APDX_SPMV_ARGS                                                                                                 \
  S.sl_w.p, S.sl_m.p, S.valptr.p, S.idxptr.p, S.val.p, S.idx.p, x, y, w, pl->f0, pl->f1, S.n_slices,                   \
      (int32_t)pl->n_free, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, stage, c.fused, check_done,                   \
      (NDOT > 0 && stage >= 0) ? c.pd : (halo_epoch ? c.pd : nullptr), epoch, halo_epoch, rg
#define APDX_SPMV_OCC(NFV, OCCV)                                                                                       \
  do {                                                                                                                 \
    if (S.sym && S.n_mirrored > 0)                                                                                     \
      k_spmv_sell<NDOT, NFV, true, OCCV><<<grid, VEC_BLOCK, 0, pl->stream>>>(APDX_SPMV_ARGS);                          \
    else k_spmv_sell<NDOT, NFV, false, OCCV><<<grid, VEC_BLOCK, 0, pl->stream>>>(APDX_SPMV_ARGS);                      \
  } while (0)
#define APDX_SPMV_NF(NFV)                                                                                              \
  do {                                                                                                                 \
    const int occ = spmv_occ();                                                                                        \
    if (occ == 2) APDX_SPMV_OCC(NFV, 2);                                                                               \
    else if (occ == 3) APDX_SPMV_OCC(NFV, 3);                                                                          \
    else if (occ == 5) APDX_SPMV_OCC(NFV, 5);                                                                          \
    else APDX_SPMV_OCC(NFV, 4);                                                                                        \
  } while (0)
  if (S.nf == 1) APDX_SPMV_NF(1);
  else if (S.nf == 2) APDX_SPMV_NF(2);
  else APDX_SPMV_NF(3);
#undef APDX_SPMV_NF
#undef APDX_SPMV_OCC
#undef APDX_SPMV_ARGS
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
  } else if (c.mbox) {
    k_allreduce_mbox_apply<<<1, 32, 0, pl->stream>>>(stage, nv, k.scal.p, k.flags.p, c.mpd);
  } else {
    APDX_CHECK(comm_allreduce_sum(k.scal.p + S_PEND, nv, pl->stream));
    k_apply_stage<<<1, 1, 0, pl->stream>>>(stage, k.scal.p, k.flags.p);
  }
  pl->stats.kernel_launches += 1;
  return APDX_OK;
}
// y = A v with the ghost entries of v refreshed first.  Multi-GPU over NCCL: the halo exchange runs on a second
// stream while the interior slices are multiplied; the boundary slices follow once the ghosts have arrived.
template <int NDOT>
static int spmv_exchange(apdx_plan *pl, double *v, double *y, const double *w, int stage, int check_done) {
  const Comm c = comm_of(pl);
  if (!c.multi) return launch_spmv<NDOT>(pl, v, y, w, stage, check_done);
  Sell &S = pl->sell;
  const char *ov = getenv("APDX_OVERLAP");
  // measured on 2 B200s: no gain over the in-order exchange, so opt-in (APDX_OVERLAP=1)
  const bool overlap = !c.p2p && !pl->hl.active && (ov && strcmp(ov, "1") == 0) && S.hi_begin > S.lo_end &&
                       (S.lo_end + (S.n_slices - S.hi_begin)) > 0;
  if (!overlap) {
    int he = 0;
    APDX_CHECK(exchange_halo(pl, v, &he));
    return launch_spmv<NDOT>(pl, v, y, w, stage, check_done, he);
  }
  cudaStream_t s = pl->stream, s2 = pl->stream2;
  APDX_CUDA(cudaEventRecord(pl->ev_fork, s));
  APDX_CUDA(cudaStreamWaitEvent(s2, pl->ev_fork, 0));
  APDX_CHECK(comm_halo_exchange(pl, v, s2));
  APDX_CUDA(cudaEventRecord(pl->ev_join, s2));
  APDX_CHECK(launch_spmv<NDOT>(pl, v, y, w, stage, check_done, 0, 1));   // interior
  APDX_CUDA(cudaStreamWaitEvent(s, pl->ev_join, 0));
  return launch_spmv<NDOT>(pl, v, y, w, stage, check_done, 0, 2);        // boundary + reduction
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
                                                                      pl->f0, pl->sell.n_rows, pl->sell.nf, 1, k.p.p);
  for (int i = 0; i < 10; ++i) APDX_CHECK(launch_spmv<1>(pl, k.p.p, k.q.p, k.p.p, -1, 0));   // warm-up
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
                                                                      pl->f0, pl->sell.n_rows, pl->sell.nf, o->jacobi, k.minv.p);
  double sc_h[S_COUNT] = {0};
  sc_h[S_TOL2] = o->rtol * o->rtol;
  sc_h[S_SS] = o->atol * o->atol;
  int32_t fl_h[F_COUNT] = {0, 0, 0, maxiter, 0};
  APDX_CUDA(cudaMemcpyAsync(k.scal.p, sc_h, sizeof(sc_h), cudaMemcpyHostToDevice, s));
  APDX_CUDA(cudaMemcpyAsync(k.flags.p, fl_h, sizeof(fl_h), cudaMemcpyHostToDevice, s));
  pl->stats.kernel_launches += 1;

  const char *cm = getenv("APDX_COMM");
  // multi-GPU CG default: the three-kernel CG with NCCL halo + all-reduces.  Measured alternatives, opt-in through
  // APDX_COMM: cg2 = single-reduction CG; p2p / fused = CUDA-IPC peer stores instead of NCCL
  const bool cg2 = !bi && c.multi && !c.p2p && cm && strcmp(cm, "cg2") == 0;

  // q = A x0 (x is the caller's buffer: its ghost entries travel with NCCL).  The Newton paths start from x0 = 0
  // (k_rhs_reduced / k_rhs_gather zero it, ghosts included) and say so: q = 0 without a product.
  if (pl->x0_is_zero) {
    APDX_CUDA(cudaMemsetAsync(bi ? k.t.p : k.q.p, 0, (size_t)pl->n_free * sizeof(double), s));
    pl->x0_is_zero = false;
  } else {
    if (c.multi) APDX_CHECK(comm_halo_exchange(pl, x, s));
    APDX_CHECK(launch_spmv<0>(pl, x, bi ? k.t.p : k.q.p, nullptr, 0, 0));
  }
  if (cg2) {
    k_cg2_init<<<VEC_GRID, VEC_BLOCK, 0, s>>>(rhs, k.q.p, k.minv.p, k.r.p, k.z.p, k.p.p, k.s.p, i0, i1, k.partial.p,
                                              k.ticket.p, k.scal.p, k.flags.p);
    APDX_CHECK(comm_halo_exchange(pl, k.z.p, s));
    APDX_CHECK(launch_spmv<1>(pl, k.z.p, k.q.p, k.z.p, -1, 0));
    APDX_CHECK(comm_allreduce_sum(k.scal.p + S_PEND, 4, s));
    k_apply_stage<<<1, 1, 0, s>>>(ST_CG2_INIT, k.scal.p, k.flags.p);
    pl->stats.kernel_launches += 2;
  } else if (!bi) {
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
  int launched = 0, he = 0, par = 0;
  const bool trace_on = getenv("APDX_TRACE") != nullptr;
  const char *ge = getenv("APDX_GRAPH");
  const bool graphs_on = !(ge && strcmp(ge, "0") == 0);
  std::vector<std::pair<cudaEvent_t, int>> trace;
  // 16-byte vector accesses in the CG vector kernels need every vector 16-byte aligned at even indices
  const bool vec16 = (((uintptr_t)x | (uintptr_t)k.r.p | (uintptr_t)k.p.p | (uintptr_t)k.q.p | (uintptr_t)k.minv.p) & 15) == 0;
  // p2p-fused CG keeps the scalar state double-buffered in st_sc / st_fl; half 0 aliases k.scal / k.flags
  const bool pfused = !bi && c.p2p && pl->sell.nf == 1 && p2p_is_heap_vector(pl, k.p.p) && !(cm && strcmp(cm, "p2p") == 0);
  while (true) {
    APDX_CUDA(cudaMemcpyAsync(fl_pin, (pfused ? k.st_fl.p + par * F_COUNT : k.flags.p), sizeof(int32_t) * F_COUNT,
                              cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaMemcpyAsync(sc_pin, (pfused ? k.st_sc.p + par * S_COUNT : k.scal.p), sizeof(double) * S_COUNT,
                              cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaStreamSynchronize(s));
    if (fl_pin[F_DONE] || launched >= maxiter) break;
    int todo = maxiter - launched < chunk ? maxiter - launched : chunk;
    const bool tracing = trace_on && launched == chunk;   // APDX_TRACE: time every kernel of the second chunk
    auto TR = [&](int label) {
      if (!tracing) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, s);
      trace.push_back({e, label});
    };
    TR(-1);
    auto one_iteration = [&]() -> int {
      if (cg2) {
        // ---- multi-GPU default: vector kernel, halo, SpMV, ONE all-reduce of (w.u, r.u, r.r), scalar stage ----
        k_cg2_vec<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.q.p, k.minv.p, k.z.p, k.p.p, k.s.p, x, k.r.p, i0, i1, k.partial.p,
                                                 k.ticket.p, k.scal.p, k.flags.p);
        TR(3);
        APDX_CHECK(comm_halo_exchange(pl, k.z.p, s));
        TR(0);
        APDX_CHECK(launch_spmv<1>(pl, k.z.p, k.q.p, k.z.p, -1, 1));
        TR(1);
        APDX_CHECK(comm_allreduce_sum(k.scal.p + S_PEND, 3, s));
        k_apply_stage<<<1, 1, 0, s>>>(ST_CG2_ITER, k.scal.p, k.flags.p);
        TR(2);
        pl->stats.kernel_launches += 2;
      } else if (!bi && pfused) {
        // ---- multi-GPU CG, three kernels per iteration, stages in the prologues (see k_cg_update_pf) ----
        P2P &P = pl->p2p;
        if (he == 0) APDX_CHECK(exchange_halo(pl, k.p.p, &he));   // first iteration: standalone push of p
        {
          // SpMV posts (p.q) with epoch e1; it reads the done flag of the current state half
          Sell &S = pl->sell;
          const unsigned grid = spmv_grid(S.n_slices);
          const int e1 = ++P.red_epoch;
          // (the opt-in p2p-fused CG is only wired for scalar problems)
#define APDX_PF_ARGS                                                                                                   \
  S.sl_w.p, S.sl_m.p, S.valptr.p, S.idxptr.p, S.val.p, S.idx.p, k.p.p, k.q.p, k.p.p, pl->f0, pl->f1, S.n_slices,       \
      (int32_t)pl->n_free, k.partial.p, k.ticket.p, k.scratch.p, k.st_fl.p + par * F_COUNT, ST_CG_PQ, 0, 1, pd, e1, he, \
      SliceRange{0, S.n_slices, 0, 0, 0, -1, 1}
          if (S.sym && S.n_mirrored > 0) k_spmv_sell<1, 1, true, 2><<<grid, VEC_BLOCK, 0, s>>>(APDX_PF_ARGS);
          else k_spmv_sell<1, 1, false, 2><<<grid, VEC_BLOCK, 0, s>>>(APDX_PF_ARGS);
#undef APDX_PF_ARGS
          TR(1);
          pl->stats.spmv_launches += 1;
          StageCtx c1{k.st_sc.p + par * S_COUNT, k.st_sc.p + (par ^ 1) * S_COUNT, k.st_fl.p + par * F_COUNT,
                      k.st_fl.p + (par ^ 1) * F_COUNT, pd, e1, ST_CG_PQ, 1};
          const int e2 = ++P.red_epoch;
          k_cg_update_pf<<<VEC_GRID, VEC_BLOCK, 0, s>>>(c1, k.p.p, k.q.p, k.minv.p, x, k.r.p, i0, i1, k.partial.p, k.ticket.p,
                                                        k.scratch.p, e2);
          TR(3);
          par ^= 1;
          StageCtx c2{k.st_sc.p + par * S_COUNT, k.st_sc.p + (par ^ 1) * S_COUNT, k.st_fl.p + par * F_COUNT,
                      k.st_fl.p + (par ^ 1) * F_COUNT, pd, e2, ST_CG_UPDATE, 2};
          he = ++P.halo_epoch;
          const int64_t off = k.p.p - P.vec_base;
          double *lo = pl->rank_lo >= 0 ? P.peer_vec[0] + off + P.peer_lo_f1 : nullptr;
          double *hi = pl->rank_hi >= 0 ? P.peer_vec[1] + off : nullptr;
          k_cg_p_pf<<<VEC_GRID, VEC_BLOCK, 0, s>>>(c2, k.r.p, k.minv.p, k.p.p, i0, i1, pl->send_lo, pl->send_hi, lo, hi,
                                                   P.peer_hflag[0], P.peer_hflag[1], he, k.ticket.p + 1);
          TR(5);
          par ^= 1;
          pl->stats.kernel_launches += 3;
        }
      } else if (!bi) {
        if (c.p2p) {
          APDX_CHECK(exchange_halo(pl, k.p.p, &he));
          TR(0);
          APDX_CHECK(launch_spmv<1>(pl, k.p.p, k.q.p, k.p.p, ST_CG_PQ, 1, he));
        } else {
          APDX_CHECK(spmv_exchange<1>(pl, k.p.p, k.q.p, k.p.p, ST_CG_PQ, 1));
        }
        TR(1);
        APDX_CHECK(finish_stage(pl, ST_CG_PQ, 1));
        TR(2);
        if (vec16) k_cg_update<true><<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.q.p, k.minv.p, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p,
                                                                   k.flags.p, fused, pd, next_red_epoch(pl));
        else k_cg_update<false><<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.q.p, k.minv.p, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p,
                                                               k.flags.p, fused, pd, next_red_epoch(pl));
        TR(3);
        APDX_CHECK(finish_stage(pl, ST_CG_UPDATE, 2));
        TR(4);
        if (vec16) k_cg_p<true><<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.minv.p, k.p.p, x, i0, i1, k.scal.p, k.flags.p, k.ticket.p + 1);
        else k_cg_p<false><<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.minv.p, k.p.p, x, i0, i1, k.scal.p, k.flags.p, k.ticket.p + 1);
        TR(5);
        pl->stats.kernel_launches += 2;
      } else {
        k_bi_p<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.q.p, k.minv.p, k.p.p, k.phat.p, i0, i1, k.scal.p, k.flags.p);
        if (c.p2p) {
          APDX_CHECK(exchange_halo(pl, k.phat.p, &he));
          APDX_CHECK(launch_spmv<1>(pl, k.phat.p, k.q.p, k.r0.p, ST_BI_R0V, 1, he));
        } else {
          APDX_CHECK(spmv_exchange<1>(pl, k.phat.p, k.q.p, k.r0.p, ST_BI_R0V, 1));
        }
        APDX_CHECK(finish_stage(pl, ST_BI_R0V, 1));
        k_bi_s<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.r.p, k.q.p, k.minv.p, k.s.p, k.shat.p, i0, i1, k.partial.p,
                                              k.ticket.p, k.scal.p, k.flags.p, fused, pd, next_red_epoch(pl));
        APDX_CHECK(finish_stage(pl, ST_BI_S, 1));
        if (c.p2p) {
          APDX_CHECK(exchange_halo(pl, k.shat.p, &he));
          APDX_CHECK(launch_spmv<2>(pl, k.shat.p, k.t.p, k.s.p, ST_BI_T, 1, he));
        } else {
          APDX_CHECK(spmv_exchange<2>(pl, k.shat.p, k.t.p, k.s.p, ST_BI_T, 1));
        }
        APDX_CHECK(finish_stage(pl, ST_BI_T, 2));
        k_bi_x<<<VEC_GRID, VEC_BLOCK, 0, s>>>(k.phat.p, k.shat.p, k.s.p, k.t.p, k.r0.p, x, k.r.p, i0, i1,
                                              k.partial.p, k.ticket.p, k.scal.p, k.flags.p, fused, pd, next_red_epoch(pl));
        APDX_CHECK(finish_stage(pl, ST_BI_X, 2));
        pl->stats.kernel_launches += 3;
      }
      return APDX_OK;
    };
    // CUDA graph of one chunk of iterations: the loop is launch-bound for small systems and at high GPU counts.
    // Kernels turn into no-ops once the device-side done flag is set, so replaying a whole chunk is always safe.
    const int mode_id = cg2 ? 1 : (c.mbox ? 3 : (c.multi ? 2 : 0));
    const bool graph_ok = graphs_on && !tracing && !c.p2p && todo == chunk;
    KrylovGraph &G = pl->kgraph[bi ? 1 : 0];
    if (graph_ok && G.exec && G.rhs == rhs && G.x == x && G.mode == mode_id && G.chunk == chunk && G.i0 == i0 && G.i1 == i1) {
      APDX_CUDA(cudaGraphLaunch(G.exec, s));
      pl->stats.kernel_launches += G.kernel_launches;
      pl->stats.spmv_launches += G.spmv_launches;
    } else if (graph_ok) {
      if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
      const double k0 = pl->stats.kernel_launches, s0 = pl->stats.spmv_launches;
      APDX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int rc = APDX_OK;
      for (int it = 0; it < todo && rc == APDX_OK; ++it) rc = one_iteration();
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc != APDX_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      APDX_REQUIRE(ce == cudaSuccess && graph, APDX_ERR_CUDA, "stream capture of the Krylov chunk failed: %s", cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&G.exec, graph, 0);
      cudaGraphDestroy(graph);
      APDX_REQUIRE(ce == cudaSuccess, APDX_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
      G.rhs = rhs; G.x = x; G.mode = mode_id; G.chunk = chunk; G.i0 = i0; G.i1 = i1;
      G.kernel_launches = pl->stats.kernel_launches - k0;
      G.spmv_launches = pl->stats.spmv_launches - s0;
      APDX_CUDA(cudaGraphLaunch(G.exec, s));
    } else {
      for (int it = 0; it < todo; ++it) APDX_CHECK(one_iteration());
    }
    launched += todo;
    if (tracing) {
      APDX_CUDA(cudaStreamSynchronize(s));
      double sum[6] = {0, 0, 0, 0, 0, 0};
      for (size_t t = 1; t < trace.size(); ++t) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, trace[t - 1].first, trace[t].first);
        if (trace[t].second >= 0) sum[trace[t].second] += ms;
      }
      const char *nm[6] = {"halo", "spmv", "stage_pq", "update", "stage_upd", "p_update"};
      fprintf(stderr, "[apdx trace] us/iteration over %d iterations:", todo);
      for (int l = 0; l < 6; ++l) fprintf(stderr, " %s %.1f", nm[l], 1e3 * sum[l] / todo);
      fprintf(stderr, "\n");
      for (auto &t : trace) cudaEventDestroy(t.first);
      trace.clear();
    }
  }
  APDX_CUDA(cudaGetLastError());
  if (c.p2p || c.mbox) {
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
