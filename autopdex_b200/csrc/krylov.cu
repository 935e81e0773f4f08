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
#include "krylov.cuh"

namespace apdx {

__global__ void k_apply_stage(int stage, double *sc, int32_t *fl) {
  if (fl[F_DONE] && stage != ST_CG_INIT && stage != ST_BI_INIT) return;
  apply_stage(stage, sc, fl);
}

// spin until *flag >= epoch (system-scope visibility); bounded so that a lost peer cannot hang the GPU
__device__ __forceinline__ void p2p_wait(const int *flag, int epoch, int *err) {
  const long long t0 = clock64();
  while (*(volatile const int *)flag < epoch) {
    if (clock64() - t0 > 20000000000ll) {  // ~10 s (ranks may enter a solve seconds apart)
      *err = 1;
      break;
    }
  }
  __threadfence_system();
}

// mbox mode: the whole all-reduce of a dot-product stage in ONE small kernel -- post the local sums into every rank's
// mailbox (peer stores over NVLink), raise the flags, wait for every rank's post, sum in rank order (identical bits on
// all ranks), run the scalar recurrence.  The reduction counter lives on the device (every rank runs the same
// sequence of reductions), so the kernel has no per-launch argument and can be replayed inside the CUDA graph of a
// Krylov chunk.  Replaces ncclAllReduce (8-24 bytes, ~22 us) + k_apply_stage.  Mailboxes are double-buffered by the
// parity of the counter: a rank can be at most one reduction ahead of the slowest one.
__global__ void k_allreduce_mbox_apply(int stage, int nv, double *sc, int32_t *fl, const P2PDev *pd) {
  if (fl[F_DONE] && stage != ST_CG_INIT && stage != ST_BI_INIT) return;
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

// Sliced-ELL SpMV (layout: sell.cu).  One warp per slice of 64 rows; lane owns the local rows lane and lane + 32, so
// each load instruction of the warp covers 256 contiguous bytes.  Offset mode: column = row + off[j] (offsets held one
// per lane, x gathers coalesced); explicit mode: one int32 column per stored entry.
// n_cols: length of x (clamp target of the padded offsets, whose values are exact zeros).
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
// a multiple of the batch B: compile-time trip counts, no predicates, offsets held one per lane and handed out by
// shuffles.  Half the instructions of the generic path.
// (Measured and removed in round 2: a schedule that gave the 8 warps of a block the same in-plane position of 2/4/8
// consecutive mesh planes, so that the plane-distance mirrored columns and x windows would hit this SM's L1 instead of
// the L2 -- L2->L1 traffic fell only from 4.39 to 4.05 GB per launch and the time did not move, 0.483 -> 0.493-0.501 ms,
// profiles/r02d_spmv_planes_*.jsonl.  Also, profiles/r02a_spmv_variants_*.jsonl: kernels compiled for 3 / 4 resident blocks per
// SM with shallower batches -- 0.500 / 0.695 ms against 0.486 ms at P256 -- and a variant that requested the next
// slice's first value batch during the mirrored phase -- 0.501 ms; nf = 3 at 96^3: 0.297 / 0.325 / 0.275 against 0.246.)
template <int B, bool SYM>
__device__ __forceinline__ void spmv_stored_fast(int32_t offl, const double *__restrict__ vpc, int32_t nb,
                                                 const double *__restrict__ xr0, const double *__restrict__ xr1, double &a0,
                                                 double &a1) {
  for (int32_t jb = 0; jb < nb; jb += B) {
    double va[B], vb[B], xa[B], xb[B];
    int32_t off[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const double *q = vpc + (size_t)(jb + u) * 64;
      va[u] = SYM ? __ldg(q) : __ldcs(q);
      vb[u] = SYM ? __ldg(q + 32) : __ldcs(q + 32);
    }
#pragma unroll
    for (int u = 0; u < B; ++u) off[u] = __shfl_sync(0xffffffffu, offl, jb + u);   // lane j holds column j's offset
#pragma unroll
    for (int u = 0; u < B; ++u) { xa[u] = __ldg(xr0 + off[u]); xb[u] = __ldg(xr1 + off[u]); }
#pragma unroll
    for (int u = 0; u < B; ++u) { a0 += va[u] * xa[u]; a1 += vb[u] * xb[u]; }
  }
}

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

template <int NDOT, int NF, bool SYM>
__global__ void __launch_bounds__(VEC_BLOCK, 2)
    k_spmv_sell(const int32_t *__restrict__ sl_w, const int32_t *__restrict__ sl_m, const int64_t *__restrict__ valptr,
                const int64_t *__restrict__ idxptr, const double *__restrict__ val, const int32_t *__restrict__ idx,
                const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, int64_t row0,
                int64_t row1, int64_t n_slices, int32_t n_cols, double *partial, unsigned int *ticket, double *sc,
                int32_t *fl, int stage, int fused, int check_done) {
  if (check_done && fl[F_DONE]) return;
  const int lane = threadIdx.x & 31;
  constexpr int WPB = VEC_BLOCK / 32;
  constexpr int U = SPMV_U;          // generic paths: columns per batch
  double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
  for (int i = 0; i < (NDOT > 0 ? NDOT : 1); ++i) acc[i] = 0.0;
  const int64_t si_begin = (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
  const int64_t si_end = n_slices;
  const int64_t si_step = (int64_t)gridDim.x * WPB;
  // The per-slice metadata is a chain of dependent loads (header -> offsets -> x gathers): the header of the warp's
  // NEXT slice is requested before the current slice is processed, its first 32 offsets half-way through.
  struct Hdr { int32_t wenc, M; int64_t vp, ip; };
  auto load_hdr = [&](int64_t s, Hdr &h) {
    h.wenc = 0; h.M = 0; h.vp = 0; h.ip = 0;
    if (s < si_end) {
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
  for (int64_t s = si_begin; s < si_end; s += si_step) {
    Hdr nxt;
    load_hdr(s + si_step, nxt);
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
      if (fast && (nb % 9 == 0 || nb % 8 == 0 || nb % 7 == 0)) {
        const int32_t offl = jc == 0 ? offl0 : __ldg(ip + jc + min(lane, nb - 1));
        if (nb % 9 == 0) spmv_stored_fast<9, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else if (nb % 8 == 0) spmv_stored_fast<8, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
        else spmv_stored_fast<7, SYM>(offl, vp + (size_t)jc * 64, nb, xr0, xr1, a0, a1);
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
    if (SYM && M > 0) {
      if (fast) {
        if (cur.M & SELL_MB7) spmv_mirrored_fast<7>(tab, M, val, xr0, xr1, lane, a0, a1);
        else spmv_mirrored_fast<8>(tab, M, val, xr0, xr1, lane, a0, a1);
      } else {
        if (cur.M & SELL_MB7) spmv_mirrored<7>(tab, M, val, x, lane, rr0, rr1, n_cols, a0, a1);
        else spmv_mirrored<8>(tab, M, val, x, lane, rr0, rr1, n_cols, a0, a1);
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
  if constexpr (NDOT > 0) reduce_finalize<NDOT>(acc, partial, ticket, sc, fl, stage, fused);
}

// ---- CG vector kernels -----------------------------------------------------------------------
// r = b - q ; z = minv r ; p = z ; sums (r.z, r.r, b.b)
__global__ void __launch_bounds__(VEC_BLOCK) k_cg_init(const double *__restrict__ b, const double *__restrict__ q,
                                                       const double *__restrict__ minv, double *__restrict__ r,
                                                       double *__restrict__ p, int64_t i0, int64_t i1,
                                                       double *partial, unsigned int *ticket, double *sc, int32_t *fl, int stage, int fused) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double bi = b[i];
    double ri = bi - q[i];
    double zi = minv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    acc[0] += ri * zi; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  reduce_finalize<3>(acc, partial, ticket, sc, fl, stage, fused);
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
                                                         unsigned int *ticket, double *sc, int32_t *fl, int fused) {
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
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_CG_UPDATE, fused);
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

// ---- BiCGSTAB vector kernels --------------------------------------------------------------------
// r = b - q ; r0 = r ; p = 0 ; v = 0 ; sums (r0.r, r.r, b.b)
__global__ void __launch_bounds__(VEC_BLOCK) k_bi_init(const double *__restrict__ b, const double *__restrict__ q,
                                                       double *__restrict__ r, double *__restrict__ r0,
                                                       double *__restrict__ p, double *__restrict__ v, int64_t i0,
                                                       int64_t i1, double *partial, unsigned int *ticket, double *sc,
                                                       int32_t *fl, int fused) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double bi = b[i];
    double ri = bi - q[i];
    r[i] = ri; r0[i] = ri; p[i] = 0.0; v[i] = 0.0;
    acc[0] += ri * ri; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  reduce_finalize<3>(acc, partial, ticket, sc, fl, ST_BI_INIT, fused);
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
                                                    double *partial, unsigned int *ticket, double *sc, int32_t *fl, int fused) {
  if (fl[F_DONE]) return;
  const double alpha = sc[S_ALPHA];
  double acc[1] = {0.0};
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    double si = r[i] - alpha * v[i];
    s[i] = si;
    shat[i] = minv[i] * si;
    acc[0] += si * si;
  }
  reduce_finalize<1>(acc, partial, ticket, sc, fl, ST_BI_S, fused);
}
// x += alpha phat + omega shat ; r = s - omega t ; sums (r0.r, r.r)
__global__ void __launch_bounds__(VEC_BLOCK) k_bi_x(const double *__restrict__ phat, const double *__restrict__ shat,
                                                    const double *__restrict__ s, const double *__restrict__ t,
                                                    const double *__restrict__ r0, double *__restrict__ x,
                                                    double *__restrict__ r, int64_t i0, int64_t i1, double *partial,
                                                    unsigned int *ticket, double *sc, int32_t *fl, int fused) {
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
  reduce_finalize<2>(acc, partial, ticket, sc, fl, ST_BI_X, fused);
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
int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
// grid of the streaming vector kernels: 8 blocks of 256 threads per SM (all resident)
unsigned vec_grid() { return (unsigned)(sm_count() * 8); }

int krylov_alloc(apdx_plan *pl) {
  KrylovWork &k = pl->kw;
  if (k.r.p) return APDX_OK;
  const int64_t n = pl->n_free;
  APDX_CHECK(k.r.alloc(n));
  APDX_CHECK(k.p.alloc(n));
  APDX_CHECK(k.q.alloc(n));
  APDX_CHECK(k.minv.alloc(n));
  // per-block partial sums of the fused dot products: up to 4 sums per kernel
  APDX_CHECK(k.partial.alloc(std::max({4 * (size_t)vec_grid(), 4 * (size_t)sm_count() * 32, (size_t)NORM_GRID + 8})));
  APDX_CHECK(k.scal.alloc(S_COUNT));
  APDX_CHECK(k.flags.alloc(F_COUNT));
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
  APDX_CHECK(k.s.alloc(n));
  APDX_CHECK(k.t.alloc(n));
  APDX_CHECK(k.phat.alloc(n));
  APDX_CHECK(k.shat.alloc(n));
  APDX_CHECK(k.r0.alloc(n));
  APDX_CUDA(cudaMemsetAsync(k.phat.p, 0, n * sizeof(double), pl->stream));
  APDX_CUDA(cudaMemsetAsync(k.shat.p, 0, n * sizeof(double), pl->stream));
  return APDX_OK;
}

// ---- launch helpers ---------------------------------------------------------------------------------------
// comm modes: single GPU (dot products finalised and the scalar stage applied inside the producing kernel), NCCL
// (ncclSend/Recv halo + ncclAllReduce + one-thread scalar kernel), mbox (the all-reduce + scalar stage of a dot product
// in ONE small kernel over peer-memory mailboxes; halo on NCCL)
struct Comm {
  bool multi, mbox;
  const P2PDev *mpd;
  int fused;
};
static Comm comm_of(apdx_plan *pl) {
  Comm c;
  c.multi = comm_active();
  c.mbox = c.multi && pl->p2p.mbox;
  c.mpd = c.mbox ? pl->p2p.dev : nullptr;
  c.fused = c.multi ? 0 : 1;
  return c;
}

// persistent SpMV grid: exactly the blocks that are resident at once (the kernel is compiled for 2 blocks of 256
// threads per SM)
static unsigned spmv_grid(int64_t n_slices) {
  const int resident = sm_count() * 2;
  const int64_t nb = (n_slices + VEC_BLOCK / 32 - 1) / (VEC_BLOCK / 32);
  return (unsigned)(nb < resident ? (nb > 0 ? nb : 1) : resident);
}

template <int NDOT>
static int launch_spmv(apdx_plan *pl, const double *x, double *y, const double *w, int stage, int check_done) {
  KrylovWork &k = pl->kw;
  Sell &S = pl->sell;
  const Comm c = comm_of(pl);
  const unsigned grid = spmv_grid(S.n_slices);
#define APDX_SPMV_ARGS                                                                                                 \
  S.sl_w.p, S.sl_m.p, S.valptr.p, S.idxptr.p, S.val.p, S.idx.p, x, y, w, pl->f0, pl->f1, S.n_slices,                   \
      (int32_t)pl->n_free, k.partial.p, k.ticket.p, k.scal.p, k.flags.p, stage, c.fused, check_done
#define APDX_SPMV_NF(NFV)                                                                                              \
  do {                                                                                                                 \
    if (S.sym && S.n_mirrored > 0) k_spmv_sell<NDOT, NFV, true><<<grid, VEC_BLOCK, 0, pl->stream>>>(APDX_SPMV_ARGS);   \
    else k_spmv_sell<NDOT, NFV, false><<<grid, VEC_BLOCK, 0, pl->stream>>>(APDX_SPMV_ARGS);                            \
  } while (0)
  if (S.nf == 1) APDX_SPMV_NF(1);
  else if (S.nf == 2) APDX_SPMV_NF(2);
  else APDX_SPMV_NF(3);
#undef APDX_SPMV_NF
#undef APDX_SPMV_ARGS
  pl->stats.spmv_launches += 1;
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

int spmv_launch(apdx_plan *pl, const double *x, double *y, const double *w, int ndot, int stage, int check_done) {
  if (ndot == 0) return launch_spmv<0>(pl, x, y, nullptr, stage, check_done);
  if (ndot == 1) return launch_spmv<1>(pl, x, y, w, stage, check_done);
  return launch_spmv<2>(pl, x, y, w, stage, check_done);
}

// after a dot-product kernel: multi-GPU reduction of the pending sums + scalar stage
int krylov_finish_stage(apdx_plan *pl, int stage, int nv) {
  const Comm c = comm_of(pl);
  if (!c.multi) return APDX_OK;
  KrylovWork &k = pl->kw;
  if (c.mbox) {
    k_allreduce_mbox_apply<<<1, 32, 0, pl->stream>>>(stage, nv, k.scal.p, k.flags.p, c.mpd);
  } else {
    APDX_CHECK(comm_allreduce_sum(k.scal.p + S_PEND, nv, pl->stream));
    k_apply_stage<<<1, 1, 0, pl->stream>>>(stage, k.scal.p, k.flags.p);
  }
  pl->stats.kernel_launches += 1;
  return APDX_OK;
}
// y = A v with the ghost entries of v refreshed first (multi-GPU: halo exchange of the interface dofs).
// Splitting the product into an interior launch overlapped with the exchange and a boundary launch was measured on
// 2 and 4 B200s in round 1 (565 vs 562 us per iteration) and removed.
template <int NDOT>
static int spmv_exchange(apdx_plan *pl, double *v, double *y, const double *w, int stage, int check_done) {
  if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, v, pl->stream));
  return launch_spmv<NDOT>(pl, v, y, w, stage, check_done);
}

int spmv_reduced(apdx_plan *pl, const double *x, double *y) {
  APDX_REQUIRE(pl->have_sell_values, APDX_ERR_STATE, "no assembled tangent: call apdx_assemble first");
  APDX_CHECK(krylov_alloc(pl));
  return spmv_exchange<0>(pl, const_cast<double *>(x), y, nullptr, 0, 0);
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
  if (o->jacobi == APDX_PRECOND_MULTIGRID) return mg_pcg_solve(pl, o, rhs, x, iters, relres);
  const bool bi = o->method == APDX_KRYLOV_BICGSTAB;
  if (bi) APDX_CHECK(krylov_alloc_bicgstab(pl));
  KrylovWork &k = pl->kw;
  cudaStream_t s = pl->stream;
  const Comm c = comm_of(pl);
  const int64_t i0 = pl->f0, i1 = pl->f1, n = pl->n_free;
  const int fused = c.fused;
  const unsigned VG = vec_grid();
  // default iteration limit of jax.scipy.sparse.linalg.cg / bicgstab: 10 n (the 32-bit counter caps it)
  const int64_t n_glob_hint = n * (int64_t)(c.multi ? comm_size() : 1);
  const int maxiter = o->maxiter > 0 ? o->maxiter : (int)std::min<int64_t>(10 * n_glob_hint, 2000000000ll);
  const int chunk = o->check_every > 0 ? o->check_every : 32;
  nvtx_push("apdx:krylov");

  k_jacobi_inv<<<(unsigned)((pl->sell.n_rows + 255) / 256), 256, 0, s>>>(pl->sell.val.p, pl->sell.valptr.p, pl->sell.diag.p,
                                                                      pl->f0, pl->sell.n_rows, pl->sell.nf, o->jacobi, k.minv.p);
  double sc_h[S_COUNT] = {0};
  sc_h[S_TOL2] = o->rtol * o->rtol;
  sc_h[S_SS] = o->atol * o->atol;
  int32_t fl_h[F_COUNT] = {0, 0, 0, maxiter, 0};
  APDX_CUDA(cudaMemcpyAsync(k.scal.p, sc_h, sizeof(sc_h), cudaMemcpyHostToDevice, s));
  APDX_CUDA(cudaMemcpyAsync(k.flags.p, fl_h, sizeof(fl_h), cudaMemcpyHostToDevice, s));
  pl->stats.kernel_launches += 1;

  // q = A x0 (x is the caller's buffer: its ghost entries travel with NCCL).  The Newton paths start from x0 = 0
  // (k_rhs_reduced / k_rhs_gather zero it, ghosts included) and say so: q = 0 without a product.
  if (pl->x0_is_zero) {
    APDX_CUDA(cudaMemsetAsync(bi ? k.t.p : k.q.p, 0, (size_t)pl->n_free * sizeof(double), s));
    pl->x0_is_zero = false;
  } else {
    APDX_CHECK(spmv_exchange<0>(pl, x, bi ? k.t.p : k.q.p, nullptr, 0, 0));
  }
  if (!bi) {
    k_cg_init<<<VG, VEC_BLOCK, 0, s>>>(rhs, k.q.p, k.minv.p, k.r.p, k.p.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p,
                                       k.flags.p, ST_CG_INIT, fused);
    pl->stats.kernel_launches += 1;
    APDX_CHECK(krylov_finish_stage(pl, ST_CG_INIT, 3));
  } else {
    k_bi_init<<<VG, VEC_BLOCK, 0, s>>>(rhs, k.t.p, k.r.p, k.r0.p, k.p.p, k.q.p, i0, i1, k.partial.p, k.ticket.p,
                                       k.scal.p, k.flags.p, fused);
    pl->stats.kernel_launches += 1;
    APDX_CHECK(krylov_finish_stage(pl, ST_BI_INIT, 3));
  }

  int32_t *fl_pin = reinterpret_cast<int32_t *>(pl->pinned);
  double *sc_pin = pl->pinned + 8;
  int launched = 0;
  const bool trace_on = getenv("APDX_TRACE") != nullptr;
  const char *ge = getenv("APDX_GRAPH");
  const bool graphs_on = !(ge && strcmp(ge, "0") == 0);
  std::vector<std::pair<cudaEvent_t, int>> trace;
  // 16-byte vector accesses in the CG vector kernels need every vector 16-byte aligned at even indices
  const bool vec16 = (((uintptr_t)x | (uintptr_t)k.r.p | (uintptr_t)k.p.p | (uintptr_t)k.q.p | (uintptr_t)k.minv.p) & 15) == 0;
  while (true) {
    APDX_CUDA(cudaMemcpyAsync(fl_pin, k.flags.p, sizeof(int32_t) * F_COUNT, cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaMemcpyAsync(sc_pin, k.scal.p, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, s));
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
      if (!bi) {
        if (c.multi) APDX_CHECK(comm_halo_exchange(pl, k.p.p, s));
        TR(0);
        APDX_CHECK(launch_spmv<1>(pl, k.p.p, k.q.p, k.p.p, ST_CG_PQ, 1));
        TR(1);
        APDX_CHECK(krylov_finish_stage(pl, ST_CG_PQ, 1));
        TR(2);
        if (vec16) k_cg_update<true><<<VG, VEC_BLOCK, 0, s>>>(k.q.p, k.minv.p, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p,
                                                             k.flags.p, fused);
        else k_cg_update<false><<<VG, VEC_BLOCK, 0, s>>>(k.q.p, k.minv.p, k.r.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p,
                                                         k.flags.p, fused);
        TR(3);
        APDX_CHECK(krylov_finish_stage(pl, ST_CG_UPDATE, 2));
        TR(4);
        if (vec16) k_cg_p<true><<<VG, VEC_BLOCK, 0, s>>>(k.r.p, k.minv.p, k.p.p, x, i0, i1, k.scal.p, k.flags.p, k.ticket.p + 1);
        else k_cg_p<false><<<VG, VEC_BLOCK, 0, s>>>(k.r.p, k.minv.p, k.p.p, x, i0, i1, k.scal.p, k.flags.p, k.ticket.p + 1);
        TR(5);
        pl->stats.kernel_launches += 2;
      } else {
        k_bi_p<<<VG, VEC_BLOCK, 0, s>>>(k.r.p, k.q.p, k.minv.p, k.p.p, k.phat.p, i0, i1, k.scal.p, k.flags.p);
        TR(5);
        if (c.multi) APDX_CHECK(comm_halo_exchange(pl, k.phat.p, s));
        TR(0);
        APDX_CHECK(launch_spmv<1>(pl, k.phat.p, k.q.p, k.r0.p, ST_BI_R0V, 1));
        TR(1);
        APDX_CHECK(krylov_finish_stage(pl, ST_BI_R0V, 1));
        TR(2);
        k_bi_s<<<VG, VEC_BLOCK, 0, s>>>(k.r.p, k.q.p, k.minv.p, k.s.p, k.shat.p, i0, i1, k.partial.p, k.ticket.p, k.scal.p,
                                        k.flags.p, fused);
        TR(3);
        APDX_CHECK(krylov_finish_stage(pl, ST_BI_S, 1));
        TR(4);
        if (c.multi) APDX_CHECK(comm_halo_exchange(pl, k.shat.p, s));
        TR(0);
        APDX_CHECK(launch_spmv<2>(pl, k.shat.p, k.t.p, k.s.p, ST_BI_T, 1));
        TR(1);
        APDX_CHECK(krylov_finish_stage(pl, ST_BI_T, 2));
        TR(2);
        k_bi_x<<<VG, VEC_BLOCK, 0, s>>>(k.phat.p, k.shat.p, k.s.p, k.t.p, k.r0.p, x, k.r.p, i0, i1, k.partial.p, k.ticket.p,
                                        k.scal.p, k.flags.p, fused);
        TR(3);
        APDX_CHECK(krylov_finish_stage(pl, ST_BI_X, 2));
        TR(4);
        pl->stats.kernel_launches += 3;
      }
      return APDX_OK;
    };
    // CUDA graph of one chunk of iterations: the loop is launch-bound for small systems and at high GPU counts.
    // Kernels turn into no-ops once the device-side done flag is set, so replaying a whole chunk is always safe.
    const int mode_id = c.mbox ? 3 : (c.multi ? 2 : 0);
    const bool graph_ok = graphs_on && !tracing && todo == chunk;
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
      const char *nm[6] = {"halo", "spmv", "stage_after_spmv", "vector_dot", "stage_after_vector", "vector_nodot"};
      fprintf(stderr, "[apdx trace] %s, us/iteration over %d iterations:", bi ? "bicgstab" : "cg", todo);
      for (int l = 0; l < 6; ++l) fprintf(stderr, " %s %.1f", nm[l], 1e3 * sum[l] / todo);
      fprintf(stderr, "\n");
      for (auto &t : trace) cudaEventDestroy(t.first);
      trace.clear();
    }
  }
  nvtx_pop();
  APDX_CUDA(cudaGetLastError());
  if (c.mbox) {
    int err = 0;
    APDX_CUDA(cudaMemcpy(&err, pl->p2p.err_d, sizeof(int), cudaMemcpyDeviceToHost));
    APDX_REQUIRE(err == 0, APDX_ERR_NCCL, "peer-to-peer wait timed out (a rank stopped posting reductions)");
  }
  const double rr = sc_pin[S_BB] > 0 ? sqrt(sc_pin[S_RR] / sc_pin[S_BB]) : sqrt(sc_pin[S_RR]);
  if (iters) *iters = fl_pin[F_ITERS];
  if (relres) *relres = rr;
  pl->stats.krylov_iters += fl_pin[F_ITERS];
  // outcome of the LAST solve (apdx_plan_stats): the reference's jax solvers return info = None and SciPy's spsolve
  // cannot fail silently, so an unconverged Krylov solve must at least be visible to the caller
  pl->stats.krylov_relres = rr;
  pl->stats.krylov_converged = (sc_pin[S_RR] <= sc_pin[S_TOL2] && !fl_pin[F_BREAKDOWN]) ? 1.0 : 0.0;
  if (fl_pin[F_BREAKDOWN]) pl->stats.krylov_converged = 0.0;
  if (fl_pin[F_BREAKDOWN]) {
    set_error("Krylov breakdown (NaN or zero inner product) after %d iterations", fl_pin[F_ITERS]);
  }
  return APDX_OK;
}

}  // namespace apdx
