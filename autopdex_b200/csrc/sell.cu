// Sliced-ELL storage of the reduced system for the Krylov solve, built once per plan on device.
//
// Slices of 64 rows (one warp of the SpMV; lane owns local rows lane and lane + 32).  Inside a slice the
// entries are stored column-major, so a column is one contiguous 512-byte run (two 256-byte warp loads).
// Column indices are compressed per slice: if the union of (col - row) over the slice's rows is
// small ("offset mode"), the slice stores that list of offsets ONCE (W ints instead of 64*W) and
// entry j of every row means column row + off[j]; rows that do not have an offset hold an explicit
// zero.  Finite-element matrices on structured or well-ordered meshes are almost entirely in
// offset mode, which cuts the per-nonzero traffic from 12 to ~8 bytes and makes the x gathers
// coalesced.  Slices with irregular columns fall back to explicit int32 columns ("explicit mode").
// Offset-mode slices store only the columns with offset >= 0 (symmetric tangents, see below).
// The reference has no device SpMV at all; this is the storage behind the `north_star`'s
// sliced-ELL SpMV (its "128-bit loads" were measured against contiguous 8-byte-per-lane loads in session 2: the
// interleaved row ownership they need doubles the L1 tag traffic of the x gathers).
#include <cub/cub.cuh>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace apdx {

constexpr int SELL_C = 64;
constexpr int32_t SELL_BIG = 0x7fffffff;

__device__ __forceinline__ int32_t warp_min(int32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int32_t warp_max(int32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- symmetric ("mirrored") storage --------------------------------------------------------------------------
// Every tangent of the in-scope models is symmetric (Poisson, capacity, linear elasticity, hyperelastic neo-Hooke;
// Dirichlet reduction [:,free][free] keeps it).  An offset-mode slice therefore stores only its columns with
// offset >= 0; a lower column -d is read from the place where A[r-d][r] is already stored: row r-d lives in an
// earlier slice s' as column +d.  For the 64 rows of a slice the partner rows are consecutive rows of at most two
// slices, so one table entry (off, posA, posB, split) per (slice, lower offset) locates all of them:
//   value of local row k  =  val[(k < split ? posA : posB) + k].
// The partner values were streamed from HBM a few MB earlier and are served by the L2 (measured with
// tools/probe_sym_spmv.cu on the 256^3 stencil: 0.41 ms vs 0.55 ms for full storage).  A lower column is mirrored only
// if, for every row of the slice, the partner row is an owned row of an offset-mode slice and the pattern is
// structurally symmetric there; otherwise the slice stores the column itself (first mesh plane, ghost columns of a
// partition, explicit-mode neighbours).  Stored columns are ordered: offsets >= 0 ascending, then stored lower ones.
// Row segments without any entry point at a block of 64 zeros appended to the value array.

struct SliceGeo {
  int64_t row0, row1, n_rows, n_slices, n_cols;
  int nf;
  __device__ __forceinline__ int64_t row(int64_t s, int k) const {
    return row0 + (s / nf) * (int64_t)SELL_C * nf + (s % nf) + (int64_t)nf * k;
  }
  // slice of the owned row with local index t (0 <= t < n_slices * 64)
  __device__ __forceinline__ int64_t slice_of(int64_t t) const { return (t / ((int64_t)SELL_C * nf)) * nf + (t % nf); }
};

__device__ __forceinline__ bool csr_has(const int32_t *__restrict__ rp, const int32_t *__restrict__ col, int64_t row,
                                        int32_t c) {
  int32_t lo = rp[row], hi = rp[row + 1];
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    const int32_t v = col[mid];
    if (v == c) return true;
    if (v < c) lo = mid + 1; else hi = mid;
  }
  return false;
}

// pass A: storage mode of every slice (1 = offsets, 0 = explicit columns) from the union of its offsets
__global__ void __launch_bounds__(256) k_sell_mode(const int32_t *__restrict__ rp, const int32_t *__restrict__ col,
                                                   SliceGeo G, uint8_t *__restrict__ sl_mode) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= G.n_slices) return;
  int64_t r[2];
  int32_t p[2], e[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r[h] = G.row(s, 2 * lane + h);
    if (r[h] < G.row1) { p[h] = rp[r[h]]; e[h] = rp[r[h] + 1]; } else { p[h] = e[h] = 0; }
  }
  const int32_t wmax = warp_max(max(e[0] - p[0], e[1] - p[1]));
  int32_t wu = 0;
  while (true) {
    int32_t c0 = p[0] < e[0] ? col[p[0]] - (int32_t)r[0] : SELL_BIG;
    int32_t c1 = p[1] < e[1] ? col[p[1]] - (int32_t)r[1] : SELL_BIG;
    int32_t m = warp_min(min(c0, c1));
    if (m == SELL_BIG) break;
    if (c0 == m) ++p[0];
    if (c1 == m) ++p[1];
    ++wu;
    if (wu > 2 * wmax) break;  // hopeless: explicit mode
  }
  // offset mode pays 8 B per stored entry, explicit mode 12 B: prefer offsets while wu <= 1.5 wmax
  if (lane == 0) sl_mode[s] = ((wmax > 0) && (2 * wu <= 3 * wmax)) ? 1 : 0;
}

// One warp per slice.  PASS 0: count (stored width W, mirrored columns M, columns with offset >= 0, sizes).
// PASS 1: fill offsets / explicit columns / sources / diagonal / the offsets of the mirror table.
// Slices interleave the nf dofs of a node: slice s = b*nf + c holds the 64 rows row0 + b*64*nf + c + nf*k,
// k = 0..63 (in THIS build kernel lane handles k = 2*lane and 2*lane + 1; the value array is indexed [column][k], so the
// SpMV is free to own rows (lane, lane + 32)), i.e. rows of ONE field component -- for vector problems the offsets col - row
// of such rows coincide (3 dn + (c' - c)), which keeps elasticity matrices in offset mode.  nf = 1 gives
// 64 consecutive rows.  A row's CSR columns ascend, hence so do its offsets: the union over the slice is
// produced by repeated warp-wide min extraction.
template <int PASS>
__global__ void __launch_bounds__(256) k_sell_layout(const int32_t *__restrict__ rp, const int32_t *__restrict__ col,
                                                     const int32_t *__restrict__ red2full, SliceGeo G, int sym,
                                                     const uint8_t *__restrict__ sl_mode, int32_t *__restrict__ sl_w,
                                                     int32_t *__restrict__ sl_m, int32_t *__restrict__ sl_u,
                                                     int32_t *__restrict__ sz_val, int32_t *__restrict__ sz_idx,
                                                     const int64_t *__restrict__ valptr, const int64_t *__restrict__ idxptr,
                                                     int32_t *__restrict__ sell_idx, int32_t *__restrict__ sell_src,
                                                     int32_t *__restrict__ sell_diag, uint8_t *__restrict__ ghost,
                                                     int64_t zero_base, unsigned long long *__restrict__ n_mirrored) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= G.n_slices) return;
  int64_t r[2];
  int32_t b[2], e[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r[h] = G.row(s, 2 * lane + h);
    if (r[h] < G.row1) { b[h] = rp[r[h]]; e[h] = rp[r[h] + 1]; } else { b[h] = e[h] = 0; }
  }
  const bool offset_mode = sl_mode[s] != 0;
  if (PASS == 1) {  // does the slice reference a column outside the owned range [row0,row1) (a ghost entry of x)?
    bool g = false;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      for (int32_t q = b[h]; q < e[h]; ++q) g |= (col[q] < G.row0 || col[q] >= G.row1);
    g = __any_sync(0xffffffffu, g);
    if (lane == 0) ghost[s] = g ? 1 : 0;
  }
  if (!offset_mode) {
    const int32_t w = warp_max(max(e[0] - b[0], e[1] - b[1]));
    if (PASS == 0) {
      if (lane == 0) { sl_w[s] = w; sl_m[s] = 0; sl_u[s] = 0; sz_val[s] = w * SELL_C; sz_idx[s] = w * SELL_C; }
      return;
    }
    const int64_t vp = valptr[s], ip = idxptr[s];
    for (int32_t j = 0; j < w; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t pos = vp + (int64_t)j * SELL_C + 2 * lane + h;
        const int64_t ipos = ip + (int64_t)j * SELL_C + 2 * lane + h;
        if (b[h] + j < e[h]) {
          int32_t c = col[b[h] + j];
          sell_idx[ipos] = c;
          sell_src[pos] = red2full[b[h] + j];
          // sell_diag holds the position of the diagonal RELATIVE to the slice's value block
          if (c == (int32_t)r[h]) sell_diag[r[h] - G.row0] = (int32_t)((int64_t)j * SELL_C + 2 * lane + h);
        } else {
          sell_idx[ipos] = r[h] < G.row1 ? (int32_t)r[h] : (int32_t)G.row0;  // padding: any valid column, value 0
          sell_src[pos] = -1;
        }
      }
    }
    return;
  }
  // ---- offset mode: walk the union of offsets in ascending order ------------------------------------------------
  int32_t p[2] = {b[0], b[1]};
  int32_t nU = 0, nL = 0, nM = 0, off_min = 0, off_max = 0;
  int32_t nU_tot = 0, w_tot = 0;
  int64_t vp = 0, ip = 0;
  if (PASS == 1) { nU_tot = sl_u[s]; w_tot = sl_w[s] & 0x7fffffff; vp = valptr[s]; ip = idxptr[s]; }
  const int32_t w_al = (w_tot + 3) & ~3;
  while (true) {
    const int32_t c0 = p[0] < e[0] ? col[p[0]] - (int32_t)r[0] : SELL_BIG;
    const int32_t c1 = p[1] < e[1] ? col[p[1]] - (int32_t)r[1] : SELL_BIG;
    const int32_t m = warp_min(min(c0, c1));
    if (m == SELL_BIG) break;
    const bool has[2] = {c0 == m, c1 == m};
    if (nU + nL + nM == 0) off_min = m;
    off_max = m;
    bool mirrored = false;
    if (m < 0 && sym) {
      bool ok = true;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (r[h] >= G.row1) continue;               // padding row: never written
        const int64_t rho = r[h] + m;               // partner row (m = -d)
        const int64_t t = rho - G.row0;
        const bool in_range = (t >= 0 && t < G.n_rows);
        if (has[h]) {
          if (!in_range) ok = false;                                   // ghost / Dirichlet-clamped column
          else if (!sl_mode[G.slice_of(t)]) ok = false;                 // partner slice stores explicit columns
          else if (!csr_has(rp, col, rho, (int32_t)r[h])) ok = false;   // structurally unsymmetric
        } else if (in_range && csr_has(rp, col, rho, (int32_t)r[h])) {
          ok = false;                                                   // partner has an entry this row lacks
        }
      }
      mirrored = __all_sync(0xffffffffu, ok);
    }
    if (mirrored) {
      if (PASS == 1 && lane == 0) {
        int32_t *tab = sell_idx + ip + w_al + 4 * (int64_t)nM;
        tab[0] = m; tab[1] = 0; tab[2] = 0; tab[3] = 0;                // positions: k_sell_mirror
      }
      ++nM;
    } else {
      const int32_t j = (m >= 0) ? nU : nU_tot + nL;                   // stored column index
      if (PASS == 1) {
        if (lane == 0) sell_idx[ip + j] = m;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int32_t src = -1;
          if (has[h]) {
            src = red2full[p[h]];
            if (m == 0) sell_diag[r[h] - G.row0] = (int32_t)((int64_t)j * SELL_C + 2 * lane + h);
          }
          sell_src[vp + (int64_t)j * SELL_C + 2 * lane + h] = src;
        }
      }
      if (m >= 0) ++nU; else ++nL;
    }
    if (has[0]) ++p[0];
    if (has[1]) ++p[1];
  }
  // the mirror table is padded to a multiple of the SpMV's mirrored batch (7 or 8 columns, whichever pads less;
  // bit 30 of sl_m = batch of 7) with entries that point at the block of zeros, so that the kernel's inner loop
  // has a compile-time trip count and no predicates
  const int32_t pad7 = (nM + 6) / 7 * 7, pad8 = (nM + 7) / 8 * 8;
  const int32_t m_pad = pad7 < pad8 ? pad7 : pad8;
  if (PASS == 0 && lane == 0) {
    const int32_t w = nU + nL;
    sl_w[s] = w | (int32_t)0x80000000;
    // every column of every row inside x: the SpMV may skip the index clamps (all but the first / last mesh planes)
    const int64_t rb = G.row(s, 0);
    const bool interior = w > 0 && rb + off_min >= 0 && rb + (int64_t)G.nf * (SELL_C - 1) + off_max < G.n_cols;
    sl_m[s] = m_pad | (pad7 < pad8 ? SELL_MB7 : 0) | (interior ? SELL_FAST : 0);
    sl_u[s] = nU;
    sz_val[s] = w * SELL_C;
    sz_idx[s] = ((w + 3) & ~3) + 4 * m_pad;   // offsets padded to 16 bytes, then the mirror table (int4 entries)
    if (nM > 0) atomicAdd(n_mirrored, (unsigned long long)nM * SELL_C);
  }
  if (PASS == 1 && lane == 0) {
    for (int32_t j = w_tot; j < w_al; ++j) sell_idx[ip + j] = 0;
    for (int32_t i = nM; i < m_pad; ++i) {
      int32_t *tab = sell_idx + ip + w_al + 4 * (int64_t)i;
      tab[0] = 0; tab[1] = (int32_t)zero_base; tab[2] = (int32_t)zero_base; tab[3] = SELL_C;
    }
  }
}

// pass D: positions of the mirrored columns.  One thread per slice.
__global__ void __launch_bounds__(256) k_sell_mirror(SliceGeo G, const uint8_t *__restrict__ sl_mode,
                                                     const int32_t *__restrict__ sl_w, const int32_t *__restrict__ sl_m,
                                                     const int32_t *__restrict__ sl_u, const int64_t *__restrict__ valptr,
                                                     const int64_t *__restrict__ idxptr, int32_t *__restrict__ sell_idx,
                                                     int64_t zero_base) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= G.n_slices) return;
  const int32_t M = sl_m[s] & SELL_MMASK;
  if (M == 0) return;
  const int32_t w_al = ((sl_w[s] & 0x7fffffff) + 3) & ~3;
  int32_t *tab = sell_idx + idxptr[s] + w_al;
  const int64_t n_blocks = G.n_slices / G.nf;
  for (int32_t i = 0; i < M; ++i) {
    const int32_t d = -tab[4 * i];
    if (d == 0) break;                                         // padding entries (already complete)
    // partner of local row k: owned-row index t_k = t0 + nf k, t0 = row(s,0) - d - row0 (may be negative)
    const int64_t t0 = G.row(s, 0) - d - G.row0;
    int64_t cp = t0 % G.nf; if (cp < 0) cp += G.nf;            // field component of the partner rows
    const int64_t q = (t0 - cp) / G.nf;                        // exact division; partner node index (may be negative)
    int64_t bA = q / SELL_C, dl = q % SELL_C;
    if (dl < 0) { dl += SELL_C; bA -= 1; }                     // floor division
    int32_t pos[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int64_t bb = bA + g;
      int64_t found = -1;
      int64_t sp = -1;
      if (bb >= 0 && bb < n_blocks) {
        sp = bb * G.nf + cp;
        if (sl_mode[sp]) {
          const int32_t *ol = sell_idx + idxptr[sp];           // offsets >= 0 first, ascending
          int32_t lo = 0, hi = sl_u[sp];
          while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            const int32_t v = ol[mid];
            if (v == d) { found = mid; break; }
            if (v < d) lo = mid + 1; else hi = mid;
          }
        }
      }
      pos[g] = found >= 0 ? (int32_t)(valptr[sp] + found * SELL_C + dl - (g ? SELL_C : 0)) : (int32_t)zero_base;
    }
    tab[4 * i + 1] = pos[0];
    tab[4 * i + 2] = pos[1];
    tab[4 * i + 3] = (int32_t)(SELL_C - dl);                   // rows k < split use posA
  }
}

// Gather lists of the assembly scatter in sliced-ELL order, composed once per plan: position t of the value array sums
// ke[gl_idx[j]], j in [gl_ptr[t], gl_ptr[t+1]) -- the element-matrix entries of the CSR entry behind t in ascending COO
// order (the fixed order of k_gather_reduce_*, elements.cu), nothing for padding.  Composing (position -> CSR entry ->
// segment -> COO list) at build time turns three dependent, sector-wasting indirections per position into two
// sequential streams: the P256 scatter moved ~21 GB per assembly that way and moves ~9.5 GB through these lists.
__global__ void k_gl_count(const int32_t *__restrict__ sell_src, const int32_t *__restrict__ seg_ptr, int64_t total,
                           int32_t *__restrict__ cnt) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > total) return;
  int32_t c = 0;
  if (t < total) {
    const int32_t u = sell_src[t];
    if (u >= 0) c = seg_ptr[u + 1] - seg_ptr[u];
  }
  cnt[t] = c;   // cnt[total] = 0: the exclusive scan leaves the grand total there
}
__global__ void k_gl_fill(const int32_t *__restrict__ sell_src, const int32_t *__restrict__ seg_ptr,
                          const uint32_t *__restrict__ perm, const int32_t *__restrict__ gl_ptr, int64_t total,
                          uint32_t *__restrict__ gl_idx) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int32_t u = sell_src[t];
  if (u < 0) return;
  const int32_t b = seg_ptr[u], n = seg_ptr[u + 1] - b, o = gl_ptr[t];
  for (int32_t j = 0; j < n; ++j) gl_idx[o + j] = perm[b + j];
}
__global__ void k_gather_reduce_sell(const double *__restrict__ ke, const uint32_t *__restrict__ gl_idx,
                                     const int32_t *__restrict__ gl_ptr, int64_t total, double *__restrict__ sell_val) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  double acc = 0.0;
  for (int32_t j = gl_ptr[t]; j < gl_ptr[t + 1]; ++j) acc += ke[gl_idx[j]];
  sell_val[t] = acc;
}

static int scan64(const int32_t *in, int64_t *out, int64_t n, cudaStream_t s) {
  size_t tb = 0;
  cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *> it(in, cub::CastOp<int64_t>());
  APDX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, out, n, s));
  DevBuf<uint8_t> tmp;
  APDX_CHECK(tmp.alloc(tb));
  APDX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, out, n, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  tmp.release();
  return APDX_OK;
}

int sell_build(apdx_plan *pl) {
  Sell &S = pl->sell;
  if (S.built) return APDX_OK;
  cudaStream_t s = pl->stream;
  const int64_t rows = pl->f1 - pl->f0;
  S.row0 = pl->f0;
  S.n_rows = rows;
  S.nf = pl->nf;
  S.n_slices = ((rows + (int64_t)SELL_C * S.nf - 1) / ((int64_t)SELL_C * S.nf)) * S.nf;
  {  // symmetric storage is the default; APDX_SELL_SYM=0 keeps every column (A/B measurements)
    const char *e = getenv("APDX_SELL_SYM");
    S.sym = !(e && e[0] == '0');
    // the premise, per set: the element tangent of the model is symmetric.  True for every model of the C ABI today
    // (potential Hessians, Galerkin forms with symmetric constitutive tensors, the capacity matrix; Neumann sets have
    // no tangent); a model that is not listed here switches the whole plan to full storage.
    for (const SetData &st : pl->sets) {
      switch (st.d.model) {
        case APDX_MODEL_POISSON_POTENTIAL: case APDX_MODEL_POISSON_WEAK: case APDX_MODEL_LINEAR_ELASTICITY:
        case APDX_MODEL_NEO_HOOKE: case APDX_MODEL_NEUMANN: case APDX_MODEL_CAPACITY: case APDX_MODEL_PATTERN_ONLY:
          break;
        default:
          S.sym = false;
      }
    }
  }
  const int64_t ns = S.n_slices;
  const SliceGeo G{pl->f0, pl->f1, rows, ns, pl->n_free, S.nf};
  APDX_CHECK(S.sl_w.alloc(ns));
  APDX_CHECK(S.sl_m.alloc(ns));
  APDX_CHECK(S.valptr.alloc(ns + 1));
  APDX_CHECK(S.idxptr.alloc(ns + 1));
  APDX_CHECK(S.diag.alloc(rows));
  APDX_CUDA(cudaMemsetAsync(S.diag.p, 0xff, rows * sizeof(int32_t), s));
  DevBuf<int32_t> szv, szi, sl_u;
  DevBuf<uint8_t> ghost, sl_mode;
  APDX_CHECK(ghost.alloc(ns > 0 ? ns : 1));
  APDX_CHECK(sl_mode.alloc(ns > 0 ? ns : 1));
  APDX_CHECK(sl_u.alloc(ns > 0 ? ns : 1));
  APDX_CHECK(szv.alloc(ns + 1));
  APDX_CHECK(szi.alloc(ns + 1));
  APDX_CUDA(cudaMemsetAsync(szv.p, 0, (ns + 1) * sizeof(int32_t), s));
  APDX_CUDA(cudaMemsetAsync(szi.p, 0, (ns + 1) * sizeof(int32_t), s));
  DevBuf<unsigned long long> n_mir;
  APDX_CHECK(n_mir.alloc(1));
  APDX_CUDA(cudaMemsetAsync(n_mir.p, 0, sizeof(unsigned long long), s));
  const unsigned grid = (unsigned)((ns * 32 + 255) / 256);
  k_sell_mode<<<grid, 256, 0, s>>>(pl->red_row_ptr.p, pl->red_col.p, G, sl_mode.p);
  k_sell_layout<0><<<grid, 256, 0, s>>>(pl->red_row_ptr.p, pl->red_col.p, pl->red2full.p, G, S.sym ? 1 : 0, sl_mode.p,
                                        S.sl_w.p, S.sl_m.p, sl_u.p, szv.p, szi.p, nullptr, nullptr, nullptr, nullptr,
                                        nullptr, nullptr, 0, n_mir.p);
  APDX_CHECK(scan64(szv.p, S.valptr.p, ns + 1, s));
  APDX_CHECK(scan64(szi.p, S.idxptr.p, ns + 1, s));
  int64_t tot[2];
  APDX_CUDA(cudaMemcpy(&tot[0], S.valptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost));
  APDX_CUDA(cudaMemcpy(&tot[1], S.idxptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost));
  S.n_val = tot[0];
  S.n_idx = tot[1];
  APDX_REQUIRE(S.n_val + SELL_C < (1ll << 31), APDX_ERR_UNSUPPORTED, "sliced-ELL storage exceeds 2^31 entries");
  APDX_CHECK(S.val.alloc(S.n_val + SELL_C));   // + the block of zeros mirrored columns without entries point at
  APDX_CUDA(cudaMemsetAsync(S.val.p + S.n_val, 0, SELL_C * sizeof(double), s));
  APDX_CHECK(S.src.alloc(S.n_val > 0 ? S.n_val : 1));
  APDX_CHECK(S.idx.alloc(S.n_idx > 0 ? S.n_idx : 1));
  k_sell_layout<1><<<grid, 256, 0, s>>>(pl->red_row_ptr.p, pl->red_col.p, pl->red2full.p, G, S.sym ? 1 : 0, sl_mode.p,
                                        S.sl_w.p, S.sl_m.p, sl_u.p, nullptr, nullptr, S.valptr.p, S.idxptr.p, S.idx.p,
                                        S.src.p, S.diag.p, ghost.p, S.n_val, n_mir.p);
  if (S.sym && ns > 0)
    k_sell_mirror<<<(unsigned)((ns + 255) / 256), 256, 0, s>>>(G, sl_mode.p, S.sl_w.p, S.sl_m.p, sl_u.p, S.valptr.p,
                                                               S.idxptr.p, S.idx.p, S.n_val);
  APDX_CUDA(cudaStreamSynchronize(s));
  {  // entries read from their transposed position (diagnostics: apdx_plan_sell_info, tests)
    unsigned long long nm = 0;
    APDX_CUDA(cudaMemcpy(&nm, n_mir.p, sizeof(nm), cudaMemcpyDeviceToHost));
    S.n_mirrored = (int64_t)nm;
  }
  if (S.n_val > 0) {  // compose the gather lists of the scatter; the position -> CSR entry map is not needed afterwards
    const unsigned g = (unsigned)((S.n_val + 1 + 255) / 256);
    APDX_CHECK(S.gl_ptr.alloc(S.n_val + 1));
    k_gl_count<<<g, 256, 0, s>>>(S.src.p, pl->seg_ptr.p, S.n_val, S.gl_ptr.p);
    size_t tb = 0;
    APDX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, S.gl_ptr.p, S.gl_ptr.p, S.n_val + 1, s));
    DevBuf<uint8_t> tmp;
    APDX_CHECK(tmp.alloc(tb));
    APDX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, S.gl_ptr.p, S.gl_ptr.p, S.n_val + 1, s));
    int32_t n_gl = 0;
    APDX_CUDA(cudaMemcpyAsync(&n_gl, S.gl_ptr.p + S.n_val, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    APDX_CUDA(cudaStreamSynchronize(s));
    tmp.release();
    APDX_CHECK(S.gl_idx.alloc(n_gl > 0 ? n_gl : 1));
    k_gl_fill<<<g, 256, 0, s>>>(S.src.p, pl->seg_ptr.p, pl->perm.p, S.gl_ptr.p, S.n_val, S.gl_idx.p);
    APDX_CUDA(cudaStreamSynchronize(s));
    S.src.release();
  }
  APDX_CUDA(cudaGetLastError());
  S.built = true;
  return APDX_OK;
}

int sell_gather_reduce(apdx_plan *pl) {
  APDX_CHECK(sell_build(pl));
  Sell &S = pl->sell;
  if (S.n_val > 0) {
    k_gather_reduce_sell<<<(unsigned)((S.n_val + 255) / 256), 256, 0, pl->stream>>>(pl->ke.p, S.gl_idx.p, S.gl_ptr.p, S.n_val,
                                                                                  S.val.p);
    pl->stats.kernel_launches += 1;
  }
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

}  // namespace apdx
