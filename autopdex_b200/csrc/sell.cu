// Sliced-ELL storage of the reduced system for the Krylov solve, built once per plan on device.
//
// Slices of 64 consecutive rows (one warp, two rows per lane -> 128-bit value loads).  Inside a
// slice the entries are stored column-major, so every warp load is one contiguous 512-byte run.
// Column indices are compressed per slice: if the union of (col - row) over the slice's rows is
// small ("offset mode"), the slice stores that list of offsets ONCE (W ints instead of 64*W) and
// entry j of every row means column row + off[j]; rows that do not have an offset hold an explicit
// zero.  Finite-element matrices on structured or well-ordered meshes are almost entirely in
// offset mode, which cuts the per-nonzero traffic from 12 to ~8 bytes and makes the x gathers
// coalesced.  Slices with irregular columns fall back to explicit int32 columns ("explicit mode").
// The reference has no device SpMV at all; this is the storage behind the
// `north_star`'s "sliced-ELL ... with 128-bit loads and warp-shuffle row reductions".
#include <cub/cub.cuh>

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

// One warp per slice.  pass 0: count (width, mode, sizes).  pass 1: fill indices / sources / diagonal.
// Slices interleave the nf dofs of a node: slice s = b*nf + c holds the 64 rows row0 + b*64*nf + c + nf*k,
// k = 0..63 (lane handles k = lane and k = lane + 32, stored side by side as one double2 so that the x gathers
// and y stores of a warp are contiguous), i.e. rows of ONE field component -- for vector problems the offsets col - row
// of such rows coincide (3 dn + (c' - c)), which keeps elasticity matrices in offset mode.  nf = 1 gives
// 64 consecutive rows.  A row's CSR columns ascend, hence so do its offsets: the union over the slice is
// produced by repeated warp-wide min extraction.
template <int PASS>
__global__ void __launch_bounds__(256) k_sell_build(const int32_t *__restrict__ rp, const int32_t *__restrict__ col,
                                                    const int32_t *__restrict__ red2full, int64_t row0, int64_t row1,
                                                    int64_t n_slices, int nf, int32_t *__restrict__ sl_w,
                                                    int32_t *__restrict__ sz_val, int32_t *__restrict__ sz_idx,
                                                    const int64_t *__restrict__ valptr, const int64_t *__restrict__ idxptr,
                                                    int32_t *__restrict__ sell_idx, int32_t *__restrict__ sell_src,
                                                    int32_t *__restrict__ sell_diag, uint8_t *__restrict__ ghost) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= n_slices) return;
  int64_t r[2];
  int32_t b[2], e[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r[h] = row0 + (s / nf) * (int64_t)SELL_C * nf + (s % nf) + (int64_t)nf * (2 * lane + h);
    if (r[h] < row1) { b[h] = rp[r[h]]; e[h] = rp[r[h] + 1]; } else { b[h] = e[h] = 0; }
  }
  const int32_t wmax = warp_max(max(e[0] - b[0], e[1] - b[1]));
  // ---- union of offsets -------------------------------------------------------------------------
  int32_t p[2] = {b[0], b[1]};
  int32_t wu = 0;
  bool offset_mode;
  if (PASS == 0) {
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
    offset_mode = (wmax > 0) && (2 * wu <= 3 * wmax);
    if (lane == 0) {
      int32_t w = offset_mode ? wu : wmax;
      sl_w[s] = w | (offset_mode ? (int32_t)0x80000000 : 0);
      sz_val[s] = w * SELL_C;
      sz_idx[s] = offset_mode ? ((w + 1) & ~1) : w * SELL_C;  // even: explicit slices read their columns as int2
    }
    return;
  }
  // ---- PASS 1: fill ------------------------------------------------------------------------------
  {  // does the slice reference a column outside the owned range [row0,row1) (a ghost entry of x)?
    bool g = false;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      for (int32_t q = b[h]; q < e[h]; ++q) g |= (col[q] < row0 || col[q] >= row1);
    g = __any_sync(0xffffffffu, g);
    if (lane == 0) ghost[s] = g ? 1 : 0;
  }
  const int32_t wenc = sl_w[s];
  offset_mode = wenc < 0;
  const int32_t w = wenc & 0x7fffffff;
  const int64_t vp = valptr[s], ip = idxptr[s];
  if (offset_mode) {
    for (int32_t j = 0; j < w; ++j) {
      int32_t c0 = p[0] < e[0] ? col[p[0]] - (int32_t)r[0] : SELL_BIG;
      int32_t c1 = p[1] < e[1] ? col[p[1]] - (int32_t)r[1] : SELL_BIG;
      int32_t m = warp_min(min(c0, c1));
      if (lane == 0) sell_idx[ip + j] = m;
      int32_t s0 = -1, s1 = -1;
      // sell_diag holds the position of the diagonal RELATIVE to the slice's value block
      if (c0 == m) { s0 = red2full[p[0]]; if (m == 0) sell_diag[r[0] - row0] = (int32_t)((int64_t)j * SELL_C + 2 * lane); ++p[0]; }
      if (c1 == m) { s1 = red2full[p[1]]; if (m == 0) sell_diag[r[1] - row0] = (int32_t)((int64_t)j * SELL_C + 2 * lane + 1); ++p[1]; }
      sell_src[vp + (int64_t)j * SELL_C + 2 * lane] = s0;
      sell_src[vp + (int64_t)j * SELL_C + 2 * lane + 1] = s1;
    }
  } else {
    for (int32_t j = 0; j < w; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t pos = vp + (int64_t)j * SELL_C + 2 * lane + h;
        const int64_t ipos = ip + (int64_t)j * SELL_C + 2 * lane + h;
        if (b[h] + j < e[h]) {
          int32_t c = col[b[h] + j];
          sell_idx[ipos] = c;
          sell_src[pos] = red2full[b[h] + j];
          if (c == (int32_t)r[h]) sell_diag[r[h] - row0] = (int32_t)((int64_t)j * SELL_C + 2 * lane + h);
        } else {
          sell_idx[ipos] = r[h] < row1 ? (int32_t)r[h] : (int32_t)row0;  // padding: any valid column, value 0
          sell_src[pos] = -1;
        }
      }
    }
  }
}

// sell_val[t] = sum of the element-matrix entries of the CSR entry behind SELL position t (0 for padding),
// fixed ascending COO order as in k_gather_reduce_* (elements.cu)
__global__ void k_gather_reduce_sell(const double *__restrict__ ke, const uint32_t *__restrict__ perm,
                                     const int32_t *__restrict__ seg_ptr, const int32_t *__restrict__ sell_src,
                                     int64_t total, double *__restrict__ sell_val) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int32_t u = sell_src[t];
  double acc = 0.0;
  if (u >= 0)
    for (int32_t j = seg_ptr[u]; j < seg_ptr[u + 1]; ++j) acc += ke[perm[j]];
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
  const int64_t ns = S.n_slices;
  APDX_CHECK(S.sl_w.alloc(ns));
  APDX_CHECK(S.valptr.alloc(ns + 1));
  APDX_CHECK(S.idxptr.alloc(ns + 1));
  APDX_CHECK(S.diag.alloc(rows));
  APDX_CUDA(cudaMemsetAsync(S.diag.p, 0xff, rows * sizeof(int32_t), s));
  DevBuf<int32_t> szv, szi;
  DevBuf<uint8_t> ghost;
  APDX_CHECK(ghost.alloc(ns > 0 ? ns : 1));
  APDX_CHECK(szv.alloc(ns + 1));
  APDX_CHECK(szi.alloc(ns + 1));
  APDX_CUDA(cudaMemsetAsync(szv.p, 0, (ns + 1) * sizeof(int32_t), s));
  APDX_CUDA(cudaMemsetAsync(szi.p, 0, (ns + 1) * sizeof(int32_t), s));
  const unsigned grid = (unsigned)((ns * 32 + 255) / 256);
  k_sell_build<0><<<grid, 256, 0, s>>>(pl->red_row_ptr.p, pl->red_col.p, pl->red2full.p, pl->f0, pl->f1, ns, S.nf, S.sl_w.p,
                                       szv.p, szi.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  APDX_CHECK(scan64(szv.p, S.valptr.p, ns + 1, s));
  APDX_CHECK(scan64(szi.p, S.idxptr.p, ns + 1, s));
  int64_t tot[2];
  APDX_CUDA(cudaMemcpy(&tot[0], S.valptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost));
  APDX_CUDA(cudaMemcpy(&tot[1], S.idxptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost));
  S.n_val = tot[0];
  S.n_idx = tot[1];
  APDX_REQUIRE(S.n_val < (1ll << 31), APDX_ERR_UNSUPPORTED, "sliced-ELL storage exceeds 2^31 entries");
  APDX_CHECK(S.val.alloc(S.n_val > 0 ? S.n_val : 1));
  APDX_CHECK(S.src.alloc(S.n_val > 0 ? S.n_val : 1));
  APDX_CHECK(S.idx.alloc(S.n_idx > 0 ? S.n_idx : 1));
  k_sell_build<1><<<grid, 256, 0, s>>>(pl->red_row_ptr.p, pl->red_col.p, pl->red2full.p, pl->f0, pl->f1, ns, S.nf, S.sl_w.p,
                                       nullptr, nullptr, S.valptr.p, S.idxptr.p, S.idx.p, S.src.p, S.diag.p, ghost.p);
  APDX_CUDA(cudaStreamSynchronize(s));
  {  // leading / trailing runs of slices that touch ghost columns; everything in between is "interior"
    std::vector<uint8_t> gh((size_t)ns);
    APDX_CUDA(cudaMemcpy(gh.data(), ghost.p, (size_t)ns, cudaMemcpyDeviceToHost));
    int64_t first_clean = 0, last_clean = ns;
    while (first_clean < ns && gh[first_clean]) ++first_clean;
    while (last_clean > first_clean && gh[last_clean - 1]) --last_clean;
    bool clean = true;
    for (int64_t q = first_clean; q < last_clean; ++q) clean = clean && !gh[q];
    // pad the runs a little so that flagged slices scattered just behind the ghost planes stay in the boundary part
    if (!clean) {
      int64_t lo = first_clean, hi = last_clean;
      for (int64_t q = first_clean; q < last_clean; ++q)
        if (gh[q]) { if (q < ns / 2) lo = q + 1; else { hi = q; break; } }
      first_clean = lo;
      last_clean = hi > lo ? hi : lo;
      clean = true;
      for (int64_t q = first_clean; q < last_clean; ++q) clean = clean && !gh[q];
      if (!clean) { first_clean = ns; last_clean = ns; }   // irregular partition: no overlap, one launch
    }
    S.lo_end = first_clean;
    S.hi_begin = last_clean;
  }
  APDX_CUDA(cudaGetLastError());
  S.built = true;
  return APDX_OK;
}

int sell_gather_reduce(apdx_plan *pl) {
  APDX_CHECK(sell_build(pl));
  Sell &S = pl->sell;
  if (S.n_val > 0) {
    k_gather_reduce_sell<<<(unsigned)((S.n_val + 255) / 256), 256, 0, pl->stream>>>(pl->ke.p, pl->perm.p, pl->seg_ptr.p,
                                                                                  S.src.p, S.n_val, S.val.p);
    pl->stats.kernel_launches += 1;
  }
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

}  // namespace apdx
