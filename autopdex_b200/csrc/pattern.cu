// Sparsity pattern, element-local -> CSR map and gather lists, built once on device by
// radix sort + unique.  Device replacement of assembler._get_indices (assembler.py:47-141)
// and of the pattern half of solver.scipy_assembling (solver.py:1207-1217): rows ascending,
// columns ascending inside a row, duplicates merged, explicit zeros kept, and the reduced
// system csr[:, free][free] with free dofs renumbered in order.
#include <cub/cub.cuh>

#include "common.cuh"

namespace apdx {

static inline unsigned grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  return (unsigned)g;
}

// COO key of entry k of one set: (row << bits) | col, reference order
// k = e*ndof^2 + i*ndof + j, row = gd[i], col = gd[j], gd = conn[e][a]*nf + c  (assembler.py:129-141)
__global__ void k_coo_keys(const int32_t *__restrict__ conn, int64_t n_entries, int nen, int nf,
                           int bits, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                           int64_t offset) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_entries) return;
  int ndof = nen * nf;
  int64_t e = k / ((int64_t)ndof * ndof);
  int rem = (int)(k - e * (int64_t)ndof * ndof);
  int i = rem / ndof, j = rem - i * ndof;
  uint64_t row = (uint64_t)conn[e * nen + i / nf] * nf + (i % nf);
  uint64_t col = (uint64_t)conn[e * nen + j / nf] * nf + (j % nf);
  keys[offset + k] = (row << bits) | col;
  vals[offset + k] = (uint32_t)(offset + k);
}

// key of entry m of the element-vector stream: the global dof it adds into (assembler.py:415-431)
__global__ void k_res_keys(const int32_t *__restrict__ conn, int64_t n_entries, int nen, int nf,
                           uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int64_t offset) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_entries) return;
  int ndof = nen * nf;
  int64_t e = m / ndof;
  int i = (int)(m - e * ndof);
  keys[offset + m] = (uint32_t)conn[e * nen + i / nf] * nf + (i % nf);
  vals[offset + m] = (uint32_t)(offset + m);
}

// The element kernels store element matrices / vectors entry-major ("SoA": ke[set][ij][element]) so that both their
// stores and the gathers of the segmented reduction are coalesced across consecutive elements.  The sort works on
// reference-order COO indices k = e*nd2 + ij (needed for the element->CSR map); afterwards the gather lists are
// rewritten to SoA addresses off + ij*n_rows + e.
struct SoaSet {
  int64_t off, koff, n_rows;   // off: first reference-order index of the set; koff: start of its block in the stream
  int32_t width;   // ndof^2 (matrix stream) or ndof (vector stream)
  int32_t sym_n;   // matrix stream stored as the upper triangle of every element matrix (sym_n = ndof; entry (a,b) and
                   // (b,a) share slot lo*n - lo(lo-1)/2 + hi - lo): entry-major for register-kernel sets (n_rows > 0),
                   // element-major for generic-kernel sets (n_rows == 0); 0 = full storage
};
struct SoaTable {
  int n;
  SoaSet s[16];
};
// stream address of the reference-order COO (or element-vector) index k
__device__ __forceinline__ int64_t soa_address(int64_t k, const SoaTable &t) {
  int q = 0;
  while (q + 1 < t.n && k >= t.s[q + 1].off) ++q;
  const int64_t local = k - t.s[q].off;
  if (t.s[q].n_rows == 0 && t.s[q].sym_n == 0) return t.s[q].koff + local;   // reference (element-major, full) layout
  const int64_t e = local / t.s[q].width;
  int64_t ij = local - e * t.s[q].width;
  if (t.s[q].sym_n > 0) {
    const int nn = t.s[q].sym_n, a = (int)(ij / nn), b = (int)(ij - (int64_t)a * nn);
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    ij = lo * nn - lo * (lo - 1) / 2 + (hi - lo);
    if (t.s[q].n_rows == 0) return t.s[q].koff + e * ((int64_t)nn * (nn + 1) / 2) + ij;   // element-major triangle
  }
  return t.s[q].koff + ij * t.s[q].n_rows + e;
}
__global__ void k_to_soa(uint32_t *__restrict__ list, int64_t n, SoaTable t) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  list[i] = (uint32_t)soa_address(list[i], t);
}
// element-matrix entries [k0, k0 + n) in the reference's COO order (assembler.py:1383: data = tangent_contributions.flatten())
__global__ void k_coo_export(const double *__restrict__ ke, int64_t k0, int64_t n, SoaTable t, double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = ke[soa_address(k0 + i, t)];
}

static SoaTable matrix_soa_table(const apdx_plan *pl) {
  SoaTable t{};
  for (auto &st : pl->sets) {
    if (st.d.n_rows == 0) continue;
    t.s[t.n++] = SoaSet{st.coo_offset, st.ke_offset, st.soa ? st.d.n_rows : 0, st.ndof_e * st.ndof_e,
                        (st.soa || st.tri) ? st.ndof_e : 0};
  }
  return t;
}

template <typename K>
__global__ void k_head_flags(const K *__restrict__ sorted, int64_t n, int32_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
}

// uid = inclusive scan of the head flags (1-based); writes segment starts, columns and the map
__global__ void k_unique_fill(const uint64_t *__restrict__ sorted, const uint32_t *__restrict__ perm,
                              const int32_t *__restrict__ uid_incl, int64_t n, int bits,
                              int32_t *__restrict__ seg_ptr, int32_t *__restrict__ col) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t u = uid_incl[i] - 1;
  bool head = (i == 0) || (uid_incl[i - 1] != uid_incl[i]);
  if (head) {
    seg_ptr[u] = (int32_t)i;
    col[u] = (int32_t)(sorted[i] & ((1ull << bits) - 1ull));
  }
}

__global__ void k_res_fill(const uint32_t *__restrict__ sorted, int64_t n, int64_t n_dofs,
                           int32_t *__restrict__ rseg_ptr) {
  // rseg_ptr[d] = first index i with sorted[i] >= d
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t cur = sorted[i];
  int64_t prev = (i == 0) ? -1 : (int64_t)sorted[i - 1];
  for (int64_t d = prev + 1; d <= cur; ++d) rseg_ptr[d] = (int32_t)i;
  if (i == n - 1)
    for (int64_t d = cur + 1; d <= n_dofs; ++d) rseg_ptr[d] = (int32_t)n;
}

// row_ptr from the sorted unique keys: row_ptr[r] = first unique entry with row >= r
__global__ void k_row_ptr(const uint64_t *__restrict__ sorted, const int32_t *__restrict__ seg_ptr,
                          int64_t nnz, int bits, int64_t n_rows, int32_t *__restrict__ row_ptr) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nnz) return;
  int64_t cur = (int64_t)(sorted[seg_ptr[u]] >> bits);
  int64_t prev = (u == 0) ? -1 : (int64_t)(sorted[seg_ptr[u - 1]] >> bits);
  for (int64_t r = prev + 1; r <= cur; ++r) row_ptr[r] = (int32_t)u;
  if (u == nnz - 1)
    for (int64_t r = cur + 1; r <= n_rows; ++r) row_ptr[r] = (int32_t)nnz;
}

__global__ void k_free_flags(const uint8_t *__restrict__ mask, int64_t n, int32_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = mask[i] ? 0 : 1;
}
__global__ void k_free_fill(const uint8_t *__restrict__ mask, const int32_t *__restrict__ excl, int64_t n,
                            int32_t *__restrict__ free_id, int32_t *__restrict__ free_list) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask[i]) {
    free_id[i] = -1;
  } else {
    free_id[i] = excl[i];
    free_list[excl[i]] = (int32_t)i;
  }
}

// keep flag of every full CSR entry: row and column both free (solver.py:1214-1217)
__global__ void k_keep_flags(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                             const int32_t *__restrict__ free_id, int64_t n_rows,
                             int32_t *__restrict__ keep) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  bool rf = free_id[r] >= 0;
  for (int32_t u = row_ptr[r]; u < row_ptr[r + 1]; ++u) keep[u] = (rf && free_id[col[u]] >= 0) ? 1 : 0;
}
__global__ void k_reduced_fill(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                               const int32_t *__restrict__ free_id, const int32_t *__restrict__ keep_excl,
                               int64_t n_rows, int64_t nnz_red, int64_t n_free,
                               int32_t *__restrict__ red_row_ptr, int32_t *__restrict__ red_col,
                               int32_t *__restrict__ red2full, int32_t *__restrict__ red_diag) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int32_t q = free_id[r];
  if (q < 0) return;
  int32_t b = row_ptr[r], e = row_ptr[r + 1];
  int32_t o = keep_excl[b];
  red_row_ptr[q] = o;
  if (q == n_free - 1) red_row_ptr[n_free] = (int32_t)nnz_red;
  int32_t diag = -1;
  for (int32_t u = b; u < e; ++u) {
    int32_t c = free_id[col[u]];
    if (c >= 0) {
      red_col[o] = c;
      red2full[o] = u;
      if (c == q) diag = o;
      ++o;
    }
  }
  red_diag[q] = diag;
}

template <typename T>
static int exclusive_scan(const int32_t *in, T *out, int64_t n, cudaStream_t s) {
  size_t tb = 0;
  APDX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s));
  DevBuf<uint8_t> tmp;
  APDX_CHECK(tmp.alloc(tb));
  APDX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  tmp.release();
  return APDX_OK;
}
// out[i] = in[0] + ... + in[i-1] for i < n (used by multigrid.cu for the row pointers of its transfer operators);
// returns after the stream has finished
int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s) { return exclusive_scan<int32_t>(in, out, n, s); }
static int inclusive_scan(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s) {
  size_t tb = 0;
  APDX_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, in, out, n, s));
  DevBuf<uint8_t> tmp;
  APDX_CHECK(tmp.alloc(tb));
  APDX_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, in, out, n, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  tmp.release();
  return APDX_OK;
}

static int bits_for(int64_t n) {
  int b = 1;
  while ((1ll << b) < n) ++b;
  return b;
}

int build_pattern(apdx_plan *pl, const uint8_t *mask_h) {
  cudaStream_t s = pl->stream;
  const int64_t n = pl->n_dofs;
  const int B = 256;

  // ---- Dirichlet maps ----------------------------------------------------------------
  APDX_CHECK(pl->mask.alloc(n));
  if (mask_h)
    APDX_CUDA(cudaMemcpyAsync(pl->mask.p, mask_h, n, cudaMemcpyHostToDevice, s));
  else
    APDX_CUDA(cudaMemsetAsync(pl->mask.p, 0, n, s));
  APDX_CHECK(pl->free_id.alloc(n));
  {
    DevBuf<int32_t> flag, excl;
    APDX_CHECK(flag.alloc(n + 1));
    APDX_CHECK(excl.alloc(n + 1));
    APDX_CUDA(cudaMemsetAsync(flag.p, 0, (n + 1) * sizeof(int32_t), s));
    k_free_flags<<<grid_for(n, B), B, 0, s>>>(pl->mask.p, n, flag.p);
    APDX_CHECK(exclusive_scan(flag.p, excl.p, n + 1, s));
    int32_t nf_ = 0;
    APDX_CUDA(cudaMemcpy(&nf_, excl.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost));
    pl->n_free = nf_;
    APDX_CHECK(pl->free_list.alloc(pl->n_free > 0 ? pl->n_free : 1));
    k_free_fill<<<grid_for(n, B), B, 0, s>>>(pl->mask.p, excl.p, n, pl->free_id.p, pl->free_list.p);
    APDX_CUDA(cudaStreamSynchronize(s));
    flag.release();
    excl.release();
  }
  APDX_REQUIRE(pl->n_free > 0, APDX_ERR_INVALID, "all dofs are Dirichlet dofs: nothing to solve");

  // ---- residual gather lists -----------------------------------------------------------
  {
    const int64_t m = pl->n_res;
    DevBuf<uint32_t> kin, kout, vin;
    APDX_CHECK(kin.alloc(m));
    APDX_CHECK(kout.alloc(m));
    APDX_CHECK(vin.alloc(m));
    APDX_CHECK(pl->rperm.alloc(m));
    for (auto &st : pl->sets) {
      int64_t cnt = st.d.n_rows * st.ndof_e;
      if (cnt == 0) continue;
      k_res_keys<<<grid_for(cnt, B), B, 0, s>>>(st.conn.p, cnt, st.d.nen, pl->nf, kin.p, vin.p, st.res_offset);
    }
    size_t tb = 0;
    int eb = bits_for(n);
    APDX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin.p, kout.p, vin.p, pl->rperm.p, m, 0, eb, s));
    DevBuf<uint8_t> tmp;
    APDX_CHECK(tmp.alloc(tb));
    APDX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kin.p, kout.p, vin.p, pl->rperm.p, m, 0, eb, s));
    APDX_CHECK(pl->rseg_ptr.alloc(n + 1));
    k_res_fill<<<grid_for(m, B), B, 0, s>>>(kout.p, m, n, pl->rseg_ptr.p);
    SoaTable t{};
    for (auto &st : pl->sets) {
      if (st.d.n_rows == 0) continue;
      t.s[t.n++] = SoaSet{st.res_offset, st.res_offset, st.soa ? st.d.n_rows : 0, st.ndof_e, 0};
    }
    k_to_soa<<<grid_for(m, B), B, 0, s>>>(pl->rperm.p, m, t);
    APDX_CUDA(cudaStreamSynchronize(s));
  }

  // ---- COO sort / unique ----------------------------------------------------------------
  const int64_t nc = pl->n_coo;
  APDX_REQUIRE(nc < (1ll << 31), APDX_ERR_UNSUPPORTED,
               "%lld COO entries exceed the 2^31 limit of one plan; partition the mesh", (long long)nc);
  const int bits = bits_for(n);
  DevBuf<uint64_t> sorted;
  APDX_CHECK(pl->perm.alloc(nc));
  APDX_CHECK(sorted.alloc(nc));
  {
    DevBuf<uint64_t> kin;
    DevBuf<uint32_t> vin;
    APDX_CHECK(kin.alloc(nc));
    APDX_CHECK(vin.alloc(nc));
    for (auto &st : pl->sets) {
      int64_t cnt = st.d.n_rows * (int64_t)st.ndof_e * st.ndof_e;
      if (cnt == 0) continue;
      k_coo_keys<<<grid_for(cnt, B), B, 0, s>>>(st.conn.p, cnt, st.d.nen, pl->nf, bits, kin.p, vin.p,
                                                 st.coo_offset);
    }
    size_t tb = 0;
    APDX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin.p, sorted.p, vin.p, pl->perm.p, nc, 0,
                                              2 * bits, s));
    DevBuf<uint8_t> tmp;
    APDX_CHECK(tmp.alloc(tb));
    APDX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kin.p, sorted.p, vin.p, pl->perm.p, nc, 0,
                                              2 * bits, s));
    APDX_CUDA(cudaStreamSynchronize(s));
  }
  {
    DevBuf<int32_t> uid;  // head flags, then their inclusive scan
    APDX_CHECK(uid.alloc(nc));
    k_head_flags<uint64_t><<<grid_for(nc, B), B, 0, s>>>(sorted.p, nc, uid.p);
    APDX_CHECK(inclusive_scan(uid.p, uid.p, nc, s));
    int32_t nnz = 0;
    APDX_CUDA(cudaMemcpy(&nnz, uid.p + nc - 1, sizeof(int32_t), cudaMemcpyDeviceToHost));
    pl->nnz = nnz;
    APDX_CHECK(pl->seg_ptr.alloc(pl->nnz + 1));
    APDX_CHECK(pl->col.alloc(pl->nnz));
    k_unique_fill<<<grid_for(nc, B), B, 0, s>>>(sorted.p, pl->perm.p, uid.p, nc, bits, pl->seg_ptr.p,
                                                 pl->col.p);
    int32_t nc32 = (int32_t)nc;
    APDX_CUDA(cudaMemcpyAsync(pl->seg_ptr.p + pl->nnz, &nc32, sizeof(int32_t), cudaMemcpyHostToDevice, s));
    APDX_CHECK(pl->row_ptr.alloc(n + 1));
    k_row_ptr<<<grid_for(pl->nnz, B), B, 0, s>>>(sorted.p, pl->seg_ptr.p, pl->nnz, bits, n, pl->row_ptr.p);
    const SoaTable t = matrix_soa_table(pl);
    k_to_soa<<<grid_for(nc, B), B, 0, s>>>(pl->perm.p, nc, t);   // after k_unique_fill, which needs reference-order indices
    APDX_CUDA(cudaStreamSynchronize(s));
  }
  sorted.release();

  // ---- reduced pattern -------------------------------------------------------------------
  {
    DevBuf<int32_t> keep, excl;
    APDX_CHECK(keep.alloc(pl->nnz + 1));
    APDX_CHECK(excl.alloc(pl->nnz + 1));
    APDX_CUDA(cudaMemsetAsync(keep.p + pl->nnz, 0, sizeof(int32_t), s));
    k_keep_flags<<<grid_for(n, B), B, 0, s>>>(pl->row_ptr.p, pl->col.p, pl->free_id.p, n, keep.p);
    APDX_CHECK(exclusive_scan(keep.p, excl.p, pl->nnz + 1, s));
    int32_t nr = 0;
    APDX_CUDA(cudaMemcpy(&nr, excl.p + pl->nnz, sizeof(int32_t), cudaMemcpyDeviceToHost));
    pl->nnz_red = nr;
    APDX_CHECK(pl->red_row_ptr.alloc(pl->n_free + 1));
    APDX_CHECK(pl->red_col.alloc(pl->nnz_red > 0 ? pl->nnz_red : 1));
    APDX_CHECK(pl->red2full.alloc(pl->nnz_red > 0 ? pl->nnz_red : 1));
    APDX_CHECK(pl->red_diag.alloc(pl->n_free));
    k_reduced_fill<<<grid_for(n, B), B, 0, s>>>(pl->row_ptr.p, pl->col.p, pl->free_id.p, excl.p, n,
                                                 pl->nnz_red, pl->n_free, pl->red_row_ptr.p,
                                                 pl->red_col.p, pl->red2full.p, pl->red_diag.p);
    APDX_CUDA(cudaStreamSynchronize(s));
  }
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

// The element-local -> CSR position map (the north-star "index map") for the COO entries
// [k0, k0 + n): position of (row, col) of entry k inside the sorted, duplicate-free pattern that the radix sort + unique
// pass built.  The scatter itself runs through the INVERSE lists (perm / seg_ptr, gather form); the forward map is
// only exported (parity tests, apdx_plan_get_elem_map), so it is materialised on request instead of holding
// 4 bytes per COO entry (4.3 GB at 256^3).
struct MapSet {
  int64_t off;
  const int32_t *conn;
  int32_t nen, ndof;
};
struct MapTable {
  int n, nf;
  MapSet s[16];
};
__global__ void k_elem_map_range(MapTable t, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, int64_t k0,
                                 int64_t n, int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = k0 + i;
  int q = 0;
  while (q + 1 < t.n && k >= t.s[q + 1].off) ++q;
  const MapSet &S = t.s[q];
  const int64_t local = k - S.off, nd2 = (int64_t)S.ndof * S.ndof;
  const int64_t e = local / nd2;
  const int rem = (int)(local - e * nd2), a = rem / S.ndof, b = rem - a * S.ndof;
  const int32_t row = S.conn[e * S.nen + a / t.nf] * t.nf + a % t.nf;
  const int32_t c = S.conn[e * S.nen + b / t.nf] * t.nf + b % t.nf;
  int32_t lo = row_ptr[row], hi = row_ptr[row + 1];
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  out[i] = lo;
}
int elem_map_export(const apdx_plan *pl, int64_t offset, int64_t count, int32_t *dst_d) {
  if (count <= 0) return APDX_OK;
  MapTable t{};
  t.nf = pl->nf;
  for (auto &st : pl->sets) {
    if (st.d.n_rows == 0) continue;
    APDX_REQUIRE(t.n < 16, APDX_ERR_UNSUPPORTED, "more than 16 non-empty sets");
    t.s[t.n++] = MapSet{st.coo_offset, st.conn.p, st.d.nen, st.ndof_e};
  }
  k_elem_map_range<<<grid_for(count, 256), 256, 0, pl->stream>>>(t, pl->row_ptr.p, pl->col.p, offset, count, dst_d);
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

// The BCOO wire format of assembler.assemble_tangent (assembler.py:749-777): values of the element-local pairs
// [offset, offset + count) in reference order, duplicates NOT summed, copied out of the element streams.
int coo_export(apdx_plan *pl, int64_t offset, int64_t count, double *dst_d) {
  if (count <= 0) return APDX_OK;
  k_coo_export<<<grid_for(count, 256), 256, 0, pl->stream>>>(pl->ke.p, offset, count, matrix_soa_table(pl), dst_d);
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

}  // namespace apdx
