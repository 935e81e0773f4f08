// Internal declarations shared by the translation units of libapdx_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/apdx_b200.h"

namespace apdx {

void set_error(const char *fmt, ...);

#define APDX_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      apdx::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                   \
                      cudaGetErrorString(err__));                                          \
      return APDX_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define APDX_CHECK(call)                                                                   \
  do {                                                                                     \
    int rc__ = (call);                                                                     \
    if (rc__ != APDX_OK) return rc__;                                                      \
  } while (0)

#define APDX_REQUIRE(cond, code, ...)                                                      \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      apdx::set_error(__VA_ARGS__);                                                        \
      return (code);                                                                       \
    }                                                                                      \
  } while (0)

// Device buffer with size bookkeeping (all plan memory goes through this).
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool owned = true;
  DevBuf() = default;
  // Scope-bound: a buffer that goes out of scope frees its memory (the build passes use many temporaries; before this
  // destructor existed the ones without an explicit release() leaked ~30 GB per 256^3 plan).  Move-only.
  ~DevBuf() { release(); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; o.owned = true; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; o.owned = true; }
    return *this;
  }
  int alloc(size_t count);
  void release();
  void adopt(T *ptr, size_t count) {  // non-owning alias (vectors living in the peer-to-peer heap)
    release();
    p = ptr;
    n = count;
    owned = false;
  }
  size_t bytes() const { return n * sizeof(T); }
};

extern size_t g_plan_bytes;  // bytes currently held by DevBufs (diagnostics)

template <typename T>
int DevBuf<T>::alloc(size_t count) {
  release();
  if (count == 0) return APDX_OK;
  cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
  if (e != cudaSuccess) {
    p = nullptr;
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? APDX_ERR_NOMEM : APDX_ERR_CUDA;
  }
  n = count;
  g_plan_bytes += bytes();
  return APDX_OK;
}
template <typename T>
void DevBuf<T>::release() {
  if (p && owned) {
    cudaFree(p);
    g_plan_bytes -= bytes();
  }
  p = nullptr;
  n = 0;
  owned = true;
}

// Run-time parameter of a set: value(row, gp, comp) = p[row*s_row + gp*s_gp + comp]
struct ParamView {
  const double *p;
  int64_t s_row;
  int32_t s_gp;
  int32_t ncomp;
};

struct SetData {
  apdx_set_desc d{};          // host copy of the descriptor (pointers invalid after create)
  int32_t ndof_e = 0;         // nen * nf
  int64_t coo_offset = 0;     // offset of this set's block in the reference-order COO numbering
  int64_t ke_offset = 0;      // offset of this set's block in the element-matrix stream `ke` (upper triangle only for soa / tri sets)
  int64_t res_offset = 0;     // offset of this set's block in the Re stream
  bool soa = false;           // element streams stored entry-major (sets handled by the register kernels)
  bool tri = false;           // generic-kernel set: element matrices stored as their upper triangle, element-major
                              // (ke[e * ndof (ndof + 1) / 2 + slot(i, j)]; every model of the kernel has a symmetric tangent)
  DevBuf<int32_t> conn;       // [n_rows][nen]
  DevBuf<double> shape_n, shape_dn, gp_w;
  std::vector<double> h_shape_n, h_shape_dn, h_gp_w;  // host copies (kernel-argument tables of the fast kernels)
  DevBuf<double> ip_n, ip_dndx, ip_w;  // integration-point tables
  DevBuf<double> params[APDX_PARAM_COUNT];
  ParamView pview[APDX_PARAM_COUNT]{};
};

constexpr int NORM_GRID = 592;   // blocks of the residual-norm reduction (api.cu); KrylovWork::partial holds its partial sums + the result

struct KrylovWork {
  DevBuf<double> r, p, q, s, t, phat, shat, r0, minv;
  DevBuf<double> partial;     // per-block partial sums of the fused dot products
  DevBuf<double> scal;        // device scalars (see krylov.cu)
  DevBuf<unsigned int> ticket;
  DevBuf<int32_t> flags;      // [0]=done [1]=iterations [2]=breakdown [3]=maxiter [4]=final
};

// ---- peer-memory (NVLink) plumbing of the multi-GPU Krylov loop (dist.cu, krylov.cu) ------------------
constexpr int P2P_MAX_RANKS = 16;
// device-visible descriptor; every pointer is dereferenceable from this GPU (peer memory mapped through CUDA IPC)
struct P2PDev {
  int rank, nranks;
  double *mbox[P2P_MAX_RANKS];   // rank r's mailboxes: double [2 slots][P2P_MAX_RANKS senders][4]
  int *mflag[P2P_MAX_RANKS];     // rank r's mailbox flags: int [2 slots][P2P_MAX_RANKS senders]
  int *err;                      // my time-out flag
  int *epoch_self;               // device-resident reduction counter of the mailbox all-reduce
  // halo inboxes of a slab partition (dist.cu: k_halo_push / k_halo_pull): a rank stores its boundary entries straight
  // into the neighbour's inbox over NVLink and raises a flag there; [2 slots][cap] doubles per side, slot = parity of
  // the exchange counter
  double *peer_inbox_lo, *peer_inbox_hi;   // rank_lo's upper / rank_hi's lower inbox (nullptr: no neighbour)
  int *peer_flag_lo, *peer_flag_hi;        // the flag pair [2 slots] that belongs to those inboxes
  double *my_inbox_lo, *my_inbox_hi;       // where rank_lo / rank_hi deposit my ghost entries
  int *my_flag_lo, *my_flag_hi;
  int *halo_epoch;                         // device-resident exchange counter
  unsigned int *halo_ticket;               // [2] last-block tickets of the push and the pull kernel
  long long cap;                           // doubles per inbox slot
};
struct P2P {
  bool mbox = false;             // the dot-product all-reduces go through the mailboxes
  bool inbox = false;            // the halo exchange goes through the peer inboxes instead of ncclSend/ncclRecv
  double inbox_us = 0, nccl_us = 0;   // set-up measurement behind that decision (per exchange, max over ranks)
  char *heap = nullptr;          // exported block: [mailboxes | flags]
  size_t heap_bytes = 0;
  void *peer_base[P2P_MAX_RANKS]{};
  P2PDev *dev = nullptr;
  int *err_d = nullptr;
};

// instantiated CUDA graph of one chunk of Krylov iterations (krylov.cu)
struct KrylovGraph {
  cudaGraphExec_t exec = nullptr;
  const void *rhs = nullptr, *x = nullptr;
  int mode = -1, chunk = 0;
  int64_t i0 = 0, i1 = 0;
  double kernel_launches = 0, spmv_launches = 0;
};

constexpr int32_t SELL_MB7 = 1 << 30;   // Sell::sl_m flag: mirror table padded to a multiple of 7 entries (else 8)
constexpr int32_t SELL_FAST = 1 << 29;  // Sell::sl_m flag: offset-mode slice whose columns all lie inside [0, n_cols): no clamps
constexpr int32_t SELL_MMASK = 0x0fffffff;

// sliced-ELL copy of the owned rows of the reduced system (sell.cu)
struct Sell {
  bool built = false;
  bool sym = true;            // lower columns read from the transposed position where possible (sell.cu)
  int64_t row0 = 0, n_rows = 0, n_slices = 0, n_val = 0, n_idx = 0, n_mirrored = 0;
  int nf = 1;                 // slices interleave the nf dofs per node (sell.cu)
  DevBuf<int32_t> sl_w;       // [n_slices] stored width | (offset mode ? 1<<31 : 0)
  DevBuf<int32_t> sl_m;       // [n_slices] padded number of mirrored lower columns | SELL_MB7 (table behind the offsets)
  DevBuf<int64_t> valptr;     // [n_slices+1] start of the slice's value block (doubles)
  DevBuf<int64_t> idxptr;     // [n_slices+1] start of the slice's index block (int32)
  DevBuf<double> val;         // [n_val + 64] column-major per slice, then 64 zeros
  DevBuf<int32_t> idx;        // [n_idx] offsets + mirror table (offset mode) or columns (explicit mode)
  DevBuf<int32_t> src;        // [n_val] full-CSR entry behind each position, -1 = padding (build only)
  DevBuf<int32_t> gl_ptr;     // [n_val+1] gather list of the assembly scatter per position (sell.cu)
  DevBuf<uint32_t> gl_idx;    // element-stream indices, grouped by position, ascending COO order
  DevBuf<int32_t> diag;       // [n_rows] diagonal position relative to the slice's value block, -1 = none
  void release() {
    sl_w.release(); sl_m.release(); valptr.release(); idxptr.release(); val.release(); idx.release(); src.release(); diag.release(); gl_ptr.release(); gl_idx.release();
    built = false;
  }
};

// ---- multigrid hierarchy (multigrid.cu) ---------------------------------------------------------------------------
struct CsrDev {
  int64_t n_rows = 0, nnz = 0;
  DevBuf<int32_t> ptr, idx;
  DevBuf<double> val;
  void release() { ptr.release(); idx.release(); val.release(); n_rows = nnz = 0; }
};
struct MgLevel {
  apdx_plan *coarse = nullptr;     // next coarser level (not owned)
  bool stream_borrowed = false;    // this plan is a coarse level: it runs on its finer level's stream
  CsrDev P, R;                     // P [n_free x n_free(coarse)], R = P^T
  DevBuf<int32_t> inject;          // [n_dofs(coarse)] fine full dof coinciding with each coarse full dof
  DevBuf<double> x, b, r, d, minv, ev, dofs;   // level work vectors [n_free] (ev: kept power-iteration vector); dofs [n_dofs]: injected state of a coarse level
  double lmax = 0.0;               // largest eigenvalue of D^-1 A (power iteration after every assembly)
  int pre = 2, post = 2, coarsest = 12;
  double ratio = 3.0, coarsest_ratio = 40.0;
  bool ready = false;              // values, minv and lmax belong to the current tangent
  void release() { P.release(); R.release(); inject.release(); x.release(); b.release(); r.release(); d.release(); minv.release(); ev.release(); dofs.release(); }
};

struct Stats {
  double asm_tangent_ms = 0, asm_residual_ms = 0, krylov_ms = 0, total_ms = 0;
  double krylov_iters = 0, spmv_launches = 0, kernel_launches = 0;
  double krylov_relres = 0, krylov_converged = 1;   // outcome of the last Krylov solve
};

// NVTX ranges around assemble / krylov / halo (visible in nsys / ncu --nvtx; no-ops without a profiler attached)
void nvtx_push(const char *name);
void nvtx_pop();
int sm_count();   // multiprocessors of the current device (grids are sized in multiples of it)

}  // namespace apdx

struct apdx_plan {
  int32_t dim = 0, nf = 0, n_sets = 0;
  int64_t n_nodes = 0, n_dofs = 0, n_free = 0, nnz = 0, nnz_red = 0, n_coo = 0, n_ke = 0, n_res = 0;
  std::vector<apdx::SetData> sets;
  cudaStream_t stream = nullptr;       // the stream all work of the plan is enqueued on
  cudaStream_t own_stream = nullptr;   // the stream the plan created (stream == own_stream unless apdx_plan_set_stream
                                       // handed it the caller's stream, or the plan is the coarse level of another plan)
  cudaEvent_t ev[4]{};

  // fields
  apdx::DevBuf<double> coords, dofs_n;
  double time_increment = 1.0;
  bool have_coords = false;

  // Dirichlet maps
  apdx::DevBuf<uint8_t> mask;        // [n_dofs] 1 = Dirichlet
  apdx::DevBuf<int32_t> free_id;     // [n_dofs] reduced index or -1
  apdx::DevBuf<int32_t> free_list;   // [n_free] full dof id

  // full CSR pattern + element map + gather lists of the deterministic scatter
  apdx::DevBuf<int32_t> row_ptr, col;      // [n_dofs+1], [nnz]
  apdx::DevBuf<uint32_t> perm;             // [n_coo] COO entries grouped by CSR entry
  apdx::DevBuf<int32_t> seg_ptr;           // [nnz+1] hmm: n_coo < 2^31 required
  apdx::DevBuf<uint32_t> rperm;            // [n_res] element-vector entries grouped by dof
  apdx::DevBuf<int32_t> rseg_ptr;          // [n_dofs+1]

  // reduced CSR pattern
  apdx::DevBuf<int32_t> red_row_ptr, red_col;  // [n_free+1], [nnz_red]
  apdx::DevBuf<int32_t> red2full;              // [nnz_red] -> full CSR entry
  apdx::DevBuf<int32_t> red_diag;              // [n_free] position of the diagonal in reduced data

  // values
  apdx::DevBuf<double> ke, re;             // element matrices / vectors (streams in COO order)
  apdx::DevBuf<double> vals, red_vals;     // summed CSR data
  bool have_values = false;
  bool have_ke = false;              // element matrices of the last tangent assembly are in `ke` (apdx_get_coo_values)

  // Newton / Krylov work
  apdx::DevBuf<double> residual, rhs_red, x_red, dofs_trial;
  apdx::KrylovWork kw;
  apdx::Sell sell;
  apdx::P2P p2p;
  apdx::KrylovGraph kgraph[2];   // [0] CG, [1] BiCGSTAB
  apdx::MgLevel mg;
  bool have_sell_values = false, have_red_values = false;
  bool x0_is_zero = false;                 // the caller of krylov_solve has just zeroed the whole initial guess
  double *pinned = nullptr;                // small pinned host staging
  apdx::Stats stats;
  std::vector<double> newton_history;      // residual norm after every step of the last apdx_newton

  // partition (multi-GPU)
  int64_t owned_begin = 0, owned_end = 0;  // local dof range
  int64_t f0 = 0, f1 = 0;                  // owned range in reduced numbering
  int32_t rank_lo = -1, rank_hi = -1;
  int64_t halo_lo = 0, halo_hi = 0;        // ghost free-dof counts below / above
  int64_t send_lo = 0, send_hi = 0;        // owned free dofs the lower / upper neighbour ghosts
  // general partition (apdx_plan_set_partition_lists: RCB / unstructured meshes): local order [owned | ghosts grouped by
  // owner]; the owned entries neighbour i ghosts are packed through send_idx[send_ptr[i] .. send_ptr[i+1]) into sendbuf,
  // its entries arrive in the contiguous block [recv_begin[i], recv_begin[i] + recv_count[i]) of the reduced numbering
  struct HaloLists {
    bool active = false;
    std::vector<int32_t> rank;
    std::vector<int64_t> send_ptr, recv_begin, recv_count;
    apdx::DevBuf<int32_t> send_idx;
    apdx::DevBuf<double> sendbuf;
  } hl;
};

namespace apdx {
// api.cu
void drop_all_krylov_graphs();
// pattern.cu
int build_pattern(apdx_plan *pl, const uint8_t *mask_h);
int coo_export(apdx_plan *pl, int64_t offset, int64_t count, double *dst_d);
int elem_map_export(const apdx_plan *pl, int64_t offset, int64_t count, int32_t *dst_d);
int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s);
// elements_fast.cu
bool fast_kernel_applies(int dim, int nf, const apdx_set_desc &d);
// elements.cu
int launch_element_kernels(apdx_plan *pl, const double *dofs_d, bool want_tangent);
int launch_gather_reduce(apdx_plan *pl, int tangent_flags, double *residual_d);
// krylov.cu
int krylov_alloc(apdx_plan *pl);
int spmv_reduced(apdx_plan *pl, const double *x, double *y);
int krylov_solve(apdx_plan *pl, const apdx_krylov_opts *o, const double *rhs, double *x,
                 int32_t *iters, double *relres);
int time_spmv(apdx_plan *pl, int reps, double *ms_avg);
// multigrid.cu
int mg_pcg_solve(apdx_plan *pl, const apdx_krylov_opts *o, const double *rhs, double *x, int32_t *iters, double *relres);
int mg_level_setup(apdx_plan *pl);     // minv + lambda_max of this level's current sliced-ELL matrix
int mg_inject(apdx_plan *fine, const double *fine_dofs, double *coarse_dofs);
int mg_link(apdx_plan *fine, apdx_plan *coarse, const int32_t *p_ptr, const int32_t *p_idx, const double *p_val,
            const int32_t *r_ptr, const int32_t *r_idx, const double *r_val, const int64_t *inject_h);
int mg_link_structured(apdx_plan *fine, apdx_plan *coarse, int dim, const int64_t *dims_f, const int64_t *dims_c,
                       int64_t plane_off_f, int64_t plane_off_c);
// sell.cu
int sell_build(apdx_plan *pl);
int sell_gather_reduce(apdx_plan *pl);
// dist.cu
bool comm_active();
int comm_size();
int comm_allreduce_sum(double *buf_d, int count, cudaStream_t s);
int comm_halo_exchange(apdx_plan *pl, double *x_d, cudaStream_t s);
int comm_exchange_planes(double *buf_d, int64_t n, int64_t lo, int64_t hi, int rank_lo, int rank_hi, cudaStream_t s);
int comm_halo_setup(apdx_plan *pl);
int comm_halo_setup_lists(apdx_plan *pl);
int p2p_setup(apdx_plan *pl);
void p2p_teardown(apdx_plan *pl);
}  // namespace apdx
