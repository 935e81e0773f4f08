// C ABI of libapdx_b200.so: plan life cycle, run-time fields, assembly, the Newton hot loop.
// See include/apdx_b200.h for the contract and the reference functions each entry replaces.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace apdx {

// NVTX v3 is header-only: the ranges cost a function-pointer test unless a profiler injected its library
void nvtx_push(const char *name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

size_t g_plan_bytes = 0;
// live plans: apdx_comm_destroy must drop their captured CUDA graphs (they hold NCCL operations of the communicator)
static std::vector<apdx_plan *> g_plans;
void drop_all_krylov_graphs() {
  for (apdx_plan *pl : g_plans) {
    if (pl->stream) cudaStreamSynchronize(pl->stream);
    for (auto &g : pl->kgraph)
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  }
}
static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- small kernels of the Newton loop (utility.mask_op semantics, utility.py:283-368) --------
__global__ void k_impose(double *__restrict__ dofs, const uint8_t *__restrict__ mask,
                         const double *__restrict__ vals, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && mask[i]) dofs[i] = vals[i];
}
__global__ void k_rhs_reduced(const double *__restrict__ residual, const int32_t *__restrict__ free_list,
                              int64_t n_free, double *__restrict__ rhs, double *__restrict__ x0) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_free) return;
  rhs[q] = -residual[free_list[q]];  // solver.py:611, :1514-1515
  x0[q] = 0.0;
}
__global__ void k_newton_update(double *__restrict__ dofs, const int32_t *__restrict__ free_list,
                                const double *__restrict__ x, double damping, int64_t n_free) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_free) dofs[free_list[q]] += damping * x[q];  // solver.py:881-885
}
__global__ void k_mixed_vector(const double *__restrict__ dofs, const int32_t *__restrict__ free_id,
                               const double *__restrict__ x, int64_t n, double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t q = free_id[i];
  out[i] = q >= 0 ? x[q] : dofs[i];  // solver.py:648-656
}
__global__ void k_rhs_gather(const double *__restrict__ rhs_full, const int32_t *__restrict__ free_list, int64_t n_free,
                             double *__restrict__ rhs, double *__restrict__ x0) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_free) return;
  rhs[q] = rhs_full[free_list[q]];
  x0[q] = 0.0;
}
__global__ void k_free_scatter(const int32_t *__restrict__ free_id, const double *__restrict__ x, int64_t n,
                               double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t q = free_id[i];
  out[i] = q >= 0 ? x[q] : 0.0;  // utility.mask_op(zeros, free_dofs_flat, u_f, 'set') (implicit_diff.py:229-232)
}
__global__ void __launch_bounds__(256) k_norm_partial(const double *__restrict__ residual,
                                                      const int32_t *__restrict__ free_list, int64_t q0, int64_t q1,
                                                      double *__restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += (int64_t)gridDim.x * blockDim.x) {
    double v = residual[free_list[q]];
    acc += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) k_norm_final(const double *__restrict__ partial, int n, double *out) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    *out = s;
  }
}
__global__ void k_lower_bound(const int32_t *__restrict__ sorted, int64_t n, int64_t key, int64_t *out) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < key) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}

// FP64 peak probe: 16 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456) out[0] = s;   // keeps the chains alive
}

static inline unsigned g1(int64_t n) { return (unsigned)((n + 255) / 256); }

static int assemble_internal(apdx_plan *pl, const double *dofs_d, int tangent_flags, double *residual_d) {
  struct Range { Range(const char *n) { nvtx_push(n); } ~Range() { nvtx_pop(); } } range(tangent_flags ? "apdx:assemble_tangent" : "apdx:assemble_residual");
  APDX_CHECK(launch_element_kernels(pl, dofs_d, tangent_flags != 0));
  // tangent_flags: bit 0 full CSR values, bit 1 reduced CSR values (cold paths), bit 2 sliced-ELL values (solver)
  APDX_CHECK(launch_gather_reduce(pl, tangent_flags & 3, residual_d));
  if (tangent_flags & 4) {
    APDX_CHECK(sell_gather_reduce(pl));
    pl->have_sell_values = true;
    pl->mg.ready = false;   // a multigrid hierarchy hanging off this plan belongs to the previous tangent
  }
  if (tangent_flags) {
    pl->have_ke = true;
    pl->have_values = (tangent_flags & 1) != 0;
    pl->have_red_values = (tangent_flags & 2) != 0;
  }
  return APDX_OK;
}

// squared 2-norm of residual[free] over the owned range (solver.py:899-903), all ranks
static int residual_norm(apdx_plan *pl, const double *residual_d, double *out) {
  cudaStream_t s = pl->stream;
  double *scratch = pl->kw.partial.p;
  k_norm_partial<<<NORM_GRID, 256, 0, s>>>(residual_d, pl->free_list.p, pl->f0, pl->f1, scratch);
  k_norm_final<<<1, 256, 0, s>>>(scratch, NORM_GRID, scratch + NORM_GRID);
  pl->stats.kernel_launches += 2;
  if (comm_active()) APDX_CHECK(comm_allreduce_sum(scratch + NORM_GRID, 1, s));
  APDX_CUDA(cudaMemcpyAsync(pl->pinned + 40, scratch + NORM_GRID, sizeof(double), cudaMemcpyDeviceToHost, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  *out = pl->pinned[40];
  return APDX_OK;
}

static int ensure_newton_buffers(apdx_plan *pl) {
  if (!pl->residual.p) {
    APDX_CHECK(pl->residual.alloc(pl->n_dofs));
    APDX_CHECK(pl->rhs_red.alloc(pl->n_free));
    APDX_CHECK(pl->x_red.alloc(pl->n_free));
  }
  return krylov_alloc(pl);
}

static float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// Multigrid set-up after a tangent assembly at dofs_d: smoother data of this level, then the coarser levels' operators
// re-discretised at the injected state (multigrid.cu)
static int mg_prepare(apdx_plan *pl, const double *dofs_d) {
  APDX_CHECK(mg_level_setup(pl));
  if (apdx_plan *c = pl->mg.coarse) {
    APDX_CHECK(ensure_newton_buffers(c));
    APDX_CHECK(mg_inject(pl, dofs_d, c->mg.dofs.p));
    // partitioned hierarchy: a coarse ghost plane may coincide with a fine plane this rank does not hold (two planes
    // beyond its owned range) -- the neighbour, which owns that coarse plane, sends its injected values
    if (comm_active())
      APDX_CHECK(comm_exchange_planes(c->mg.dofs.p, c->n_dofs, c->owned_begin, c->n_dofs - c->owned_end, c->rank_lo, c->rank_hi,
                                      pl->stream));
    APDX_CHECK(assemble_internal(c, c->mg.dofs.p, 4, c->residual.p));
    pl->stats.kernel_launches += c->stats.kernel_launches;
    c->stats = Stats();
    APDX_CHECK(mg_prepare(c, c->mg.dofs.p));
  }
  return APDX_OK;
}

// one linear step on dofs (already holding the imposed Dirichlet values): assemble, solve; x_red = delta
static int linear_step_internal(apdx_plan *pl, const apdx_krylov_opts *opts, const double *dofs_d, int32_t *kiters) {
  cudaStream_t s = pl->stream;
  APDX_CUDA(cudaEventRecord(pl->ev[0], s));
  APDX_CHECK(assemble_internal(pl, dofs_d, 4, pl->residual.p));
  k_rhs_reduced<<<g1(pl->n_free), 256, 0, s>>>(pl->residual.p, pl->free_list.p, pl->n_free, pl->rhs_red.p, pl->x_red.p);
  pl->x0_is_zero = true;
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaEventRecord(pl->ev[1], s));
  if (opts->jacobi == APDX_PRECOND_MULTIGRID) APDX_CHECK(mg_prepare(pl, dofs_d));   // counted with the Krylov time
  int32_t it = 0;
  double rr = 0;
  APDX_CHECK(krylov_solve(pl, opts, pl->rhs_red.p, pl->x_red.p, &it, &rr));
  APDX_CUDA(cudaEventRecord(pl->ev[2], s));
  APDX_CUDA(cudaEventSynchronize(pl->ev[2]));
  pl->stats.asm_tangent_ms += elapsed(pl->ev[0], pl->ev[1]);
  pl->stats.krylov_ms += elapsed(pl->ev[1], pl->ev[2]);
  if (kiters) *kiters = it;
  return APDX_OK;
}

}  // namespace apdx

using namespace apdx;

extern "C" {

int apdx_abi_version(void) { return APDX_ABI_VERSION; }
const char *apdx_last_error(void) { return g_err; }

int apdx_device_count(int *count) {
  APDX_CUDA(cudaGetDeviceCount(count));
  return APDX_OK;
}
int apdx_set_device(int device) {
  APDX_CUDA(cudaSetDevice(device));
  return APDX_OK;
}
int apdx_malloc(void **ptr_d, size_t bytes) {
  APDX_CUDA(cudaMalloc(ptr_d, bytes ? bytes : 8));
  return APDX_OK;
}
int apdx_free(void *ptr_d) {
  APDX_CUDA(cudaFree(ptr_d));
  return APDX_OK;
}
int apdx_host_alloc(void **ptr_h, size_t bytes) {
  APDX_CUDA(cudaMallocHost(ptr_h, bytes ? bytes : 8));
  return APDX_OK;
}
int apdx_host_free(void *ptr_h) {
  APDX_CUDA(cudaFreeHost(ptr_h));
  return APDX_OK;
}
int apdx_host_register(void *ptr_h, size_t bytes) {
  cudaError_t e = cudaHostRegister(ptr_h, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();  // page-locking is an optimisation: clear the error state (e.g. RLIMIT_MEMLOCK) and report
    set_error("cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return APDX_ERR_CUDA;
  }
  return APDX_OK;
}
int apdx_host_unregister(void *ptr_h) {
  APDX_CUDA(cudaHostUnregister(ptr_h));
  return APDX_OK;
}
int apdx_memcpy_h2d(void *dst_d, const void *src_h, size_t bytes) {
  APDX_CUDA(cudaMemcpy(dst_d, src_h, bytes, cudaMemcpyHostToDevice));
  return APDX_OK;
}
int apdx_memcpy_d2h(void *dst_h, const void *src_d, size_t bytes) {
  APDX_CUDA(cudaMemcpy(dst_h, src_d, bytes, cudaMemcpyDeviceToHost));
  return APDX_OK;
}
int apdx_memset(void *dst_d, int value, size_t bytes) {
  APDX_CUDA(cudaMemset(dst_d, value, bytes));
  return APDX_OK;
}
int apdx_synchronize(void) {
  APDX_CUDA(cudaDeviceSynchronize());
  return APDX_OK;
}
int apdx_mem_info(size_t *free_bytes, size_t *total_bytes) {
  APDX_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  return APDX_OK;
}

// ---- plan ------------------------------------------------------------------------------------------
static int validate_set(const apdx_set_desc &d, int dim, int nf, int idx) {
  APDX_REQUIRE(d.kind >= APDX_SET_DOMAIN && d.kind <= APDX_SET_INTPOINT, APDX_ERR_INVALID, "set %d: bad kind %d", idx, d.kind);
  APDX_REQUIRE(d.model >= APDX_MODEL_POISSON_POTENTIAL && d.model <= APDX_MODEL_PATTERN_ONLY, APDX_ERR_UNSUPPORTED,
               "set %d: model id %d is not a supported closed-form model", idx, d.model);
  APDX_REQUIRE(d.n_rows >= 0, APDX_ERR_INVALID, "set %d: negative row count", idx);
  APDX_REQUIRE(d.conn_itemsize == 4 || d.conn_itemsize == 8, APDX_ERR_INVALID, "set %d: connectivity itemsize must be 4 or 8", idx);
  APDX_REQUIRE(d.conn_h || d.n_rows == 0, APDX_ERR_INVALID, "set %d: connectivity pointer is NULL", idx);
  if (d.model == APDX_MODEL_PATTERN_ONLY) {   // no arithmetic: the set only contributes its all-pairs block to the pattern
    APDX_REQUIRE(d.kind == APDX_SET_DOMAIN, APDX_ERR_INVALID, "set %d: a pattern-only set is a domain set", idx);
    APDX_REQUIRE(d.nen >= 1 && d.nen <= 64, APDX_ERR_UNSUPPORTED, "set %d: %d nodes per pattern-only row not supported", idx, d.nen);
    return APDX_OK;
  }
  APDX_REQUIRE(d.nen >= 2 && d.nen <= 27, APDX_ERR_UNSUPPORTED, "set %d: %d nodes per element not supported", idx, d.nen);
  const bool scalar = d.model == APDX_MODEL_POISSON_POTENTIAL || d.model == APDX_MODEL_POISSON_WEAK || d.model == APDX_MODEL_CAPACITY;
  const bool vector = d.model == APDX_MODEL_LINEAR_ELASTICITY || d.model == APDX_MODEL_NEO_HOOKE;
  if (scalar) APDX_REQUIRE(nf == 1, APDX_ERR_UNSUPPORTED, "set %d: scalar model needs one dof per node, got %d", idx, nf);
  if (vector) APDX_REQUIRE(nf == dim, APDX_ERR_UNSUPPORTED, "set %d: elasticity model needs nf == dim (%d), got %d", idx, dim, nf);
  if (vector) {
    bool ok = (dim == 3 && d.mode == APDX_MODE_3D) || (d.mode == APDX_MODE_LAME && d.model == APDX_MODEL_LINEAR_ELASTICITY) ||
              (dim == 2 && (d.mode == APDX_MODE_PLAIN_STRAIN || (d.mode == APDX_MODE_PLAIN_STRESS && d.model == APDX_MODEL_LINEAR_ELASTICITY)));
    APDX_REQUIRE(ok, APDX_ERR_UNSUPPORTED, "set %d: elasticity mode %d not available for dim %d / this model", idx, d.mode, dim);
  }
  if (d.kind == APDX_SET_INTPOINT) {
    APDX_REQUIRE(d.n_gp == 1, APDX_ERR_INVALID, "set %d: integration-point sets have n_gp = 1", idx);
  } else {
    APDX_REQUIRE(d.n_gp >= 1 && d.n_gp <= 64, APDX_ERR_UNSUPPORTED, "set %d: %d Gauss points not supported", idx, d.n_gp);
    APDX_REQUIRE(d.shape_n_h && d.shape_dn_h && d.gp_w_h, APDX_ERR_INVALID, "set %d: shape tables missing", idx);
    int want = d.kind == APDX_SET_DOMAIN ? dim : dim - 1;
    APDX_REQUIRE(d.dim_ref == want, APDX_ERR_INVALID, "set %d: reference dimension %d, expected %d", idx, d.dim_ref, want);
    if (d.kind == APDX_SET_SURFACE)
      APDX_REQUIRE(d.model == APDX_MODEL_NEUMANN, APDX_ERR_UNSUPPORTED, "set %d: surface elements support neumann_weak only", idx);
    if (d.kind == APDX_SET_DOMAIN)
      APDX_REQUIRE(d.model != APDX_MODEL_NEUMANN, APDX_ERR_UNSUPPORTED,
                   "set %d: neumann_weak is only available in 'sparse' / surface sets", idx);
  }
  return APDX_OK;
}

int apdx_plan_create(apdx_plan **plan, int32_t dim, int64_t n_nodes, int32_t nf, int32_t n_sets,
                     const apdx_set_desc *sets, const uint8_t *dirichlet_mask_h) {
  APDX_REQUIRE(plan && sets, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(dim == 2 || dim == 3, APDX_ERR_UNSUPPORTED, "dim %d not supported", dim);
  APDX_REQUIRE(nf == 1 || nf == dim, APDX_ERR_UNSUPPORTED, "nf=%d with dim=%d not supported", nf, dim);
  APDX_REQUIRE(n_nodes > 0 && n_sets > 0, APDX_ERR_INVALID, "empty problem");
  APDX_REQUIRE(n_sets <= 16, APDX_ERR_UNSUPPORTED, "more than 16 connectivity sets per plan");
  APDX_REQUIRE(n_nodes * nf < (1ll << 31), APDX_ERR_UNSUPPORTED, "more than 2^31 dofs per plan: partition the mesh");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  APDX_REQUIRE(ce == cudaSuccess && ndev > 0, APDX_ERR_CUDA,
               "no CUDA device available (%s): the b200 backend has no CPU path", cudaGetErrorString(ce));
  for (int i = 0; i < n_sets; ++i) APDX_CHECK(validate_set(sets[i], dim, nf, i));

  apdx_plan *pl = new (std::nothrow) apdx_plan();
  APDX_REQUIRE(pl, APDX_ERR_NOMEM, "out of host memory");
  int rc = APDX_OK;
  auto fail = [&](int code) {
    apdx_plan_destroy(pl);
    return code;
  };
  pl->dim = dim; pl->nf = nf; pl->n_sets = n_sets; pl->n_nodes = n_nodes; pl->n_dofs = n_nodes * nf;
  if (cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    return fail(APDX_ERR_CUDA);
  }
  pl->stream = pl->own_stream;
  for (auto &e : pl->ev) cudaEventCreate(&e);
  if (cudaMallocHost((void **)&pl->pinned, 64 * sizeof(double)) != cudaSuccess) {
    set_error("cudaMallocHost failed");
    return fail(APDX_ERR_CUDA);
  }
  pl->sets.resize(n_sets);
  int64_t coo = 0, res = 0, kes = 0;
  for (int i = 0; i < n_sets; ++i) {
    SetData &st = pl->sets[i];
    st.d = sets[i];
    st.ndof_e = st.d.nen * nf;
    st.soa = fast_kernel_applies(dim, nf, st.d);
    st.tri = !st.soa && st.d.model != APDX_MODEL_PATTERN_ONLY;   // generic kernel (elements.cu, phase C)
    st.coo_offset = coo;
    st.res_offset = res;
    st.ke_offset = kes;
    // element-matrix stream: register-kernel sets store the upper triangle only (pattern.cu: soa_address)
    kes += st.d.n_rows * (int64_t)((st.soa || st.tri) ? st.ndof_e * (st.ndof_e + 1) / 2 : st.ndof_e * st.ndof_e);
    coo += st.d.n_rows * (int64_t)st.ndof_e * st.ndof_e;
    res += st.d.n_rows * (int64_t)st.ndof_e;
    const int64_t cn = st.d.n_rows * st.d.nen;
    if (cn > 0) {
      if ((rc = st.conn.alloc(cn)) != APDX_OK) return fail(rc);
      std::vector<int32_t> tmp;
      const int32_t *src32 = nullptr;
      if (st.d.conn_itemsize == 8) {
        tmp.resize(cn);
        const int64_t *c64 = static_cast<const int64_t *>(st.d.conn_h);
        for (int64_t k = 0; k < cn; ++k) {
          if (c64[k] < 0 || c64[k] >= n_nodes) {   // check the 64-bit value: narrowing first could alias a corrupt id into range
            set_error("set %d: node id %lld out of range [0,%lld)", i, (long long)c64[k], (long long)n_nodes);
            return fail(APDX_ERR_INVALID);
          }
          tmp[k] = (int32_t)c64[k];
        }
        src32 = tmp.data();
      } else {
        src32 = static_cast<const int32_t *>(st.d.conn_h);
      }
      for (int64_t k = 0; k < cn; ++k)
        if (src32[k] < 0 || src32[k] >= n_nodes) {
          set_error("set %d: node id %d out of range [0,%lld)", i, src32[k], (long long)n_nodes);
          return fail(APDX_ERR_INVALID);
        }
      if (cudaMemcpy(st.conn.p, src32, cn * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("connectivity upload failed");
        return fail(APDX_ERR_CUDA);
      }
    }
    if (st.d.kind != APDX_SET_INTPOINT && st.d.model != APDX_MODEL_PATTERN_ONLY) {
      size_t a = (size_t)st.d.n_gp * st.d.nen, b = a * st.d.dim_ref, c = st.d.n_gp;
      if ((rc = st.shape_n.alloc(a)) != APDX_OK || (rc = st.shape_dn.alloc(b)) != APDX_OK ||
          (rc = st.gp_w.alloc(c)) != APDX_OK)
        return fail(rc);
      st.h_shape_n.assign(st.d.shape_n_h, st.d.shape_n_h + a);
      st.h_shape_dn.assign(st.d.shape_dn_h, st.d.shape_dn_h + b);
      st.h_gp_w.assign(st.d.gp_w_h, st.d.gp_w_h + c);
      cudaMemcpy(st.shape_n.p, st.d.shape_n_h, a * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(st.shape_dn.p, st.d.shape_dn_h, b * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(st.gp_w.p, st.d.gp_w_h, c * 8, cudaMemcpyHostToDevice);
    }
    st.d.conn_h = nullptr; st.d.shape_n_h = st.d.shape_dn_h = st.d.gp_w_h = nullptr;
  }
  pl->n_coo = coo;
  pl->n_ke = kes;
  pl->n_res = res;
  if (coo == 0) {
    set_error("no element in any set");
    return fail(APDX_ERR_INVALID);
  }
  if ((rc = build_pattern(pl, dirichlet_mask_h)) != APDX_OK) return fail(rc);
  // vals / red_vals (CSR-ordered copies for apdx_assemble / apdx_get_values) are allocated on first use (elements.cu)
  if ((rc = pl->ke.alloc(pl->n_ke > 0 ? pl->n_ke : 1)) != APDX_OK || (rc = pl->re.alloc(pl->n_res)) != APDX_OK)
    return fail(rc);
  // surface sets never write their (all-zero) tangent block: zero the stream once
  cudaMemsetAsync(pl->ke.p, 0, pl->ke.bytes(), pl->stream);
  cudaMemsetAsync(pl->re.p, 0, pl->re.bytes(), pl->stream);
  cudaStreamSynchronize(pl->stream);
  pl->owned_begin = 0; pl->owned_end = pl->n_dofs; pl->f0 = 0; pl->f1 = pl->n_free;
  if (cudaGetLastError() != cudaSuccess) {
    set_error("CUDA error during plan creation");
    return fail(APDX_ERR_CUDA);
  }
  g_plans.push_back(pl);
  *plan = pl;
  return APDX_OK;
}

int apdx_plan_destroy(apdx_plan *pl) {
  if (!pl) return APDX_OK;
  g_plans.erase(std::remove(g_plans.begin(), g_plans.end(), pl), g_plans.end());
  // a borrowed stream (the finer level's, or the caller's) may be gone already: the cudaFree calls below synchronise the device
  if (pl->stream && pl->stream == pl->own_stream) cudaStreamSynchronize(pl->stream);
  pl->mg.release();
  for (auto &st : pl->sets) {
    st.conn.release(); st.shape_n.release(); st.shape_dn.release(); st.gp_w.release();
    st.ip_n.release(); st.ip_dndx.release(); st.ip_w.release();
    for (auto &p : st.params) p.release();
  }
  pl->coords.release(); pl->dofs_n.release(); pl->mask.release(); pl->free_id.release(); pl->free_list.release();
  pl->row_ptr.release(); pl->col.release(); pl->perm.release(); pl->seg_ptr.release();
  pl->rperm.release(); pl->rseg_ptr.release(); pl->red_row_ptr.release(); pl->red_col.release();
  pl->red2full.release(); pl->red_diag.release(); pl->ke.release(); pl->re.release(); pl->vals.release();
  for (auto &g : pl->kgraph) if (g.exec) cudaGraphExecDestroy(g.exec);
  p2p_teardown(pl);
  pl->hl.send_idx.release(); pl->hl.sendbuf.release();
  pl->sell.release();
  pl->red_vals.release(); pl->residual.release(); pl->rhs_red.release(); pl->x_red.release(); pl->dofs_trial.release();
  KrylovWork &k = pl->kw;
  k.r.release(); k.p.release(); k.q.release(); k.s.release(); k.t.release(); k.phat.release();
  k.shat.release(); k.r0.release(); k.minv.release(); k.partial.release(); k.scal.release(); k.ticket.release();
  k.flags.release();
  if (pl->pinned) cudaFreeHost(pl->pinned);
  for (auto &e : pl->ev) if (e) cudaEventDestroy(e);
  if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
  delete pl;
  return APDX_OK;
}

int apdx_plan_query(const apdx_plan *pl, int64_t out[8]) {
  APDX_REQUIRE(pl && out, APDX_ERR_INVALID, "NULL argument");
  out[0] = pl->n_dofs; out[1] = pl->n_free; out[2] = pl->nnz; out[3] = pl->nnz_red; out[4] = pl->n_coo;
  out[5] = pl->f0; out[6] = pl->f1; out[7] = (int64_t)g_plan_bytes;
  return APDX_OK;
}

static int copy_i32_as_i64(const int32_t *src_d, int64_t n, int64_t *dst_h) {
  std::vector<int32_t> tmp((size_t)n);
  APDX_CUDA(cudaMemcpy(tmp.data(), src_d, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) dst_h[i] = tmp[i];
  return APDX_OK;
}

int apdx_plan_get_csr(const apdx_plan *pl, int reduced, int64_t *indptr_h, int64_t *indices_h) {
  APDX_REQUIRE(pl && indptr_h && indices_h, APDX_ERR_INVALID, "NULL argument");
  if (reduced) {
    APDX_CHECK(copy_i32_as_i64(pl->red_row_ptr.p, pl->n_free + 1, indptr_h));
    if (pl->nnz_red > 0) APDX_CHECK(copy_i32_as_i64(pl->red_col.p, pl->nnz_red, indices_h));
  } else {
    APDX_CHECK(copy_i32_as_i64(pl->row_ptr.p, pl->n_dofs + 1, indptr_h));
    APDX_CHECK(copy_i32_as_i64(pl->col.p, pl->nnz, indices_h));
  }
  return APDX_OK;
}

int apdx_plan_get_elem_map(const apdx_plan *pl, int64_t offset, int64_t count, int64_t *pos_h) {
  APDX_REQUIRE(pl && pos_h, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(offset >= 0 && count >= 0 && offset + count <= pl->n_coo, APDX_ERR_INVALID, "range outside the COO stream");
  const int64_t chunk = 1ll << 24;
  apdx::DevBuf<int32_t> tmp;
  APDX_CHECK(tmp.alloc(std::min<int64_t>(chunk, std::max<int64_t>(count, 1))));
  for (int64_t o = 0; o < count; o += chunk) {
    const int64_t c = std::min<int64_t>(chunk, count - o);
    APDX_CHECK(apdx::elem_map_export(pl, offset + o, c, tmp.p));
    APDX_CUDA(cudaStreamSynchronize(pl->stream));
    APDX_CHECK(copy_i32_as_i64(tmp.p, c, pos_h + o));
  }
  return APDX_OK;
}

// ---- fields ------------------------------------------------------------------------------------------
int apdx_set_coords(apdx_plan *pl, const double *coords_h) {
  APDX_REQUIRE(pl && coords_h, APDX_ERR_INVALID, "NULL argument");
  if (!pl->coords.p) APDX_CHECK(pl->coords.alloc(pl->n_nodes * pl->dim));
  APDX_CUDA(cudaMemcpyAsync(pl->coords.p, coords_h, pl->coords.bytes(), cudaMemcpyHostToDevice, pl->stream));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  pl->have_coords = true;
  return APDX_OK;
}

int apdx_set_param(apdx_plan *pl, int32_t set, int32_t param, int32_t layout, int32_t ncomp, const double *values_h) {
  APDX_REQUIRE(pl && values_h, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(set >= 0 && set < pl->n_sets, APDX_ERR_INVALID, "set index %d out of range", set);
  APDX_REQUIRE(param >= 0 && param < APDX_PARAM_COUNT, APDX_ERR_INVALID, "parameter id %d out of range", param);
  SetData &st = pl->sets[set];
  const bool vec = (param == APDX_PARAM_BODY_LOAD || param == APDX_PARAM_TRACTION);
  APDX_REQUIRE(ncomp == (vec ? pl->nf : 1), APDX_ERR_INVALID, "parameter %d expects %d components", param, vec ? pl->nf : 1);
  size_t count;
  ParamView v{};
  v.ncomp = ncomp;
  if (layout == APDX_LAYOUT_CONST) { count = ncomp; v.s_row = 0; v.s_gp = 0; }
  else if (layout == APDX_LAYOUT_PER_GP) { count = (size_t)st.d.n_gp * ncomp; v.s_row = 0; v.s_gp = ncomp; }
  else if (layout == APDX_LAYOUT_PER_ROW_GP) { count = (size_t)st.d.n_rows * st.d.n_gp * ncomp; v.s_row = (int64_t)st.d.n_gp * ncomp; v.s_gp = ncomp; }
  else APDX_REQUIRE(false, APDX_ERR_INVALID, "bad layout %d", layout);
  if (st.params[param].n != count) APDX_CHECK(st.params[param].alloc(count));
  APDX_CUDA(cudaMemcpyAsync(st.params[param].p, values_h, count * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  v.p = st.params[param].p;
  st.pview[param] = v;
  return APDX_OK;
}

int apdx_set_intpoint_tables(apdx_plan *pl, int32_t set, const double *n_h, const double *dndx_h, const double *w_h) {
  APDX_REQUIRE(pl && n_h && dndx_h && w_h, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(set >= 0 && set < pl->n_sets, APDX_ERR_INVALID, "set index %d out of range", set);
  SetData &st = pl->sets[set];
  APDX_REQUIRE(st.d.kind == APDX_SET_INTPOINT, APDX_ERR_INVALID, "set %d is not an integration-point set", set);
  size_t a = (size_t)st.d.n_rows * st.d.nen;
  if (!st.ip_n.p) {
    APDX_CHECK(st.ip_n.alloc(a));
    APDX_CHECK(st.ip_dndx.alloc(a * pl->dim));
    APDX_CHECK(st.ip_w.alloc(st.d.n_rows));
  }
  APDX_CUDA(cudaMemcpyAsync(st.ip_n.p, n_h, a * 8, cudaMemcpyHostToDevice, pl->stream));
  APDX_CUDA(cudaMemcpyAsync(st.ip_dndx.p, dndx_h, a * pl->dim * 8, cudaMemcpyHostToDevice, pl->stream));
  APDX_CUDA(cudaMemcpyAsync(st.ip_w.p, w_h, st.d.n_rows * 8, cudaMemcpyHostToDevice, pl->stream));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  return APDX_OK;
}

int apdx_set_time_increment(apdx_plan *pl, double dt) {
  APDX_REQUIRE(pl && dt != 0.0, APDX_ERR_INVALID, "bad time increment");
  pl->time_increment = dt;
  return APDX_OK;
}

int apdx_set_dofs_n(apdx_plan *pl, const double *dofs_n_h) {
  APDX_REQUIRE(pl && dofs_n_h, APDX_ERR_INVALID, "NULL argument");
  if (!pl->dofs_n.p) APDX_CHECK(pl->dofs_n.alloc(pl->n_dofs));
  APDX_CUDA(cudaMemcpyAsync(pl->dofs_n.p, dofs_n_h, pl->dofs_n.bytes(), cudaMemcpyHostToDevice, pl->stream));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  return APDX_OK;
}

// ---- assembly / linear algebra ---------------------------------------------------------------------------
int apdx_assemble(apdx_plan *pl, const double *dofs_d, int want_tangent, double *residual_d) {
  APDX_REQUIRE(pl && dofs_d, APDX_ERR_INVALID, "NULL argument");
  pl->stats = Stats();
  APDX_CUDA(cudaEventRecord(pl->ev[0], pl->stream));
  APDX_CHECK(assemble_internal(pl, dofs_d, want_tangent ? 5 : 0, residual_d));
  APDX_CUDA(cudaEventRecord(pl->ev[1], pl->stream));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  (want_tangent ? pl->stats.asm_tangent_ms : pl->stats.asm_residual_ms) = elapsed(pl->ev[0], pl->ev[1]);
  return APDX_OK;
}

int apdx_get_values(const apdx_plan *pl, int reduced, double *values_h) {
  APDX_REQUIRE(pl && values_h, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(pl->have_values || pl->have_sell_values, APDX_ERR_STATE, "no assembled tangent: call apdx_assemble first");
  if (reduced) {
    // cold path: the solver keeps the reduced system in sliced-ELL form; the CSR-ordered copy is summed on demand
    // from the element matrices of the last assembly
    apdx_plan *mpl = const_cast<apdx_plan *>(pl);
    if (!pl->have_red_values) {
      APDX_CHECK(launch_gather_reduce(mpl, 2, nullptr));
      APDX_CUDA(cudaStreamSynchronize(pl->stream));
      mpl->have_red_values = true;
    }
    if (pl->nnz_red > 0) APDX_CUDA(cudaMemcpy(values_h, pl->red_vals.p, pl->nnz_red * 8, cudaMemcpyDeviceToHost));
  } else {
    APDX_REQUIRE(pl->have_values, APDX_ERR_STATE, "full CSR values not assembled: call apdx_assemble first");
    APDX_CUDA(cudaMemcpy(values_h, pl->vals.p, pl->nnz * 8, cudaMemcpyDeviceToHost));
  }
  return APDX_OK;
}

int apdx_get_coo_values(apdx_plan *pl, int64_t offset, int64_t count, double *values_h) {
  APDX_REQUIRE(pl && (values_h || count == 0), APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(pl->have_ke, APDX_ERR_STATE, "no element matrices: call apdx_assemble with want_tangent first");
  APDX_REQUIRE(offset >= 0 && count >= 0 && offset + count <= pl->n_coo, APDX_ERR_INVALID,
               "COO range [%lld, %lld) outside [0, %lld)", (long long)offset, (long long)(offset + count), (long long)pl->n_coo);
  const int64_t chunk = 1ll << 24;   // 128 MB of staging
  apdx::DevBuf<double> tmp;
  APDX_CHECK(tmp.alloc(std::min<int64_t>(chunk, std::max<int64_t>(count, 1))));
  for (int64_t o = 0; o < count; o += chunk) {
    const int64_t c = std::min<int64_t>(chunk, count - o);
    APDX_CHECK(apdx::coo_export(pl, offset + o, c, tmp.p));
    APDX_CUDA(cudaMemcpyAsync(values_h + o, tmp.p, c * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
    APDX_CUDA(cudaStreamSynchronize(pl->stream));
  }
  return APDX_OK;
}

int apdx_spmv(apdx_plan *pl, const double *x_d, double *y_d) {
  APDX_REQUIRE(pl && x_d && y_d, APDX_ERR_INVALID, "NULL argument");
  APDX_CHECK(spmv_reduced(pl, x_d, y_d));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  return APDX_OK;
}

// ---- caller-owned streams ("every compute call takes a cudaStream_t and is asynchronous on it") ----------
int apdx_plan_set_stream(apdx_plan *pl, void *cuda_stream) {
  APDX_REQUIRE(pl, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(!pl->mg.stream_borrowed, APDX_ERR_STATE, "this plan is a coarse multigrid level: set the stream of the finest plan");
  APDX_CUDA(cudaStreamSynchronize(pl->stream));   // work already enqueued finishes on the stream it was enqueued on
  pl->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : pl->own_stream;
  for (apdx_plan *l = pl->mg.coarse; l; l = l->mg.coarse) l->stream = pl->stream;
  return APDX_OK;
}
int apdx_assemble_async(apdx_plan *pl, const double *dofs_d, int want_tangent, double *residual_d) {
  APDX_REQUIRE(pl && dofs_d, APDX_ERR_INVALID, "NULL argument");
  pl->stats = Stats();
  return assemble_internal(pl, dofs_d, want_tangent ? 5 : 0, residual_d);   // no host synchronisation, no timing
}
int apdx_spmv_async(apdx_plan *pl, const double *x_d, double *y_d) {
  APDX_REQUIRE(pl && x_d && y_d, APDX_ERR_INVALID, "NULL argument");
  return spmv_reduced(pl, x_d, y_d);
}
int apdx_stream_create(void **cuda_stream) {
  APDX_REQUIRE(cuda_stream, APDX_ERR_INVALID, "NULL argument");
  cudaStream_t s = nullptr;
  APDX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *cuda_stream = s;
  return APDX_OK;
}
int apdx_stream_synchronize(void *cuda_stream) {
  APDX_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
  return APDX_OK;
}
int apdx_stream_destroy(void *cuda_stream) {
  if (cuda_stream) APDX_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(cuda_stream)));
  return APDX_OK;
}

int apdx_time_spmv(apdx_plan *pl, int32_t reps, double *ms_avg) {
  APDX_REQUIRE(pl && ms_avg && reps > 0, APDX_ERR_INVALID, "bad argument");
  return time_spmv(pl, reps, ms_avg);
}

int apdx_measure_fp64_peak(double *tflops) {
  APDX_REQUIRE(tflops, APDX_ERR_INVALID, "NULL argument");
  double *d = nullptr;
  APDX_CUDA(cudaMalloc((void **)&d, sizeof(double)));
  cudaEvent_t e0, e1;
  APDX_CUDA(cudaEventCreate(&e0));
  APDX_CUDA(cudaEventCreate(&e1));
  const int blocks = sm_count() * 8, iters = 20000;
  k_fp64_peak<<<blocks, 256>>>(d, 1000);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    APDX_CUDA(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, 256>>>(d, iters);
    APDX_CUDA(cudaEventRecord(e1));
    APDX_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 16.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return APDX_OK;
}

int apdx_krylov(apdx_plan *pl, const apdx_krylov_opts *opts, const double *rhs_d, double *x_d, int32_t *iters,
                double *relres) {
  APDX_REQUIRE(pl && opts && rhs_d && x_d, APDX_ERR_INVALID, "NULL argument");
  cudaStream_t s = pl->stream;
  APDX_CHECK(krylov_alloc(pl));
  pl->x0_is_zero = false;   // x_d is the caller's initial guess
  APDX_CUDA(cudaEventRecord(pl->ev[1], s));
  APDX_CHECK(krylov_solve(pl, opts, rhs_d, x_d, iters, relres));
  APDX_CUDA(cudaEventRecord(pl->ev[2], s));
  APDX_CUDA(cudaEventSynchronize(pl->ev[2]));
  pl->stats.krylov_ms += elapsed(pl->ev[1], pl->ev[2]);
  return APDX_OK;
}

// ---- Newton ---------------------------------------------------------------------------------------------------
int apdx_linear_step(apdx_plan *pl, const apdx_krylov_opts *opts, const double *dofs_d, const double *dirichlet_values_d,
                     double *delta_d, int32_t *krylov_iters) {
  APDX_REQUIRE(pl && opts && dofs_d && delta_d, APDX_ERR_INVALID, "NULL argument");
  APDX_CHECK(ensure_newton_buffers(pl));
  if (!pl->dofs_trial.p) APDX_CHECK(pl->dofs_trial.alloc(pl->n_dofs));
  cudaStream_t s = pl->stream;
  pl->stats = Stats();
  APDX_CUDA(cudaMemcpyAsync(pl->dofs_trial.p, dofs_d, pl->n_dofs * 8, cudaMemcpyDeviceToDevice, s));
  if (dirichlet_values_d) {
    k_impose<<<g1(pl->n_dofs), 256, 0, s>>>(pl->dofs_trial.p, pl->mask.p, dirichlet_values_d, pl->n_dofs);  // solver.py:586-604
    pl->stats.kernel_launches += 1;
  }
  APDX_CHECK(linear_step_internal(pl, opts, pl->dofs_trial.p, krylov_iters));
  if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, pl->x_red.p, s));
  k_mixed_vector<<<g1(pl->n_dofs), 256, 0, s>>>(pl->dofs_trial.p, pl->free_id.p, pl->x_red.p, pl->n_dofs, delta_d);
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaStreamSynchronize(s));
  pl->stats.total_ms = pl->stats.asm_tangent_ms + pl->stats.krylov_ms;
  return APDX_OK;
}

int apdx_tangent_solve(apdx_plan *pl, const apdx_krylov_opts *opts, const double *dofs_d, const double *rhs_d,
                       int transpose, double *out_d, int32_t *krylov_iters) {
  APDX_REQUIRE(pl && opts && rhs_d && out_d, APDX_ERR_INVALID, "NULL argument");
  (void)transpose;   // A = A^T for every in-scope model (see the header)
  APDX_CHECK(ensure_newton_buffers(pl));
  cudaStream_t s = pl->stream;
  pl->stats = Stats();
  APDX_CUDA(cudaEventRecord(pl->ev[0], s));
  if (dofs_d) APDX_CHECK(assemble_internal(pl, dofs_d, 4, pl->residual.p));
  APDX_REQUIRE(pl->have_sell_values, APDX_ERR_STATE, "no assembled tangent: pass dofs_d or run a Newton step first");
  k_rhs_gather<<<g1(pl->n_free), 256, 0, s>>>(rhs_d, pl->free_list.p, pl->n_free, pl->rhs_red.p, pl->x_red.p);
  pl->x0_is_zero = true;
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaEventRecord(pl->ev[1], s));
  if (opts->jacobi == APDX_PRECOND_MULTIGRID && !pl->mg.ready) {
    APDX_REQUIRE(dofs_d, APDX_ERR_STATE, "multigrid: pass dofs_d (the hierarchy is re-discretised at the state)");
    APDX_CHECK(mg_prepare(pl, dofs_d));
  }
  int32_t it = 0;
  double rr = 0;
  APDX_CHECK(krylov_solve(pl, opts, pl->rhs_red.p, pl->x_red.p, &it, &rr));
  if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, pl->x_red.p, s));
  k_free_scatter<<<g1(pl->n_dofs), 256, 0, s>>>(pl->free_id.p, pl->x_red.p, pl->n_dofs, out_d);
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaEventRecord(pl->ev[2], s));
  APDX_CUDA(cudaEventSynchronize(pl->ev[2]));
  pl->stats.asm_tangent_ms = elapsed(pl->ev[0], pl->ev[1]);
  pl->stats.krylov_ms = elapsed(pl->ev[1], pl->ev[2]);
  pl->stats.total_ms = pl->stats.asm_tangent_ms + pl->stats.krylov_ms;
  if (krylov_iters) *krylov_iters = it;
  return APDX_OK;
}

int apdx_newton(apdx_plan *pl, const apdx_krylov_opts *opts, double *dofs_d, const double *dirichlet_values_d,
                double newton_tol, int32_t maxiter, double damping, int32_t *iters, double *res_norm, int32_t *diverged) {
  APDX_REQUIRE(pl && opts && dofs_d && iters && res_norm && diverged, APDX_ERR_INVALID, "NULL argument");
  APDX_CHECK(ensure_newton_buffers(pl));
  cudaStream_t s = pl->stream;
  pl->stats = Stats();
  cudaEvent_t t0 = pl->ev[3];
  APDX_CUDA(cudaEventRecord(t0, s));
  // carry of solver.damped_newton (solver.py:944-946): (dofs, itt=0, not_stop=True, res_norm=0.0, diverged=False)
  int32_t itt = 0;
  bool not_stop = true, div = false;
  double rn = 0.0;
  pl->newton_history.clear();
  while (not_stop) {
    const double rn_old = rn;
    // --- lin_solve_fun: impose Dirichlet values, assemble, solve (solver.py:586-656) ---
    if (dirichlet_values_d) {
      k_impose<<<g1(pl->n_dofs), 256, 0, s>>>(dofs_d, pl->mask.p, dirichlet_values_d, pl->n_dofs);
      pl->stats.kernel_launches += 1;
    }
    APDX_CHECK(linear_step_internal(pl, opts, dofs_d, nullptr));
    // --- damped update of the free dofs; Dirichlet entries keep the imposed values (:879-892) ---
    if (comm_active()) APDX_CHECK(comm_halo_exchange(pl, pl->x_red.p, s));  // ghost dofs follow their owners
    k_newton_update<<<g1(pl->n_free), 256, 0, s>>>(dofs_d, pl->free_list.p, pl->x_red.p, damping, pl->n_free);
    pl->stats.kernel_launches += 1;
    // --- residual at the updated state and its norm over the free dofs (:898-903) ---
    APDX_CUDA(cudaEventRecord(pl->ev[0], s));
    APDX_CHECK(assemble_internal(pl, dofs_d, 0, pl->residual.p));
    APDX_CUDA(cudaEventRecord(pl->ev[1], s));
    double rn2 = 0.0;
    APDX_CHECK(residual_norm(pl, pl->residual.p, &rn2));
    pl->stats.asm_residual_ms += elapsed(pl->ev[0], pl->ev[1]);
    rn = sqrt(rn2);
    pl->newton_history.push_back(rn);   // "Residual after Newton iteration {itt+1}" (solver.py:906-909)
    not_stop = rn > newton_tol;  // :904 (NaN compares false, handled by the divergence flag below)
    bool next_step;
    if (itt < maxiter) {  // :930
      bool d = (rn / rn_old > 10.0) && (itt > 1);  // :915-917
      if (rn != rn || rn2 != rn2 || isinf(rn2)) d = true;  // :919-923
      next_step = !d;
      div = d;
    } else {  // :926-928
      next_step = false;
      div = true;
    }
    itt += 1;
    not_stop = not_stop && next_step;
  }
  APDX_CUDA(cudaEventRecord(pl->ev[2], s));
  APDX_CUDA(cudaEventSynchronize(pl->ev[2]));
  pl->stats.total_ms = elapsed(t0, pl->ev[2]);
  *iters = itt;
  *res_norm = rn;
  *diverged = div ? 1 : 0;
  return APDX_OK;
}

int apdx_plan_set_coarse(apdx_plan *fine, apdx_plan *coarse, const int32_t *p_indptr_h, const int32_t *p_indices_h,
                         const double *p_data_h, const int32_t *r_indptr_h, const int32_t *r_indices_h,
                         const double *r_data_h, const int64_t *inject_h) {
  APDX_REQUIRE(fine && coarse && fine != coarse && p_indptr_h && r_indptr_h && inject_h, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(p_indptr_h[fine->n_free] == r_indptr_h[coarse->n_free], APDX_ERR_INVALID,
               "P and R = P^T disagree on the number of entries (%d, %d)", p_indptr_h[fine->n_free], r_indptr_h[coarse->n_free]);
  return mg_link(fine, coarse, p_indptr_h, p_indices_h, p_data_h, r_indptr_h, r_indices_h, r_data_h, inject_h);
}

int apdx_plan_set_coarse_structured(apdx_plan *fine, apdx_plan *coarse, int32_t dim, const int64_t *dims_f,
                                    const int64_t *dims_c, int64_t plane_off_f, int64_t plane_off_c) {
  APDX_REQUIRE(fine && coarse && fine != coarse && dims_f && dims_c, APDX_ERR_INVALID, "NULL argument");
  return mg_link_structured(fine, coarse, dim, dims_f, dims_c, plane_off_f, plane_off_c);
}

int apdx_plan_get_transfer(const apdx_plan *fine, int32_t which, int64_t *n_rows, int64_t *nnz, int32_t *indptr_h,
                           int32_t *indices_h, double *data_h, int32_t *inject_h) {
  APDX_REQUIRE(fine && fine->mg.coarse, APDX_ERR_STATE, "no coarse level linked to this plan");
  const apdx::CsrDev &M = which ? fine->mg.R : fine->mg.P;
  if (n_rows) *n_rows = M.n_rows;
  if (nnz) *nnz = M.nnz;
  APDX_CUDA(cudaStreamSynchronize(fine->stream));
  if (indptr_h) APDX_CUDA(cudaMemcpy(indptr_h, M.ptr.p, (size_t)(M.n_rows + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (indices_h && M.nnz > 0) APDX_CUDA(cudaMemcpy(indices_h, M.idx.p, (size_t)M.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (data_h && M.nnz > 0) APDX_CUDA(cudaMemcpy(data_h, M.val.p, (size_t)M.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  if (inject_h)
    APDX_CUDA(cudaMemcpy(inject_h, fine->mg.inject.p, (size_t)fine->mg.coarse->n_dofs * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return APDX_OK;
}

int apdx_plan_set_multigrid(apdx_plan *pl, int32_t pre_degree, int32_t post_degree, int32_t coarsest_degree,
                            double smoother_ratio, double coarsest_ratio) {
  APDX_REQUIRE(pl, APDX_ERR_INVALID, "NULL argument");
  for (apdx_plan *l = pl; l; l = l->mg.coarse) {
    if (pre_degree > 0) l->mg.pre = pre_degree;
    if (post_degree > 0) l->mg.post = post_degree;
    if (coarsest_degree > 0) l->mg.coarsest = coarsest_degree;
    if (smoother_ratio > 1.0) l->mg.ratio = smoother_ratio;
    if (coarsest_ratio > 1.0) l->mg.coarsest_ratio = coarsest_ratio;
  }
  return APDX_OK;
}

int apdx_plan_sell_info(const apdx_plan *pl, int64_t out[6]) {
  APDX_REQUIRE(pl && out, APDX_ERR_INVALID, "null argument");
  APDX_REQUIRE(pl->sell.built, APDX_ERR_STATE, "no sliced-ELL matrix yet: assemble a tangent for the solver first");
  const apdx::Sell &S = pl->sell;
  out[0] = S.n_slices; out[1] = S.n_val; out[2] = S.n_idx; out[3] = S.n_mirrored; out[4] = S.sym ? 1 : 0; out[5] = S.nf;
  return APDX_OK;
}

int apdx_plan_stats(const apdx_plan *pl, double out[8]) {
  APDX_REQUIRE(pl && out, APDX_ERR_INVALID, "NULL argument");
  out[0] = pl->stats.asm_tangent_ms; out[1] = pl->stats.asm_residual_ms; out[2] = pl->stats.krylov_ms;
  out[3] = pl->stats.krylov_iters; out[4] = pl->stats.spmv_launches; out[5] = pl->stats.total_ms;
  out[6] = pl->stats.kernel_launches;
  out[7] = (double)(pl->sell.n_val * 8 + pl->sell.n_idx * 4 + pl->sell.n_slices * 24);  // bytes of the sliced-ELL matrix
  return APDX_OK;
}

int apdx_plan_comm_info(const apdx_plan *pl, double out[4]) {
  APDX_REQUIRE(pl && out, APDX_ERR_INVALID, "NULL argument");
  out[0] = pl->p2p.mbox ? 1.0 : 0.0; out[1] = pl->p2p.inbox ? 1.0 : 0.0; out[2] = pl->p2p.inbox_us; out[3] = pl->p2p.nccl_us;
  return APDX_OK;
}

int apdx_plan_newton_history(const apdx_plan *pl, double *res_norms, int32_t capacity, int32_t *count) {
  APDX_REQUIRE(pl && count && (res_norms || capacity == 0), APDX_ERR_INVALID, "NULL argument");
  *count = (int32_t)pl->newton_history.size();
  for (int32_t i = 0; i < *count && i < capacity; ++i) res_norms[i] = pl->newton_history[(size_t)i];
  return APDX_OK;
}

int apdx_plan_last_krylov(const apdx_plan *pl, double *relres, int32_t *converged) {
  APDX_REQUIRE(pl && relres && converged, APDX_ERR_INVALID, "NULL argument");
  *relres = pl->stats.krylov_relres;
  *converged = pl->stats.krylov_converged != 0.0 ? 1 : 0;
  return APDX_OK;
}

// A partition whose set-up fails on ONE rank must fail on ALL of them: the ranks that passed would otherwise wait for
// the failed one inside the collectives of the halo set-up (measured the hard way: a 13-plane brick on 8 ranks left
// rank 0 with Dirichlet dofs only, and seven GPUs idling until the job's timeout).  Sum of the local failure flags.
static int partition_agree(apdx_plan *pl, bool ok_local, const char *what) {
  double bad = ok_local ? 0.0 : 1.0;
  if (comm_active()) {
    double *d = nullptr;
    APDX_CUDA(cudaMalloc((void **)&d, sizeof(double)));
    APDX_CUDA(cudaMemcpyAsync(d, &bad, sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    int rc = comm_allreduce_sum(d, 1, pl->stream);
    if (rc == APDX_OK && cudaMemcpyAsync(&bad, d, sizeof(double), cudaMemcpyDeviceToHost, pl->stream) != cudaSuccess) rc = APDX_ERR_CUDA;
    if (rc == APDX_OK && cudaStreamSynchronize(pl->stream) != cudaSuccess) rc = APDX_ERR_CUDA;
    cudaFree(d);
    APDX_CHECK(rc);
  }
  if (!ok_local) APDX_REQUIRE(false, APDX_ERR_INVALID, "%s", what);
  APDX_REQUIRE(bad == 0.0, APDX_ERR_INVALID, "partition set-up failed on %d other rank(s): %s there", (int)bad, what);
  return APDX_OK;
}

int apdx_plan_set_partition(apdx_plan *pl, int64_t owned_dof_begin, int64_t owned_dof_end, int32_t rank_lo,
                            int32_t rank_hi) {
  APDX_REQUIRE(pl, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(0 <= owned_dof_begin && owned_dof_begin < owned_dof_end && owned_dof_end <= pl->n_dofs, APDX_ERR_INVALID,
               "bad owned range");
  APDX_REQUIRE(!pl->kw.r.p, APDX_ERR_STATE, "apdx_plan_set_partition must precede the first solve");
  int64_t *out = reinterpret_cast<int64_t *>(pl->pinned + 48);
  int64_t *tmp_d = nullptr;
  APDX_CUDA(cudaMalloc((void **)&tmp_d, 2 * sizeof(int64_t)));
  k_lower_bound<<<1, 1, 0, pl->stream>>>(pl->free_list.p, pl->n_free, owned_dof_begin, tmp_d);
  k_lower_bound<<<1, 1, 0, pl->stream>>>(pl->free_list.p, pl->n_free, owned_dof_end, tmp_d + 1);
  APDX_CUDA(cudaMemcpyAsync(out, tmp_d, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, pl->stream));
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  cudaFree(tmp_d);
  for (auto &g : pl->kgraph) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  pl->sell.release();
  pl->have_sell_values = false;
  pl->owned_begin = owned_dof_begin; pl->owned_end = owned_dof_end;
  pl->f0 = out[0]; pl->f1 = out[1];
  pl->rank_lo = rank_lo; pl->rank_hi = rank_hi;
  pl->halo_lo = pl->f0; pl->halo_hi = pl->n_free - pl->f1;
  APDX_CHECK(partition_agree(pl, pl->f1 > pl->f0 && (rank_lo >= 0 || pl->halo_lo == 0) && (rank_hi >= 0 || pl->halo_hi == 0),
                             "a rank owns no free dof (every dof of its part is a Dirichlet dof), or has ghost dofs on a "
                             "side without a neighbour"));
  if (comm_active()) {
    APDX_CHECK(comm_halo_setup(pl));
    APDX_CHECK(p2p_setup(pl));
  }
  return APDX_OK;
}

int apdx_plan_set_partition_lists(apdx_plan *pl, int64_t owned_dof_end, int32_t n_neighbours,
                                  const int32_t *neighbour_rank, const int64_t *send_ptr_h, const int64_t *send_dof_h,
                                  const int64_t *recv_dof_begin_h, const int64_t *recv_dof_end_h) {
  APDX_REQUIRE(pl, APDX_ERR_INVALID, "NULL argument");
  APDX_REQUIRE(0 < owned_dof_end && owned_dof_end <= pl->n_dofs, APDX_ERR_INVALID, "bad owned range");
  APDX_REQUIRE(n_neighbours >= 0, APDX_ERR_INVALID, "negative neighbour count");
  APDX_REQUIRE(n_neighbours == 0 || (neighbour_rank && send_ptr_h && recv_dof_begin_h && recv_dof_end_h),
               APDX_ERR_INVALID, "NULL neighbour list");
  APDX_REQUIRE(!pl->kw.r.p, APDX_ERR_STATE, "apdx_plan_set_partition_lists must precede the first solve");
  const int nn = n_neighbours;
  APDX_REQUIRE(nn == 0 || send_ptr_h[0] == 0, APDX_ERR_INVALID, "send_ptr must start at 0");
  for (int i = 0; i < nn; ++i) {
    APDX_REQUIRE(send_ptr_h[i + 1] >= send_ptr_h[i], APDX_ERR_INVALID, "send_ptr must be non-decreasing");
    APDX_REQUIRE(neighbour_rank[i] >= 0 && (!comm_active() || neighbour_rank[i] < comm_size()), APDX_ERR_INVALID,
                 "neighbour rank %d out of range", neighbour_rank[i]);
    APDX_REQUIRE(owned_dof_end <= recv_dof_begin_h[i] && recv_dof_begin_h[i] <= recv_dof_end_h[i] &&
                     recv_dof_end_h[i] <= pl->n_dofs, APDX_ERR_INVALID, "ghost block of neighbour %d outside the ghost range", i);
  }
  APDX_REQUIRE(nn == 0 || send_ptr_h[nn] == 0 || send_dof_h, APDX_ERR_INVALID, "NULL send list");
  // reduced numbering: free_id[dof] = reduced index or -1 (Dirichlet)
  std::vector<int32_t> fid((size_t)pl->n_dofs);
  APDX_CUDA(cudaStreamSynchronize(pl->stream));
  APDX_CUDA(cudaMemcpy(fid.data(), pl->free_id.p, fid.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  auto lower = [&](int64_t dof) -> int64_t {   // number of free dofs below `dof`
    for (int64_t j = dof; j < pl->n_dofs; ++j)
      if (fid[(size_t)j] >= 0) return fid[(size_t)j];
    return pl->n_free;
  };
  auto &H = pl->hl;
  H.rank.assign(neighbour_rank, neighbour_rank + nn);
  H.send_ptr.assign((size_t)nn + 1, 0);
  H.recv_begin.assign((size_t)nn, 0);
  H.recv_count.assign((size_t)nn, 0);
  std::vector<int32_t> sidx;
  for (int i = 0; i < nn; ++i) {
    for (int64_t k = send_ptr_h[i]; k < send_ptr_h[i + 1]; ++k) {
      const int64_t d = send_dof_h[k];
      APDX_REQUIRE(0 <= d && d < owned_dof_end, APDX_ERR_INVALID, "send list of neighbour %d holds dof %lld, which this rank does not own",
                   i, (long long)d);
      if (fid[(size_t)d] >= 0) sidx.push_back(fid[(size_t)d]);
    }
    H.send_ptr[(size_t)i + 1] = (int64_t)sidx.size();
    H.recv_begin[(size_t)i] = lower(recv_dof_begin_h[i]);
    H.recv_count[(size_t)i] = lower(recv_dof_end_h[i]) - H.recv_begin[(size_t)i];
  }
  APDX_CHECK(H.send_idx.alloc(sidx.size() > 0 ? sidx.size() : 1));
  APDX_CHECK(H.sendbuf.alloc(sidx.size() > 0 ? sidx.size() : 1));
  if (!sidx.empty())
    APDX_CUDA(cudaMemcpy(H.send_idx.p, sidx.data(), sidx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  H.active = true;
  for (auto &g : pl->kgraph) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  pl->sell.release();
  pl->have_sell_values = false;
  pl->owned_begin = 0; pl->owned_end = owned_dof_end;
  pl->f0 = 0; pl->f1 = lower(owned_dof_end);
  pl->rank_lo = pl->rank_hi = -1;
  pl->halo_lo = pl->halo_hi = 0;
  pl->send_lo = pl->send_hi = 0;
  APDX_CHECK(partition_agree(pl, pl->f1 > pl->f0, "a rank owns no free dof (every dof of its part is a Dirichlet dof)"));
  if (comm_active()) {
    APDX_CHECK(comm_halo_setup_lists(pl));
    APDX_CHECK(p2p_setup(pl));
  }
  return APDX_OK;
}

}  // extern "C"
