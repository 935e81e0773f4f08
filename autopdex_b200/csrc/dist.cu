// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.  The reference has no
// distributed code at all; this implements SURVEY.md 8e: slab partitions whose
// ghost planes are contiguous ranges of the reduced numbering, exchanged with grouped
// ncclSend/ncclRecv before every SpMV, and ncclAllReduce for dot products / the Newton norm.
// NCCL is loaded with dlopen so that the single-GPU library has no link-time dependency on it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace apdx {

// minimal NCCL declarations (ABI-stable subset of nccl.h, NCCL >= 2.7)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};
static Nccl g_nccl;

#define APDX_NCCL(call)                                                                       \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != 0) {                                                                           \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                            \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");           \
      return APDX_ERR_NCCL;                                                                   \
    }                                                                                         \
  } while (0)

static int load_nccl() {
  if (g_nccl.lib) return APDX_OK;
  // APDX_NCCL_LIB: path of another NCCL build (like APDX_LIB for this library)
  const char *names[] = {getenv("APDX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  APDX_REQUIRE(g_nccl.lib, APDX_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                        \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                          \
  APDX_REQUIRE(g_nccl.field, APDX_ERR_NCCL, "symbol %s missing in libnccl", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(AllGather, "ncclAllGather");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return APDX_OK;
}

bool comm_active() { return g_nccl.comm != nullptr && g_nccl.nranks > 1; }
int comm_size() { return g_nccl.nranks; }

int comm_allreduce_sum(double *buf_d, int count, cudaStream_t s) {
  APDX_NCCL(g_nccl.AllReduce(buf_d, buf_d, (size_t)count, ncclFloat64, ncclSum, g_nccl.comm, s));
  return APDX_OK;
}

// general partition: sendbuf[k] = x[send_idx[k]] (owned entries the neighbours ghost, grouped by neighbour)
__global__ void k_halo_pack(const double *__restrict__ x, const int32_t *__restrict__ send_idx, int64_t n,
                            double *__restrict__ sendbuf) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    sendbuf[k] = x[send_idx[k]];
}

static int comm_halo_exchange_lists(apdx_plan *pl, double *x_d, cudaStream_t s) {
  auto &H = pl->hl;
  const int nn = (int)H.rank.size();
  if (nn == 0) return APDX_OK;
  const int64_t n_send = H.send_ptr[nn];
  if (n_send > 0) {
    const int64_t cap = (int64_t)sm_count() * 8;
    const unsigned grid = (unsigned)((n_send + 255) / 256 < cap ? (n_send + 255) / 256 : cap);
    k_halo_pack<<<grid, 256, 0, s>>>(x_d, H.send_idx.p, n_send, H.sendbuf.p);
    pl->stats.kernel_launches += 1;
    APDX_CUDA(cudaGetLastError());
  }
  APDX_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < nn; ++i) {
    const int64_t cs = H.send_ptr[i + 1] - H.send_ptr[i];
    if (cs > 0) APDX_NCCL(g_nccl.Send(H.sendbuf.p + H.send_ptr[i], (size_t)cs, ncclFloat64, H.rank[i], g_nccl.comm, s));
    if (H.recv_count[i] > 0)
      APDX_NCCL(g_nccl.Recv(x_d + H.recv_begin[i], (size_t)H.recv_count[i], ncclFloat64, H.rank[i], g_nccl.comm, s));
  }
  APDX_NCCL(g_nccl.GroupEnd());
  return APDX_OK;
}

// every pair of neighbours must agree on the message lengths (a mismatch would be undefined behaviour inside NCCL):
// each rank tells neighbour i how many entries it expects from it and compares the answer with what it will send
int comm_halo_setup_lists(apdx_plan *pl) {
  auto &H = pl->hl;
  const int nn = (int)H.rank.size();
  if (nn == 0) return APDX_OK;
  int64_t *buf = nullptr;
  APDX_CUDA(cudaMalloc((void **)&buf, 2 * (size_t)nn * sizeof(int64_t)));
  APDX_CUDA(cudaMemcpy(buf, H.recv_count.data(), (size_t)nn * sizeof(int64_t), cudaMemcpyHostToDevice));
  cudaStream_t s = pl->stream;
  APDX_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < nn; ++i) {
    APDX_NCCL(g_nccl.Send(buf + i, 1, ncclInt64, H.rank[i], g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf + nn + i, 1, ncclInt64, H.rank[i], g_nccl.comm, s));
  }
  APDX_NCCL(g_nccl.GroupEnd());
  APDX_CUDA(cudaStreamSynchronize(s));
  std::vector<int64_t> expect(nn);
  APDX_CUDA(cudaMemcpy(expect.data(), buf + nn, (size_t)nn * sizeof(int64_t), cudaMemcpyDeviceToHost));
  cudaFree(buf);
  for (int i = 0; i < nn; ++i)
    APDX_REQUIRE(expect[i] == H.send_ptr[i + 1] - H.send_ptr[i], APDX_ERR_INVALID,
                 "halo lists disagree: rank %d expects %lld entries from this rank, the send list holds %lld", H.rank[i],
                 (long long)expect[i], (long long)(H.send_ptr[i + 1] - H.send_ptr[i]));
  return APDX_OK;
}

// ---- halo exchange through peer inboxes (slab partitions) ---------------------------------------------------------
// ncclSend/ncclRecv of one node plane (528 KB at 256^3) costs ~39 us per CG iteration on 8 GPUs -- latency, not bandwidth
// (profiles/r02b_trace_n8.txt).  Here every rank stores its boundary entries straight into its neighbours' inboxes over
// NVLink and raises a flag there (k_halo_push), then waits for its own flags and copies the inboxes into the ghost
// entries of x (k_halo_pull).  Two small kernels, no NCCL, replayable inside the Krylov graph: the exchange counter
// lives on the device and its parity selects one of two inbox slots (a rank can be at most one exchange ahead of a
// neighbour: its next push needs the neighbour's previous one).  Two kernels rather than one, so that no block ever
// waits for a peer while blocks of the same grid still have to post (no co-residency assumption).
constexpr int HALO_BLOCKS = 16;
__device__ __forceinline__ void halo_wait(const int *flag, int epoch, int *err) {
  const long long t0 = clock64();
  while (*(volatile const int *)flag < epoch) {
    if (clock64() - t0 > 20000000000ll) {  // ~10 s (ranks may enter a solve seconds apart): a lost peer must not hang the GPU
      *err = 1;
      break;
    }
  }
  __threadfence_system();
}
__global__ void __launch_bounds__(256) k_halo_push(const double *__restrict__ x, int64_t f0, int64_t f1, int64_t send_lo,
                                                   int64_t send_hi, const P2PDev *pd) {
  __shared__ bool last;
  const int e = *pd->halo_epoch + 1;
  const size_t slot = (size_t)(e & 1) * (size_t)pd->cap;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (pd->peer_inbox_lo)
    for (int64_t i = tid; i < send_lo; i += nth) pd->peer_inbox_lo[slot + i] = x[f0 + i];
  if (pd->peer_inbox_hi)
    for (int64_t i = tid; i < send_hi; i += nth) pd->peer_inbox_hi[slot + i] = x[f1 - send_hi + i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(pd->halo_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {   // every block's stores are visible system-wide: raise the flags
    pd->halo_ticket[0] = 0u;
    __threadfence_system();
    if (pd->peer_flag_lo) *(volatile int *)(pd->peer_flag_lo + (e & 1)) = e;
    if (pd->peer_flag_hi) *(volatile int *)(pd->peer_flag_hi + (e & 1)) = e;
  }
}
__global__ void __launch_bounds__(256) k_halo_pull(double *__restrict__ x, int64_t f1, int64_t halo_lo, int64_t halo_hi,
                                                   const P2PDev *pd) {
  __shared__ bool last;
  const int e = *pd->halo_epoch + 1;
  const size_t slot = (size_t)(e & 1) * (size_t)pd->cap;
  if (threadIdx.x == 0) {
    if (pd->my_inbox_lo && halo_lo > 0) halo_wait(pd->my_flag_lo + (e & 1), e, pd->err);
    if (pd->my_inbox_hi && halo_hi > 0) halo_wait(pd->my_flag_hi + (e & 1), e, pd->err);
  }
  __syncthreads();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (pd->my_inbox_lo)   // written by a peer: read past the L1 (a line of the slot's previous use may still sit there)
    for (int64_t i = tid; i < halo_lo; i += nth) x[i] = __ldcg(pd->my_inbox_lo + slot + i);
  if (pd->my_inbox_hi)
    for (int64_t i = tid; i < halo_hi; i += nth) x[f1 + i] = __ldcg(pd->my_inbox_hi + slot + i);
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(pd->halo_ticket + 1, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {   // every block has read the counter (it is read before the ticket is taken)
    pd->halo_ticket[1] = 0u;
    *pd->halo_epoch = e;
  }
}
static int comm_halo_exchange_inbox(apdx_plan *pl, double *x_d, cudaStream_t s) {
  const P2PDev *pd = pl->p2p.dev;
  k_halo_push<<<HALO_BLOCKS, 256, 0, s>>>(x_d, pl->f0, pl->f1, pl->rank_lo >= 0 ? pl->send_lo : 0,
                                          pl->rank_hi >= 0 ? pl->send_hi : 0, pd);
  k_halo_pull<<<HALO_BLOCKS, 256, 0, s>>>(x_d, pl->f1, pl->rank_lo >= 0 ? pl->halo_lo : 0, pl->rank_hi >= 0 ? pl->halo_hi : 0, pd);
  pl->stats.kernel_launches += 2;
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

static int comm_halo_exchange_nccl(apdx_plan *pl, double *x_d, cudaStream_t s);

// x_d is a vector in the local reduced numbering: [ghost_lo | owned | ghost_hi] (slabs) or [owned | ghosts by owner]
int comm_halo_exchange(apdx_plan *pl, double *x_d, cudaStream_t s) {
  if (pl->hl.active) return comm_halo_exchange_lists(pl, x_d, s);
  if (pl->rank_lo < 0 && pl->rank_hi < 0) return APDX_OK;
  if (pl->p2p.inbox) return comm_halo_exchange_inbox(pl, x_d, s);
  return comm_halo_exchange_nccl(pl, x_d, s);
}
static int comm_halo_exchange_nccl(apdx_plan *pl, double *x_d, cudaStream_t s) {
  APDX_NCCL(g_nccl.GroupStart());
  if (pl->rank_lo >= 0) {
    if (pl->send_lo > 0) APDX_NCCL(g_nccl.Send(x_d + pl->f0, (size_t)pl->send_lo, ncclFloat64, pl->rank_lo, g_nccl.comm, s));
    if (pl->halo_lo > 0) APDX_NCCL(g_nccl.Recv(x_d, (size_t)pl->halo_lo, ncclFloat64, pl->rank_lo, g_nccl.comm, s));
  }
  if (pl->rank_hi >= 0) {
    if (pl->send_hi > 0) APDX_NCCL(g_nccl.Send(x_d + pl->f1 - pl->send_hi, (size_t)pl->send_hi, ncclFloat64, pl->rank_hi, g_nccl.comm, s));
    if (pl->halo_hi > 0) APDX_NCCL(g_nccl.Recv(x_d + pl->f1, (size_t)pl->halo_hi, ncclFloat64, pl->rank_hi, g_nccl.comm, s));
  }
  APDX_NCCL(g_nccl.GroupEnd());
  return APDX_OK;
}

// Slab neighbours swap boundary blocks of ANY device vector laid out [ghost_lo | owned | ghost_hi] (full-dof vectors:
// the injected state of a coarse multigrid level, coordinates / masks of a partitioned hierarchy): the first `lo` owned
// entries go to rank_lo and its answer fills [0, lo); the last `hi` owned entries go to rank_hi and its answer fills
// [n - hi, n).  Both sides of an interface must pass the same count (one node plane).
int comm_exchange_planes(double *buf_d, int64_t n, int64_t lo, int64_t hi, int rank_lo, int rank_hi, cudaStream_t s) {
  if (!comm_active()) return APDX_OK;
  APDX_REQUIRE(lo >= 0 && hi >= 0 && n - lo - hi >= lo && n - lo - hi >= hi, APDX_ERR_INVALID,
               "plane exchange: the owned block of a vector of %lld entries with %lld + %lld ghost entries is smaller than a plane",
               (long long)n, (long long)lo, (long long)hi);
  if ((rank_lo < 0 || lo == 0) && (rank_hi < 0 || hi == 0)) return APDX_OK;
  APDX_NCCL(g_nccl.GroupStart());
  if (rank_lo >= 0 && lo > 0) {
    APDX_NCCL(g_nccl.Send(buf_d + lo, (size_t)lo, ncclFloat64, rank_lo, g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf_d, (size_t)lo, ncclFloat64, rank_lo, g_nccl.comm, s));
  }
  if (rank_hi >= 0 && hi > 0) {
    APDX_NCCL(g_nccl.Send(buf_d + n - 2 * hi, (size_t)hi, ncclFloat64, rank_hi, g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf_d + n - hi, (size_t)hi, ncclFloat64, rank_hi, g_nccl.comm, s));
  }
  APDX_NCCL(g_nccl.GroupEnd());
  return APDX_OK;
}

// neighbours tell each other how many owned entries the other side ghosts, and where their owned range ends
int comm_halo_setup(apdx_plan *pl) {
  int64_t *buf = nullptr;
  APDX_CUDA(cudaMalloc((void **)&buf, 8 * sizeof(int64_t)));
  // send block: [halo_lo, f1] to the lower neighbour, [halo_hi, f1] to the upper one
  int64_t h[8] = {pl->halo_lo, pl->f1, pl->halo_hi, pl->f1, 0, 0, 0, 0};
  APDX_CUDA(cudaMemcpy(buf, h, sizeof(h), cudaMemcpyHostToDevice));
  cudaStream_t s = pl->stream;
  APDX_NCCL(g_nccl.GroupStart());
  if (pl->rank_lo >= 0) {
    APDX_NCCL(g_nccl.Send(buf + 0, 2, ncclInt64, pl->rank_lo, g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf + 4, 2, ncclInt64, pl->rank_lo, g_nccl.comm, s));
  }
  if (pl->rank_hi >= 0) {
    APDX_NCCL(g_nccl.Send(buf + 2, 2, ncclInt64, pl->rank_hi, g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf + 6, 2, ncclInt64, pl->rank_hi, g_nccl.comm, s));
  }
  APDX_NCCL(g_nccl.GroupEnd());
  APDX_CUDA(cudaStreamSynchronize(s));
  APDX_CUDA(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(buf);
  pl->send_lo = pl->rank_lo >= 0 ? h[4] : 0;  // lower neighbour's upper-ghost count
  pl->send_hi = pl->rank_hi >= 0 ? h[6] : 0;  // upper neighbour's lower-ghost count
  APDX_REQUIRE(pl->send_lo <= pl->f1 - pl->f0 && pl->send_hi <= pl->f1 - pl->f0, APDX_ERR_INVALID,
               "neighbour ghosts more dofs than this rank owns");
  return APDX_OK;
}

// ---- peer memory: CUDA IPC mapping of every rank's mailbox / flag block ----------------------------------------
constexpr size_t P2P_HDR = 4096;
static inline double *hdr_mbox(void *base) { return reinterpret_cast<double *>(base); }
static inline int *hdr_mflag(void *base) { return reinterpret_cast<int *>(static_cast<char *>(base) + 2 * P2P_MAX_RANKS * 4 * sizeof(double)); }
static inline int *hdr_err(void *base) { return hdr_mflag(base) + 2 * P2P_MAX_RANKS; }
static inline int *hdr_epoch(void *base) { return hdr_err(base) + 1; }
// halo part of the block: flags and counters at byte 2048, the inboxes behind the 4 KB header
static inline int *hdr_hflag_lo(void *base) { return reinterpret_cast<int *>(static_cast<char *>(base) + 2048); }   // [2 slots]
static inline int *hdr_hflag_hi(void *base) { return hdr_hflag_lo(base) + 2; }                                      // [2 slots]
static inline int *hdr_hepoch(void *base) { return hdr_hflag_lo(base) + 4; }
static inline unsigned int *hdr_hticket(void *base) { return reinterpret_cast<unsigned int *>(hdr_hflag_lo(base) + 5); }  // [2]
static inline double *hdr_inbox_lo(void *base) { return reinterpret_cast<double *>(static_cast<char *>(base) + P2P_HDR); }
static inline double *hdr_inbox_hi(void *base, int64_t cap) { return hdr_inbox_lo(base) + 2 * cap; }

__global__ void k_halo_test_fill(double *__restrict__ a, double *__restrict__ b, int64_t n, double rank) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    a[i] = b[i] = rank * 1.0e7 + (double)i;
}

// A block may still be mapped by the peers when its plan dies, so it is only parked here; apdx_comm_destroy
// (a collective call) frees the parked blocks after a barrier.
static std::vector<P2P> g_parked;

void p2p_teardown(apdx_plan *pl) {
  P2P &P = pl->p2p;
  if (!P.heap) return;
  g_parked.push_back(P);
  P = P2P();
}

static void p2p_free_parked() {
  for (P2P &P : g_parked) {
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
      if (P.peer_base[r] && P.peer_base[r] != P.heap) cudaIpcCloseMemHandle(P.peer_base[r]);
    if (P.dev) cudaFree(P.dev);
    cudaFree(P.heap);
  }
  g_parked.clear();
}

// Decide, on the machine the plan runs on, whether its halo exchange goes through the peer inboxes: one exchange of a
// rank- and index-dependent pattern through NCCL and one through the inboxes must deliver identical ghost entries on
// every rank (and no wait may time out); then both are timed (CUDA events, maximum over the ranks) and the faster one
// is kept.  `force`: APDX_HALO=inbox keeps the inboxes whenever they are correct.  Collective.
static int halo_inbox_selftest(apdx_plan *pl, bool force) {
  P2P &P = pl->p2p;
  cudaStream_t s = pl->stream;
  const int64_t n = pl->n_free;
  const int64_t hl = pl->rank_lo >= 0 ? pl->halo_lo : 0, hh = pl->rank_hi >= 0 ? pl->halo_hi : 0;
  double *xa = nullptr, *xb = nullptr;
  APDX_CUDA(cudaMalloc((void **)&xa, (size_t)n * sizeof(double)));
  APDX_CUDA(cudaMalloc((void **)&xb, (size_t)n * sizeof(double)));
  k_halo_test_fill<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8), 256, 0, s>>>(xa, xb, n, (double)g_nccl.rank);
  int rc = comm_halo_exchange_nccl(pl, xa, s);
  if (rc == APDX_OK) rc = comm_halo_exchange_inbox(pl, xb, s);
  if (rc == APDX_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = APDX_ERR_CUDA;
  double v[3] = {0.0, 0.0, 0.0};   // (bad, inbox us, nccl us)
  if (rc == APDX_OK) {
    std::vector<double> ga((size_t)(hl + hh)), gb((size_t)(hl + hh));
    if (hl > 0) { cudaMemcpy(ga.data(), xa, hl * sizeof(double), cudaMemcpyDeviceToHost); cudaMemcpy(gb.data(), xb, hl * sizeof(double), cudaMemcpyDeviceToHost); }
    if (hh > 0) { cudaMemcpy(ga.data() + hl, xa + pl->f1, hh * sizeof(double), cudaMemcpyDeviceToHost); cudaMemcpy(gb.data() + hl, xb + pl->f1, hh * sizeof(double), cudaMemcpyDeviceToHost); }
    int err = 0;
    cudaMemcpy(&err, P.err_d, sizeof(int), cudaMemcpyDeviceToHost);
    if (err || (!ga.empty() && memcmp(ga.data(), gb.data(), ga.size() * sizeof(double)) != 0)) v[0] = 1.0;
    if (err) cudaMemset(P.err_d, 0, sizeof(int));
    const int reps = 20;
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaEventRecord(e0, s);
    for (int i = 0; i < reps && rc == APDX_OK; ++i) rc = comm_halo_exchange_inbox(pl, xb, s);
    cudaEventRecord(e1, s);
    for (int i = 0; i < reps && rc == APDX_OK; ++i) rc = comm_halo_exchange_nccl(pl, xa, s);
    cudaEventRecord(e2, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = APDX_ERR_CUDA;
    float t_in = 0.f, t_nc = 0.f;
    cudaEventElapsedTime(&t_in, e0, e1);
    cudaEventElapsedTime(&t_nc, e1, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    v[1] = 1e3 * t_in / reps; v[2] = 1e3 * t_nc / reps;
    cudaMemcpy(&err, P.err_d, sizeof(int), cudaMemcpyDeviceToHost);
    if (err) { v[0] = 1.0; cudaMemset(P.err_d, 0, sizeof(int)); }
  }
  cudaFree(xa);
  cudaFree(xb);
  cudaGetLastError();
  if (rc != APDX_OK) v[0] = 1.0;   // still take part in the collective decision
  double *dv = nullptr;
  APDX_CUDA(cudaMalloc((void **)&dv, sizeof(v)));
  APDX_CUDA(cudaMemcpy(dv, v, sizeof(v), cudaMemcpyHostToDevice));
  APDX_NCCL(g_nccl.AllReduce(dv, dv, 3, ncclFloat64, ncclMax, g_nccl.comm, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  APDX_CUDA(cudaMemcpy(v, dv, sizeof(v), cudaMemcpyDeviceToHost));
  cudaFree(dv);
  P.inbox_us = v[1]; P.nccl_us = v[2];
  P.inbox = v[0] == 0.0 && (force || v[1] < v[2]);
  if (g_nccl.rank == 0 && getenv("APDX_TRACE"))
    fprintf(stderr, "[apdx trace] halo exchange of %lld + %lld entries: peer inboxes %.1f us, nccl %.1f us, %s -> %s\n", (long long)hl,
            (long long)hh, v[1], v[2], v[0] == 0.0 ? "identical ghost entries" : "MISMATCH or time-out", P.inbox ? "inboxes" : "nccl");
  return APDX_OK;
}

// Dot-product all-reduces of the Krylov loop through peer-memory mailboxes over NVLink (k_allreduce_mbox_apply,
// krylov.cu) -- the default since round 2: 256^3 Newton step 145.3 -> 143.7 ms on 2, 89.3 -> 86.1 ms on 4 and
// 65.9 -> 58.0 ms on 8 B200s against ncclAllReduce + a one-thread stage kernel (profiles/r02b_bench_n{2,4,8}*_sample.json).
// APDX_COMM=nccl keeps the NCCL all-reduce; if CUDA IPC peer mapping is unavailable every rank stays on NCCL (collective
// decision below).  Collective: every rank of the communicator must call it.
int p2p_setup(apdx_plan *pl) {
  P2P &P = pl->p2p;
  const char *mode = getenv("APDX_COMM");
  const bool mbox = !(mode && strcmp(mode, "nccl") == 0);
  if (!mbox) return APDX_OK;
  if (g_nccl.nranks > P2P_MAX_RANKS) return APDX_OK;
  p2p_teardown(pl);
  cudaStream_t s = pl->stream;
  const int me = g_nccl.rank, nr = g_nccl.nranks;
  // halo inboxes (slab partitions; APDX_HALO=nccl keeps ncclSend/ncclRecv): one capacity for every rank, so that the
  // block has the same layout everywhere
  const char *hmode = getenv("APDX_HALO");
  int64_t cap = 0;
  {
    double need = (!pl->hl.active && !(hmode && strcmp(hmode, "nccl") == 0))
                      ? (double)std::max(pl->rank_lo >= 0 ? pl->halo_lo : 0, pl->rank_hi >= 0 ? pl->halo_hi : 0) : 0.0;
    double *dn = nullptr;
    APDX_CUDA(cudaMalloc((void **)&dn, sizeof(double)));
    APDX_CUDA(cudaMemcpy(dn, &need, sizeof(double), cudaMemcpyHostToDevice));
    APDX_NCCL(g_nccl.AllReduce(dn, dn, 1, ncclFloat64, ncclMax, g_nccl.comm, s));
    APDX_CUDA(cudaStreamSynchronize(s));
    APDX_CUDA(cudaMemcpy(&need, dn, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dn);
    cap = (int64_t)need;
  }
  P.heap_bytes = (P2P_HDR + 4 * (size_t)cap * sizeof(double) + 4095) & ~(size_t)4095;
  APDX_CUDA(cudaMalloc((void **)&P.heap, P.heap_bytes));
  APDX_CUDA(cudaMemset(P.heap, 0, P.heap_bytes));
  // exchange IPC handles
  cudaIpcMemHandle_t mine;
  cudaError_t ce = cudaIpcGetMemHandle(&mine, P.heap);
  char *hd = nullptr;
  APDX_CUDA(cudaMalloc((void **)&hd, (size_t)(nr + 1) * sizeof(cudaIpcMemHandle_t)));
  int ok_local = (ce == cudaSuccess) ? 1 : 0;
  if (!ok_local) {
    cudaGetLastError();   // the fallback is not an error: do not leave it for the next cudaGetLastError() check
    memset(&mine, 0, sizeof(mine));
  }
  APDX_CUDA(cudaMemcpy(hd + (size_t)nr * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
  APDX_NCCL(g_nccl.AllGather(hd + (size_t)nr * sizeof(mine), hd, sizeof(mine), 0 /*ncclInt8*/, g_nccl.comm, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  std::vector<cudaIpcMemHandle_t> all(nr);
  APDX_CUDA(cudaMemcpy(all.data(), hd, (size_t)nr * sizeof(mine), cudaMemcpyDeviceToHost));
  cudaFree(hd);
  for (int r = 0; r < nr && ok_local; ++r) {
    if (r == me) { P.peer_base[r] = P.heap; continue; }
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok_local = 0;
      break;
    }
    P.peer_base[r] = ptr;
  }
  // everybody must succeed, otherwise everybody stays on the NCCL path
  double *dok = nullptr;
  APDX_CUDA(cudaMalloc((void **)&dok, sizeof(double)));
  double okv = ok_local ? 0.0 : 1.0;
  APDX_CUDA(cudaMemcpy(dok, &okv, sizeof(double), cudaMemcpyHostToDevice));
  APDX_NCCL(g_nccl.AllReduce(dok, dok, 1, ncclFloat64, ncclSum, g_nccl.comm, s));
  APDX_CUDA(cudaStreamSynchronize(s));
  APDX_CUDA(cudaMemcpy(&okv, dok, sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(dok);
  if (okv != 0.0) {
    if (me == 0) fprintf(stderr, "[apdx_b200] CUDA IPC peer mapping unavailable (%s): Krylov loop stays on NCCL\n",
                         cudaGetErrorString(ce));
    for (int r = 0; r < nr; ++r)
      if (P.peer_base[r] && P.peer_base[r] != P.heap) cudaIpcCloseMemHandle(P.peer_base[r]);
    cudaFree(P.heap);
    P = P2P();
    return APDX_OK;
  }
  P.err_d = hdr_err(P.heap);
  P2PDev d{};
  d.rank = me; d.nranks = nr;
  for (int r = 0; r < nr; ++r) { d.mbox[r] = hdr_mbox(P.peer_base[r]); d.mflag[r] = hdr_mflag(P.peer_base[r]); }
  d.err = P.err_d;
  d.epoch_self = hdr_epoch(P.heap);
  d.cap = cap;
  d.halo_epoch = hdr_hepoch(P.heap);
  d.halo_ticket = hdr_hticket(P.heap);
  if (cap > 0) {
    if (pl->rank_lo >= 0) {   // my lowest owned entries are rank_lo's UPPER ghosts, and the other way round
      d.peer_inbox_lo = hdr_inbox_hi(P.peer_base[pl->rank_lo], cap); d.peer_flag_lo = hdr_hflag_hi(P.peer_base[pl->rank_lo]);
      d.my_inbox_lo = hdr_inbox_lo(P.heap); d.my_flag_lo = hdr_hflag_lo(P.heap);
    }
    if (pl->rank_hi >= 0) {
      d.peer_inbox_hi = hdr_inbox_lo(P.peer_base[pl->rank_hi]); d.peer_flag_hi = hdr_hflag_lo(P.peer_base[pl->rank_hi]);
      d.my_inbox_hi = hdr_inbox_hi(P.heap, cap); d.my_flag_hi = hdr_hflag_hi(P.heap);
    }
  }
  APDX_CUDA(cudaMalloc((void **)&P.dev, sizeof(P2PDev)));
  APDX_CUDA(cudaMemcpy(P.dev, &d, sizeof(P2PDev), cudaMemcpyHostToDevice));
  P.mbox = true;
  if (cap > 0) APDX_CHECK(halo_inbox_selftest(pl, hmode && strcmp(hmode, "inbox") == 0));
  return APDX_OK;
}

}  // namespace apdx

using namespace apdx;

extern "C" {

int apdx_comm_unique_id(uint8_t id_out[128]) {
  APDX_CHECK(load_nccl());
  ncclUniqueId id;
  APDX_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return APDX_OK;
}

int apdx_comm_init(const uint8_t id_in[128], int32_t rank, int32_t nranks) {
  APDX_CHECK(load_nccl());
  APDX_REQUIRE(!g_nccl.comm, APDX_ERR_STATE, "communicator already initialised");
  APDX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, APDX_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  ncclUniqueId id;
  memcpy(id.internal, id_in, 128);
  APDX_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
  g_nccl.rank = rank;
  g_nccl.nranks = nranks;
  return APDX_OK;
}

int apdx_comm_allreduce_host(double *inout_h, int32_t count, int32_t op) {
  APDX_REQUIRE(inout_h && count > 0, APDX_ERR_INVALID, "bad argument");
  if (!comm_active()) return APDX_OK;
  double *buf = nullptr;
  APDX_CUDA(cudaMalloc((void **)&buf, count * sizeof(double)));
  APDX_CUDA(cudaMemcpy(buf, inout_h, count * sizeof(double), cudaMemcpyHostToDevice));
  APDX_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, op == 1 ? ncclMax : ncclSum, g_nccl.comm, 0));
  APDX_CUDA(cudaStreamSynchronize(0));
  APDX_CUDA(cudaMemcpy(inout_h, buf, count * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(buf);
  return APDX_OK;
}

int apdx_comm_exchange_planes(double *buf_d, int64_t n, int64_t lo_count, int64_t hi_count, int32_t rank_lo, int32_t rank_hi) {
  APDX_REQUIRE(buf_d || n == 0, APDX_ERR_INVALID, "NULL argument");
  APDX_CHECK(comm_exchange_planes(buf_d, n, lo_count, hi_count, rank_lo, rank_hi, 0));
  APDX_CUDA(cudaStreamSynchronize(0));
  return APDX_OK;
}

int apdx_comm_destroy(void) {
  if (g_nccl.comm) {
    drop_all_krylov_graphs();   // captured graphs hold NCCL operations of this communicator
    cudaDeviceSynchronize();
    if (g_nccl.nranks > 1) {  // barrier on EVERY rank (parked state may differ): no rank may still be storing into a block that is about to be freed
      double *b = nullptr;
      APDX_CUDA(cudaMalloc((void **)&b, sizeof(double)));
      APDX_CUDA(cudaMemset(b, 0, sizeof(double)));
      APDX_NCCL(g_nccl.AllReduce(b, b, 1, ncclFloat64, ncclSum, g_nccl.comm, 0));
      APDX_CUDA(cudaStreamSynchronize(0));
      cudaFree(b);
      p2p_free_parked();
    }
    APDX_NCCL(g_nccl.CommDestroy(g_nccl.comm));
    g_nccl.comm = nullptr;
    g_nccl.nranks = 1;
  }
  return APDX_OK;
}

}  // extern "C"
