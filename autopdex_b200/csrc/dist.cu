// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.  The reference has no
// distributed code at all; this implements SURVEY.md 8e: slab partitions whose
// ghost planes are contiguous ranges of the reduced numbering, exchanged with grouped
// ncclSend/ncclRecv before every SpMV, and ncclAllReduce for dot products / the Newton norm.
// NCCL is loaded with dlopen so that the single-GPU library has no link-time dependency on it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

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

// x_d is a vector in the local reduced numbering: [ghost_lo | owned | ghost_hi] (slabs) or [owned | ghosts by owner]
int comm_halo_exchange(apdx_plan *pl, double *x_d, cudaStream_t s) {
  if (pl->hl.active) return comm_halo_exchange_lists(pl, x_d, s);
  if (pl->rank_lo < 0 && pl->rank_hi < 0) return APDX_OK;
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
  P.heap_bytes = P2P_HDR;
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
  APDX_CUDA(cudaMalloc((void **)&P.dev, sizeof(P2PDev)));
  APDX_CUDA(cudaMemcpy(P.dev, &d, sizeof(P2PDev), cudaMemcpyHostToDevice));
  P.mbox = true;
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
