// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.  The reference has no
// distributed code at all; this implements SURVEY.md 8e: slab partitions whose
// ghost planes are contiguous ranges of the reduced numbering, exchanged with grouped
// ncclSend/ncclRecv before every SpMV, and ncclAllReduce for dot products / the Newton norm.
// NCCL is loaded with dlopen so that the single-GPU library has no link-time dependency on it.
#include <dlfcn.h>
#include <string.h>

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
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
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

// x_d is a vector in the local reduced numbering: [ghost_lo | owned | ghost_hi]
int comm_halo_exchange(apdx_plan *pl, double *x_d, cudaStream_t s) {
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

// neighbours tell each other how many owned entries the other side ghosts
int comm_halo_setup(apdx_plan *pl) {
  int64_t *buf = nullptr;
  APDX_CUDA(cudaMalloc((void **)&buf, 4 * sizeof(int64_t)));
  int64_t h[4] = {pl->halo_lo, pl->halo_hi, 0, 0};
  APDX_CUDA(cudaMemcpy(buf, h, sizeof(h), cudaMemcpyHostToDevice));
  cudaStream_t s = pl->stream;
  APDX_NCCL(g_nccl.GroupStart());
  if (pl->rank_lo >= 0) {
    APDX_NCCL(g_nccl.Send(buf + 0, 1, ncclInt64, pl->rank_lo, g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf + 2, 1, ncclInt64, pl->rank_lo, g_nccl.comm, s));
  }
  if (pl->rank_hi >= 0) {
    APDX_NCCL(g_nccl.Send(buf + 1, 1, ncclInt64, pl->rank_hi, g_nccl.comm, s));
    APDX_NCCL(g_nccl.Recv(buf + 3, 1, ncclInt64, pl->rank_hi, g_nccl.comm, s));
  }
  APDX_NCCL(g_nccl.GroupEnd());
  APDX_CUDA(cudaStreamSynchronize(s));
  APDX_CUDA(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(buf);
  pl->send_lo = pl->rank_lo >= 0 ? h[2] : 0;  // lower neighbour's upper-ghost count
  pl->send_hi = pl->rank_hi >= 0 ? h[3] : 0;  // upper neighbour's lower-ghost count
  APDX_REQUIRE(pl->send_lo <= pl->f1 - pl->f0 && pl->send_hi <= pl->f1 - pl->f0, APDX_ERR_INVALID,
               "neighbour ghosts more dofs than this rank owns");
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

int apdx_comm_destroy(void) {
  if (g_nccl.comm) {
    APDX_NCCL(g_nccl.CommDestroy(g_nccl.comm));
    g_nccl.comm = nullptr;
    g_nccl.nranks = 1;
  }
  return APDX_OK;
}

}  // extern "C"
