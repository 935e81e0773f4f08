// TEST INFRASTRUCTURE (see include/cuda_runtime.h): a stand-in for libnccl.so.2 so that the multi-GPU code of the product
// (dist.cu: halo exchange, all-reduces, partition set-up; krylov.cu: the multi-rank Krylov loop and its CUDA graphs) can
// run as N host processes against the emulated build.  The product dlopen()s NCCL at run time; the multi-rank emulated
// runs point APDX_NCCL_LIB at this library.
//
// Ranks are processes connected by a full mesh of Unix-domain sockets (rendezvous directory named by the unique id).
// Operations are enqueued on the emulated stream: executed at once, or recorded while the stream captures and replayed
// by cudaGraphLaunch -- like the real library.  All-reduces sum in rank order on rank 0 (identical bits on every rank).
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct ncclComm {
  int rank = 0, nranks = 1;
  std::vector<int> fd;   // socket to every peer (-1 for self)
  std::string dir;
};
typedef ncclComm *ncclComm_t;

typedef int (*enqueue_fn)(void *stream, void (*fn)(void *), void *arg);
static enqueue_fn g_enqueue = nullptr;

static bool find_enqueue() {
  if (g_enqueue) return true;
  const char *lib = getenv("APDX_LIB");
  void *h = lib ? dlopen(lib, RTLD_NOW | RTLD_NOLOAD) : nullptr;
  if (!h) h = dlopen(nullptr, RTLD_NOW);
  g_enqueue = h ? (enqueue_fn)dlsym(h, "emu_stream_enqueue") : nullptr;
  return g_enqueue != nullptr;
}

static size_t dtype_size(int dt) {
  switch (dt) {
    case 0: case 1: return 1;          // int8 / uint8
    case 2: case 3: case 7: return 4;  // int32 / uint32 / float32
    case 4: case 5: case 8: return 8;  // int64 / uint64 / float64
    case 6: return 2;                  // float16
    default: return 0;
  }
}

static void die(const char *what) {
  fprintf(stderr, "fake nccl: %s: %s\n", what, strerror(errno));
  _exit(70);
}
static void write_all(int fd, const void *p, size_t n) {
  const char *c = (const char *)p;
  while (n) {
    ssize_t k = send(fd, c, n, MSG_NOSIGNAL);
    if (k < 0) { if (errno == EINTR) continue; die("send"); }
    c += k; n -= (size_t)k;
  }
}
static void read_all(int fd, void *p, size_t n) {
  char *c = (char *)p;
  while (n) {
    ssize_t k = recv(fd, c, n, 0);
    if (k == 0) { fprintf(stderr, "fake nccl: peer closed the connection\n"); _exit(71); }
    if (k < 0) { if (errno == EINTR) continue; die("recv"); }
    c += k; n -= (size_t)k;
  }
}

extern "C" {

const char *ncclGetErrorString(ncclResult_t r) { return r == 0 ? "no error" : "fake nccl error"; }

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof(*id));
  timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  snprintf(id->internal, sizeof(id->internal), "/tmp/emunccl-%d-%lx", (int)getpid(), (unsigned long)ts.tv_nsec);
  return 0;
}

ncclResult_t ncclCommInitRank(ncclComm_t *out, int nranks, ncclUniqueId id, int rank) {
  if (!find_enqueue()) { fprintf(stderr, "fake nccl: emu_stream_enqueue not found (set APDX_LIB to the emulated build)\n"); return 1; }
  ncclComm *c = new ncclComm();
  c->rank = rank; c->nranks = nranks; c->fd.assign((size_t)nranks, -1);
  c->dir = std::string(id.internal, strnlen(id.internal, sizeof(id.internal)));
  mkdir(c->dir.c_str(), 0700);
  auto path_of = [&](int r) { return c->dir + "/r" + std::to_string(r); };
  int lfd = socket(AF_UNIX, SOCK_STREAM, 0);
  if (lfd < 0) die("socket");
  sockaddr_un a{};
  a.sun_family = AF_UNIX;
  snprintf(a.sun_path, sizeof(a.sun_path), "%s", path_of(rank).c_str());
  unlink(a.sun_path);
  if (bind(lfd, (sockaddr *)&a, sizeof(a)) != 0) die("bind");
  if (listen(lfd, nranks) != 0) die("listen");
  for (int r = 0; r < rank; ++r) {   // connect to the lower ranks
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un b{};
    b.sun_family = AF_UNIX;
    snprintf(b.sun_path, sizeof(b.sun_path), "%s", path_of(r).c_str());
    int tries = 0;
    while (connect(fd, (sockaddr *)&b, sizeof(b)) != 0) {
      if (++tries > 3000) die("connect");
      usleep(10000);
    }
    int32_t me = rank;
    write_all(fd, &me, sizeof(me));
    c->fd[(size_t)r] = fd;
  }
  for (int k = rank + 1; k < nranks; ++k) {   // accept the higher ranks
    int fd = accept(lfd, nullptr, nullptr);
    if (fd < 0) die("accept");
    int32_t who = -1;
    read_all(fd, &who, sizeof(who));
    c->fd[(size_t)who] = fd;
  }
  close(lfd);
  unlink(a.sun_path);
  *out = c;
  return 0;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) {
  if (!c) return 0;
  for (int fd : c->fd) if (fd >= 0) close(fd);
  rmdir(c->dir.c_str());
  delete c;
  return 0;
}

}  // extern "C"

// ---- operations ------------------------------------------------------------------------------------------------------
struct Op { virtual void run() = 0; virtual ~Op() {} };
static void run_op(void *p) { static_cast<Op *>(p)->run(); }
static ncclResult_t enqueue(void *stream, Op *op) {
  const int kept = g_enqueue(stream, run_op, op);   // 1 = recorded into a capturing stream's graph (may run many times)
  if (!kept) delete op;
  return 0;
}

struct AllReduceOp : Op {
  ncclComm *c; const void *send; void *recv; size_t count; int dt, op;
  void run() override {
    const size_t es = dtype_size(dt), bytes = count * es;
    if (recv != send) memmove(recv, send, bytes);
    if (c->nranks == 1) return;
    if (c->rank != 0) {
      write_all(c->fd[0], recv, bytes);
      read_all(c->fd[0], recv, bytes);
      return;
    }
    std::vector<char> tmp(bytes);
    for (int r = 1; r < c->nranks; ++r) {
      read_all(c->fd[(size_t)r], tmp.data(), bytes);
      for (size_t i = 0; i < count; ++i) {
        if (dt == 8) {
          double *a = (double *)recv + i, b = ((const double *)tmp.data())[i];
          *a = op == 2 ? (*a < b ? b : *a) : *a + b;
        } else if (dt == 4) {
          int64_t *a = (int64_t *)recv + i, b = ((const int64_t *)tmp.data())[i];
          *a = op == 2 ? (*a < b ? b : *a) : *a + b;
        } else {
          fprintf(stderr, "fake nccl: all-reduce of dtype %d not implemented\n", dt);
          _exit(72);
        }
      }
    }
    for (int r = 1; r < c->nranks; ++r) write_all(c->fd[(size_t)r], recv, bytes);
  }
};

struct AllGatherOp : Op {
  ncclComm *c; const void *send; void *recv; size_t bytes;
  void run() override {
    char *out = (char *)recv;
    memmove(out + (size_t)c->rank * bytes, send, bytes);
    if (c->nranks == 1) return;
    const size_t total = bytes * (size_t)c->nranks;
    if (c->rank != 0) {
      write_all(c->fd[0], out + (size_t)c->rank * bytes, bytes);
      read_all(c->fd[0], out, total);
      return;
    }
    for (int r = 1; r < c->nranks; ++r) read_all(c->fd[(size_t)r], out + (size_t)r * bytes, bytes);
    for (int r = 1; r < c->nranks; ++r) write_all(c->fd[(size_t)r], out, total);
  }
};

struct P2POp { bool is_send; char *buf; size_t bytes, done; int peer; };
struct GroupOp : Op {
  ncclComm *c = nullptr;
  std::vector<P2POp> ops;
  void run() override {
    for (auto &o : ops) o.done = 0;
    // progress every (peer, direction) queue in order, without blocking on any single one
    while (true) {
      std::vector<pollfd> pf;
      std::vector<size_t> which;
      std::vector<char> busy_s((size_t)c->nranks, 0), busy_r((size_t)c->nranks, 0);
      for (size_t i = 0; i < ops.size(); ++i) {
        P2POp &o = ops[i];
        if (o.done == o.bytes) continue;
        char &busy = (o.is_send ? busy_s : busy_r)[(size_t)o.peer];
        if (busy) continue;   // an earlier message of the same queue is still in flight
        busy = 1;
        pf.push_back(pollfd{c->fd[(size_t)o.peer], (short)(o.is_send ? POLLOUT : POLLIN), 0});
        which.push_back(i);
      }
      if (pf.empty()) break;
      if (poll(pf.data(), pf.size(), 60000) <= 0) { fprintf(stderr, "fake nccl: send/recv group timed out on rank %d\n", c->rank); _exit(73); }
      for (size_t k = 0; k < pf.size(); ++k) {
        if (!(pf[k].revents & (POLLIN | POLLOUT | POLLHUP | POLLERR))) continue;
        P2POp &o = ops[which[k]];
        ssize_t n = o.is_send ? send(pf[k].fd, o.buf + o.done, o.bytes - o.done, MSG_DONTWAIT | MSG_NOSIGNAL)
                              : recv(pf[k].fd, o.buf + o.done, o.bytes - o.done, MSG_DONTWAIT);
        if (n < 0) { if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) continue; die("group send/recv"); }
        if (n == 0 && !o.is_send) { fprintf(stderr, "fake nccl: peer closed during a group\n"); _exit(71); }
        o.done += (size_t)n;
      }
    }
  }
};

static int g_group_depth = 0;
static GroupOp *g_group = nullptr;
static void *g_group_stream = nullptr;

extern "C" {

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, int dt, int op, ncclComm_t c, void *stream) {
  AllReduceOp *o = new AllReduceOp();
  o->c = c; o->send = send; o->recv = recv; o->count = count; o->dt = dt; o->op = op;
  return enqueue(stream, o);
}
ncclResult_t ncclAllGather(const void *send, void *recv, size_t sendcount, int dt, ncclComm_t c, void *stream) {
  AllGatherOp *o = new AllGatherOp();
  o->c = c; o->send = send; o->recv = recv; o->bytes = sendcount * dtype_size(dt);
  return enqueue(stream, o);
}
ncclResult_t ncclGroupEnd();
ncclResult_t ncclGroupStart() {
  if (g_group_depth++ == 0) { g_group = new GroupOp(); g_group_stream = nullptr; }
  return 0;
}
static ncclResult_t p2p(bool is_send, void *buf, size_t count, int dt, int peer, ncclComm_t c, void *stream) {
  if (peer < 0 || peer >= c->nranks || peer == c->rank) return 2;
  const bool own_group = g_group_depth == 0;
  if (own_group) ncclGroupStart();
  g_group->c = c;
  g_group_stream = stream;
  g_group->ops.push_back(P2POp{is_send, (char *)buf, count * dtype_size(dt), 0, peer});
  return own_group ? ncclGroupEnd() : 0;
}
ncclResult_t ncclGroupEnd() {
  if (--g_group_depth > 0) return 0;
  GroupOp *g = g_group;
  g_group = nullptr;
  if (!g->c || g->ops.empty()) { delete g; return 0; }
  return enqueue(g_group_stream, g);
}
ncclResult_t ncclSend(const void *buf, size_t count, int dt, int peer, ncclComm_t c, void *stream) {
  return p2p(true, const_cast<void *>(buf), count, dt, peer, c, stream);
}
ncclResult_t ncclRecv(void *buf, size_t count, int dt, int peer, ncclComm_t c, void *stream) {
  return p2p(false, buf, count, dt, peer, c, stream);
}

}  // extern "C"
