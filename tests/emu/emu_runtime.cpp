// TEST INFRASTRUCTURE (see include/cuda_runtime.h): the host-side half of the CUDA stand-in.
//   * fibers: one per thread of the running block, switched by a 12-instruction context switch; rendezvous points
//     (__syncthreads, shuffles, votes) park a fiber until every live thread of the block / warp has arrived;
//   * device memory: malloc'ed blocks filled with 0xFF (NaN doubles, -1 integers: nothing may rely on zeroed memory)
//     between red zones that are verified on free; EMU_GUARD=1 puts every block right in front of an inaccessible page
//     (reads or writes past the end fault, with the kernel / block / thread printed), EMU_GUARD=2 right behind one;
//   * streams are synchronous; capture records closures, cudaGraphLaunch replays them; cudaMalloc / cudaFree /
//     synchronous copies / synchronisation on a capturing stream return the error the real runtime returns.
#include <cuda_runtime.h>
#include <signal.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace emu {

uint3 tid{0, 0, 0}, bid{0, 0, 0};
dim3 bdim, gdim;

// ---------------------------------------------------------------------------------------------------------------------
// fibers
// ---------------------------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

enum State : int { NEW = 0, RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber {
  void *sp = nullptr;
  char *stack = nullptr;
  State state = DONE;
  unsigned wait_mask = 0;
};
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr int MAX_THREADS = 1024;
static Fiber g_fib[MAX_THREADS];
static void *g_sched_sp = nullptr;
static int g_cur = -1;            // running fiber (linear thread index in the block), -1 = scheduler / host code
static int g_nthreads = 0;
static const std::function<void()> *g_body = nullptr;
static const char *g_kernel = "(none)";
static uint64_t g_slots[MAX_THREADS / 32][32];
static std::vector<char> g_dyn_smem;
static char *g_dyn_guarded = nullptr;   // EMU_GUARD != 0: the dynamic shared memory of a block ends right in front of an inaccessible page
static char *g_dyn_ptr = nullptr;
static constexpr size_t DYN_REGION = 256 * 1024;
static long long g_clock = 0;
static long long g_strict_violations = 0;

static void set_tid(int t) {
  tid.x = (unsigned)t % bdim.x;
  tid.y = ((unsigned)t / bdim.x) % bdim.y;
  tid.z = (unsigned)t / (bdim.x * bdim.y);
}

static void fiber_entry() {
  (*g_body)();
  Fiber &f = g_fib[g_cur];
  f.state = DONE;
  emu_switch(&f.sp, g_sched_sp);
  abort();   // a finished fiber is never resumed
}

static void prepare(Fiber &f) {
  if (!f.stack) {
    void *m = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { perror("emu: mmap of a fiber stack"); abort(); }
    mprotect(m, 4096, PROT_NONE);   // stack overflow faults instead of corrupting the neighbour
    f.stack = (char *)m;
  }
  uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
  void **sp = (void **)top;
  *--sp = nullptr;                  // return address slot of fiber_entry (never used)
  *--sp = (void *)&fiber_entry;     // popped by the `ret` of emu_switch
  for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
  f.sp = sp;
  f.state = NEW;
  f.wait_mask = 0;
}

static void yield_to_scheduler(State st, unsigned mask) {
  if (g_cur < 0) { fprintf(stderr, "emu: device synchronisation called outside a kernel\n"); abort(); }
  Fiber &f = g_fib[g_cur];
  f.state = st;
  f.wait_mask = mask;
  emu_switch(&f.sp, g_sched_sp);
}

void sync_block() { yield_to_scheduler(WAIT_BLOCK, 0); }
void sync_warp(unsigned mask) { yield_to_scheduler(WAIT_WARP, mask); }
uint64_t *warp_slots() { return g_slots[g_cur >> 5]; }
int lane_id() { return g_cur & 31; }
void *dyn_smem() { return g_dyn_ptr; }
long long clock() {   // ~2 "cycles" per nanosecond of real time: device-side time-outs (peer waits) keep their meaning
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 2 * ((long long)ts.tv_sec * 1000000000ll + ts.tv_nsec) + (++g_clock & 1);
}
unsigned live_mask() {
  unsigned m = 0;
  const int w0 = g_cur & ~31;
  for (int l = 0; l < 32 && w0 + l < g_nthreads; ++l)
    if (g_fib[w0 + l].state != DONE) m |= 1u << l;
  return m;
}

static int exec_order() {   // 0 = ascending, 1 = reverse, 2 = shuffled blocks + reversed threads
  static int o = -1;
  if (o < 0) {
    const char *e = getenv("EMU_ORDER");
    o = !e ? 0 : (strcmp(e, "reverse") == 0 ? 1 : (strcmp(e, "shuffle") == 0 ? 2 : 0));
  }
  return o;
}

static void run_block() {
  const int n = g_nthreads;
  for (int t = 0; t < n; ++t) prepare(g_fib[t]);
  const int nwarps = (n + 31) / 32;
  int done = 0;
  while (done < n) {
    bool progressed = false;
    for (int tt = 0; tt < n; ++tt) {
      const int t = exec_order() ? n - 1 - tt : tt;
      Fiber &f = g_fib[t];
      if (f.state != NEW && f.state != RUNNABLE) continue;
      g_cur = t;
      set_tid(t);
      emu_switch(&g_sched_sp, f.sp);
      g_cur = -1;
      progressed = true;
      if (f.state == DONE) ++done;
    }
    if (done == n) break;
    // release complete rendezvous: block barrier = every live thread waits at it
    int at_block = 0, live = 0;
    for (int t = 0; t < n; ++t) {
      if (g_fib[t].state != DONE) ++live;
      if (g_fib[t].state == WAIT_BLOCK) ++at_block;
    }
    bool released = false;
    if (at_block > 0 && at_block == live) {
      for (int t = 0; t < n; ++t)
        if (g_fib[t].state == WAIT_BLOCK) g_fib[t].state = RUNNABLE;
      released = true;
    }
    for (int w = 0; w < nwarps; ++w) {
      int waiting = 0;
      unsigned mask = 0, livem = 0, existm = 0;
      for (int l = 0; l < 32 && w * 32 + l < n; ++l) {
        const Fiber &f = g_fib[w * 32 + l];
        existm |= 1u << l;
        if (f.state != DONE) livem |= 1u << l;
        if (f.state == WAIT_WARP) { ++waiting; mask |= f.wait_mask; }
      }
      if (waiting == 0) continue;
      // every live lane named by the mask must have arrived (lanes outside the mask may be elsewhere)
      bool all = true;
      for (int l = 0; l < 32 && w * 32 + l < n; ++l) {
        const Fiber &f = g_fib[w * 32 + l];
        if (((mask >> l) & 1u) && f.state != DONE && f.state != WAIT_WARP) all = false;
      }
      if (all) {
        if (mask & existm & ~livem) ++g_strict_violations;   // a *_sync mask names a lane that has already exited
        for (int l = 0; l < 32 && w * 32 + l < n; ++l)
          if (g_fib[w * 32 + l].state == WAIT_WARP) g_fib[w * 32 + l].state = RUNNABLE;
        released = true;
      }
    }
    if (!progressed && !released) {
      fprintf(stderr, "emu: deadlock in kernel %s, block (%u,%u,%u): %d of %d threads live, %d at __syncthreads\n", g_kernel,
              bid.x, bid.y, bid.z, live, n, at_block);
      abort();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// streams, events, graphs
// ---------------------------------------------------------------------------------------------------------------------
struct Graph { std::vector<std::function<void()>> nodes; };
struct GraphExec { std::vector<std::function<void()>> nodes; };
struct Stream { Graph *capture = nullptr; bool invalidated = false; };
struct Event { double t_ms = 0; };
static Stream g_default_stream;
static cudaError_t g_last_error = cudaSuccess;
static long long g_launches = 0;

static std::vector<Stream *> g_captures;   // streams of this process that are capturing
static Stream *S(cudaStream_t s) { return s ? s : &g_default_stream; }
bool capturing(cudaStream_t s) { return S(s)->capture != nullptr; }
static cudaError_t fail(cudaError_t e) { g_last_error = e; return e; }
// a call that is illegal while a capture is under way invalidates the capture, as the real runtime does
static cudaError_t illegal_in_capture(cudaStream_t) {
  for (Stream *c : g_captures) c->invalidated = true;
  return fail(cudaErrorStreamCaptureUnsupported);
}
static bool any_capture() { return !g_captures.empty(); }

static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void run_grid(const Cfg &c, const char *name, const std::function<void()> &body) {
  const size_t nthreads = (size_t)c.block.x * c.block.y * c.block.z;
  if (nthreads == 0 || nthreads > (size_t)MAX_THREADS || c.grid.x == 0 || c.grid.y == 0 || c.grid.z == 0) {
    fprintf(stderr, "emu: invalid launch configuration of %s: grid (%u,%u,%u) block (%u,%u,%u)\n", name, c.grid.x, c.grid.y,
            c.grid.z, c.block.x, c.block.y, c.block.z);
    g_last_error = cudaErrorInvalidValue;   // cudaErrorInvalidConfiguration on the device
    return;
  }
  if (c.smem > 227 * 1024) { g_last_error = cudaErrorInvalidValue; return; }
  ++g_launches;
  g_kernel = name;
  g_body = &body;
  g_nthreads = (int)nthreads;
  bdim = c.block;
  gdim = c.grid;
  {
    const char *gm = getenv("EMU_GUARD");
    if (gm && atoi(gm) != 0) {   // an over-run of the dynamic shared memory is an out-of-range shared address on the device
      if (!g_dyn_guarded) {
        char *m = (char *)mmap(nullptr, DYN_REGION + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) { g_last_error = cudaErrorMemoryAllocation; return; }
        mprotect(m + DYN_REGION, 4096, PROT_NONE);
        g_dyn_guarded = m;
      }
      const size_t payload = (c.smem + 15) & ~(size_t)15;
      g_dyn_ptr = g_dyn_guarded + DYN_REGION - payload;
      memset(g_dyn_ptr, 0xFF, payload);
    } else {
      g_dyn_smem.assign(c.smem + 16, (char)0xFF);
      g_dyn_ptr = g_dyn_smem.data();
    }
  }
  // EMU_ORDER=reverse|shuffle: blocks (and the threads inside a block) run in another order -- nothing in the product may
  // depend on the order in which the hardware happens to schedule them (a poor man's racecheck for order dependence)
  const size_t nblocks = (size_t)c.grid.x * c.grid.y * c.grid.z;
  const int order = exec_order();
  uint64_t lcg = 0x9E3779B97F4A7C15ull * (uint64_t)(g_launches + 1);
  std::vector<uint32_t> perm;
  if (order == 2) {
    perm.resize(nblocks);
    for (size_t i = 0; i < nblocks; ++i) perm[i] = (uint32_t)i;
    for (size_t i = nblocks; i > 1; --i) {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      std::swap(perm[i - 1], perm[(size_t)((lcg >> 33) % i)]);
    }
  }
  for (size_t k = 0; k < nblocks; ++k) {
    const size_t b = order == 1 ? nblocks - 1 - k : (order == 2 ? perm[k] : k);
    bid = uint3{(unsigned)(b % c.grid.x), (unsigned)((b / c.grid.x) % c.grid.y), (unsigned)(b / ((size_t)c.grid.x * c.grid.y))};
    run_block();
  }
  g_body = nullptr;
  g_kernel = "(none)";
}

void submit(const Cfg &c, const char *name, std::function<void()> body) {
  Stream *s = S(c.stream);
  if (s->capture) {
    Cfg cc = c;
    std::string nm(name);
    s->capture->nodes.push_back([cc, nm, body]() { run_grid(cc, nm.c_str(), body); });
  } else {
    run_grid(c, name, body);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// device memory
// ---------------------------------------------------------------------------------------------------------------------
struct Block { size_t bytes; void *base; size_t map_bytes; std::string shm; };
static std::map<void *, size_t> g_ipc_maps;   // peer blocks mapped by cudaIpcOpenMemHandle
static std::map<void *, Block> g_blocks;
static size_t g_live_bytes = 0;
static constexpr size_t RED = 64;
static int guard_mode() {
  static int m = -1;
  if (m < 0) { const char *e = getenv("EMU_GUARD"); m = e ? atoi(e) : 0; }
  return m;
}
static void segv_handler(int sig, siginfo_t *si, void *) {
  char buf[512];
  int n = snprintf(buf, sizeof(buf), "\nemu: signal %d at address %p in kernel %s, block (%u,%u,%u), thread %d -- out-of-bounds access?\n", sig,
                   si->si_addr, g_kernel, bid.x, bid.y, bid.z, g_cur);
  if (n > 0) (void)!write(2, buf, (size_t)n);
  _exit(139);
}
static void install_handler() {
  static bool done = false;
  if (done) return;
  done = true;
  static char altstack[65536];
  stack_t ss{};
  ss.ss_sp = altstack;
  ss.ss_size = sizeof(altstack);
  sigaltstack(&ss, nullptr);
  struct sigaction sa{};
  sa.sa_sigaction = segv_handler;
  sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
  sigaction(SIGBUS, &sa, nullptr);
}

}  // namespace emu

using namespace emu;

const char *cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotSupported: return "operation not supported (emulated runtime)";
    case cudaErrorStreamCaptureUnsupported: return "operation not permitted when stream is capturing";
    case cudaErrorStreamCaptureInvalidated: return "operation failed due to a previous error during capture";
    default: return "unknown error";
  }
}
cudaError_t cudaGetLastError() { cudaError_t e = g_last_error; g_last_error = cudaSuccess; return e; }
cudaError_t cudaPeekAtLastError() { return g_last_error; }
// every process emulates its own device; the index only has to be one a box could have (multi-rank runs: LOCAL_RANK)
static int g_device = 0;
cudaError_t cudaGetDeviceCount(int *n) { const char *e = getenv("EMU_DEVICES"); *n = e ? atoi(e) : 8; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) {
  int n = 0;
  cudaGetDeviceCount(&n);
  if (d < 0 || d >= n) return fail(cudaErrorInvalidValue);
  g_device = d;
  return cudaSuccess;
}
cudaError_t cudaGetDevice(int *d) { *d = g_device; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
  if (a == cudaDevAttrMultiProcessorCount) {
    const char *e = getenv("EMU_SMS");
    *v = e ? atoi(e) : 2;
    return cudaSuccess;
  }
  if (a == cudaDevAttrMaxSharedMemoryPerBlockOptin) { *v = 227 * 1024; return cudaSuccess; }
  return fail(cudaErrorInvalidValue);
}
cudaError_t cudaDeviceSynchronize() { return any_capture() ? illegal_in_capture(nullptr) : cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t *f, size_t *t) {
  *t = (size_t)192 << 30;
  *f = *t - g_live_bytes;
  return cudaSuccess;
}

cudaError_t cudaMalloc(void **p, size_t bytes) {
  install_handler();
  if (any_capture()) return illegal_in_capture(nullptr);
  if (bytes == 0) { *p = nullptr; return cudaSuccess; }
  Block b{bytes, nullptr, 0, std::string()};
  char *user = nullptr;
  const int gm = guard_mode();
  if (gm == 1 || gm == 2) {
    const size_t pg = 4096, payload = (bytes + 15) & ~(size_t)15;
    const size_t body = (payload + pg - 1) / pg * pg;
    b.map_bytes = body + 2 * pg;
    char *m = (char *)mmap(nullptr, b.map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return fail(cudaErrorMemoryAllocation);
    mprotect(m, pg, PROT_NONE);
    mprotect(m + pg + body, pg, PROT_NONE);
    b.base = m;
    user = gm == 1 ? m + pg + body - payload : m + pg;
    memset(m + pg, 0xFF, body);
  } else {
    char *m = (char *)malloc(bytes + 2 * RED + 256);
    if (!m) return fail(cudaErrorMemoryAllocation);
    b.base = m;
    user = (char *)(((uintptr_t)m + RED + 255) & ~(uintptr_t)255);
    memset(user - RED, 0xA5, RED);
    memset(user, 0xFF, bytes);
    memset(user + bytes, 0xA5, RED);
  }
  g_blocks[user] = b;
  g_live_bytes += bytes;
  *p = user;
  return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
  if (!p) return cudaSuccess;
  if (any_capture()) return illegal_in_capture(nullptr);
  auto it = g_blocks.find(p);
  if (it == g_blocks.end()) {
    fprintf(stderr, "emu: cudaFree of %p, which is not a live device allocation (double free?)\n", p);
    abort();
  }
  Block b = it->second;
  g_blocks.erase(it);
  g_live_bytes -= b.bytes;
  if (b.map_bytes) {
    munmap(b.base, b.map_bytes);
    if (!b.shm.empty()) shm_unlink(b.shm.c_str());
  } else {
    const unsigned char *u = (const unsigned char *)p;
    for (size_t i = 0; i < RED; ++i)
      if (u[-(ptrdiff_t)RED + (ptrdiff_t)i] != 0xA5 || u[b.bytes + i] != 0xA5) {
        fprintf(stderr, "emu: red zone of a %zu-byte device block overwritten (%s the block)\n", b.bytes,
                u[b.bytes + i] != 0xA5 ? "behind" : "in front of");
        abort();
      }
    memset(p, 0xEE, b.bytes);   // use-after-free reads garbage
    free(b.base);
  }
  return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 8); return *p ? cudaSuccess : fail(cudaErrorMemoryAllocation); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }

cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) {
  if (any_capture()) return illegal_in_capture(nullptr);   // legacy-stream synchronisation
  if (bytes) memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t s) {
  Stream *st = S(s);
  if (st->capture) {
    st->capture->nodes.push_back([dst, src, bytes]() { if (bytes) memmove(dst, src, bytes); });
    return cudaSuccess;
  }
  if (bytes) memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemset(void *dst, int v, size_t bytes) {
  if (any_capture()) return illegal_in_capture(nullptr);
  if (bytes) memset(dst, v, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t s) {
  Stream *st = S(s);
  if (st->capture) {
    st->capture->nodes.push_back([dst, v, bytes]() { if (bytes) memset(dst, v, bytes); });
    return cudaSuccess;
  }
  if (bytes) memset(dst, v, bytes);
  return cudaSuccess;
}

cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new Stream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new Stream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  if (s && s != &g_default_stream) { delete s->capture; delete s; }
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { return S(s)->capture ? illegal_in_capture(s) : cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode) {
  Stream *st = S(s);
  if (st->capture) return fail(cudaErrorInvalidValue);
  st->capture = new Graph();
  st->invalidated = false;
  g_captures.push_back(st);
  return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g) {
  Stream *st = S(s);
  if (!st->capture) { *g = nullptr; return fail(cudaErrorInvalidValue); }
  Graph *gr = st->capture;
  st->capture = nullptr;
  for (size_t i = 0; i < g_captures.size(); ++i)
    if (g_captures[i] == st) { g_captures.erase(g_captures.begin() + (long)i); break; }
  if (st->invalidated) {
    delete gr;
    *g = nullptr;
    st->invalidated = false;
    return fail(cudaErrorStreamCaptureInvalidated);
  }
  *g = gr;
  return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long) {
  if (!g) return fail(cudaErrorInvalidValue);
  *e = new GraphExec{g->nodes};
  return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s) {
  if (!e) return fail(cudaErrorInvalidValue);
  if (S(s)->capture) return illegal_in_capture(s);
  for (auto &n : e->nodes) n();
  return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new Event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new Event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  if (!e) return fail(cudaErrorInvalidValue);
  if (S(s)->capture) return cudaSuccess;   // an event node; timing events of a capture are not queried by the product
  e->t_ms = now_ms();
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  if (!a || !b) return fail(cudaErrorInvalidValue);
  *ms = (float)(b->t_ms - a->t_ms);
  return cudaSuccess;
}
// CUDA IPC: a page-aligned block of the guard modes is re-mapped in place onto a POSIX shared-memory object, which the
// peers (other host processes = other emulated GPUs) map in turn -- peer stores and volatile spins then work for real.
// Blocks of the malloc mode cannot be shared: not supported, and the product falls back to NCCL (that path is tested too).
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  auto it = g_blocks.find(p);
  if (it == g_blocks.end()) return fail(cudaErrorInvalidValue);
  Block &b = it->second;
  if (!b.map_bytes || ((uintptr_t)p & 4095)) return fail(cudaErrorNotSupported);
  const size_t len = (b.bytes + 4095) & ~(size_t)4095;
  if (b.shm.empty()) {
    static int counter = 0;
    char name[48];
    snprintf(name, sizeof(name), "/emuipc-%d-%d", (int)getpid(), counter++);
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)len) != 0) { if (fd >= 0) close(fd); return fail(cudaErrorNotSupported); }
    std::vector<char> keep((const char *)p, (const char *)p + b.bytes);
    void *m = mmap(p, len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
    close(fd);
    if (m != p) return fail(cudaErrorUnknown);
    memcpy(p, keep.data(), b.bytes);
    b.shm = name;
  }
  memset(h, 0, sizeof(*h));
  snprintf(h->reserved, 48, "%s", b.shm.c_str());
  memcpy(h->reserved + 48, &len, sizeof(len));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
  size_t len = 0;
  memcpy(&len, h.reserved + 48, sizeof(len));
  if (h.reserved[0] != '/' || len == 0) return fail(cudaErrorInvalidValue);
  int fd = shm_open(h.reserved, O_RDWR, 0600);
  if (fd < 0) return fail(cudaErrorInvalidValue);
  void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail(cudaErrorMemoryAllocation);
  g_ipc_maps[m] = len;
  *p = m;
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) {
  auto it = g_ipc_maps.find(p);
  if (it == g_ipc_maps.end()) return fail(cudaErrorInvalidValue);
  munmap(p, it->second);
  g_ipc_maps.erase(it);
  return cudaSuccess;
}

// ---- introspection for the tests ------------------------------------------------------------------------------------
extern "C" {
long long emu_kernel_launches(void) { return emu::g_launches; }
long long emu_live_device_bytes(void) { return (long long)emu::g_live_bytes; }
long long emu_live_device_blocks(void) { return (long long)emu::g_blocks.size(); }
long long emu_strict_violations(void) { return emu::g_strict_violations; }   // *_sync masks naming exited lanes
// verify the red zones of every live block (EMU_GUARD=0): 0 = intact
int emu_check_red_zones(void) {
  int bad = 0;
  for (auto &kv : emu::g_blocks) {
    if (kv.second.map_bytes) continue;
    const unsigned char *u = (const unsigned char *)kv.first;
    for (size_t i = 0; i < emu::RED; ++i)
      if (u[-(ptrdiff_t)emu::RED + (ptrdiff_t)i] != 0xA5 || u[kv.second.bytes + i] != 0xA5) { ++bad; break; }
  }
  return bad;
}
// tests/emu/fake_nccl.cpp: run fn(arg) in stream order -- now, or as a node of the graph the stream is capturing (1 = kept)
int emu_stream_enqueue(void *stream, void (*fn)(void *), void *arg) {
  emu::Stream *st = emu::S((cudaStream_t)stream);
  if (st->capture) {
    st->capture->nodes.push_back([fn, arg]() { fn(arg); });
    return 1;
  }
  fn(arg);
  return 0;
}
// self-test of the capture rules (tests/test_emu_cuda_source.py): 0 = the stand-in is as strict as the runtime
int emu_selftest_capture(void) {
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return 1;
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return 2;
  void *p = nullptr;
  if (cudaMalloc(&p, 64) != cudaErrorStreamCaptureUnsupported) return 3;      // an allocation while capturing ...
  cudaGraph_t g = nullptr;
  if (cudaStreamEndCapture(s, &g) != cudaErrorStreamCaptureInvalidated || g) return 4;   // ... invalidates the capture
  cudaGetLastError();
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return 5;
  if (cudaStreamSynchronize(s) != cudaErrorStreamCaptureUnsupported) return 6;
  if (cudaStreamEndCapture(s, &g) != cudaErrorStreamCaptureInvalidated) return 7;
  cudaGetLastError();
  // a clean capture records and replays, and executes nothing while recording
  int host = 0, *dev = nullptr;
  if (cudaMalloc((void **)&dev, sizeof(int)) != cudaSuccess) return 8;
  cudaMemset(dev, 0, sizeof(int));
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return 9;
  cudaMemsetAsync(dev, 1, sizeof(int), s);
  if (*dev != 0) return 10;
  if (cudaStreamEndCapture(s, &g) != cudaSuccess || !g) return 11;
  cudaGraphExec_t e = nullptr;
  if (cudaGraphInstantiate(&e, g, 0) != cudaSuccess) return 12;
  cudaGraphDestroy(g);
  if (cudaGraphLaunch(e, s) != cudaSuccess) return 13;
  cudaMemcpy(&host, dev, sizeof(int), cudaMemcpyDeviceToHost);
  if (host != 0x01010101) return 14;
  cudaGraphExecDestroy(e);
  cudaFree(dev);
  cudaStreamDestroy(s);
  return 0;
}
}
