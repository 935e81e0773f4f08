// TEST INFRASTRUCTURE -- never part of the product, never loaded by autopdex_b200 on its own.
//
// A functional stand-in for the CUDA runtime + the device-side built-ins, just large enough to compile the product's
// .cu translation units with g++ and EXECUTE the same kernel source on the host: every thread of a block runs as a
// fiber (own stack), __syncthreads / warp shuffles / votes are real rendezvous points between the fibers, blocks run one
// after the other, streams are in-order and synchronous, stream capture records closures and cudaGraphLaunch replays
// them (nothing executes during capture, calls that are illegal during capture return an error, as on the device).
//
// Purpose: the build container has no GPU.  tests/emu/build.py turns `k<<<g, b, sm, s>>>(args)` into
// `emu::launch(emu::cfg(g, b, sm, s), k)(args)` and links libapdx_b200_emu.so, which tests/test_emu_*.py drive through
// the unchanged C ABI and compare with the oracle.  That checks indexing, buffer sizes, launch logic, host/device
// control flow and numerics of the CUDA source; it says nothing about performance, memory-model races or anything the
// emulation serialises.  The -m gpu tests on a B200 remain the parity tests proper.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#define APDX_EMULATED 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __grid_constant__

// ---- vector types --------------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace emu {
extern uint3 tid, bid;
extern dim3 bdim, gdim;
void sync_block();                                   // __syncthreads
void sync_warp(unsigned mask);                       // __syncwarp and the rendezvous inside shuffles / votes
uint64_t *warp_slots();                              // 32 x 8-byte exchange slots of the calling fiber's warp
int lane_id();
void *dyn_smem();
long long clock();
}  // namespace emu

#define threadIdx (emu::tid)
#define blockIdx (emu::bid)
#define blockDim (emu::bdim)
#define gridDim (emu::gdim)
#define warpSize 32

// ---- device built-ins ----------------------------------------------------------------------------------------------
static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::sync_warp(mask); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }   // peers are other processes (shared mappings)
static inline long long clock64() { return emu::clock(); }

template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldca(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }

namespace emu {
template <typename T>
static inline T exchange(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
  uint64_t *s = warp_slots();
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  s[lane_id()] = raw;
  sync_warp(mask);
  raw = s[src_lane & 31];
  sync_warp(mask);
  T r;
  memcpy(&r, &raw, sizeof(T));
  return r;
}
static inline unsigned ballot(unsigned mask, int pred) {
  uint64_t *s = warp_slots();
  s[lane_id()] = pred ? 1u : 0u;
  sync_warp(mask);
  unsigned b = 0;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && s[l]) b |= 1u << l;
  sync_warp(mask);
  return b;
}
unsigned live_mask();   // lanes of the calling fiber's warp that have not exited
}  // namespace emu

template <typename T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  const int l = emu::lane_id();
  return emu::exchange(m, v, (l & ~(width - 1)) | (src & (width - 1)));
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  const int l = emu::lane_id();
  const int src = l ^ x;
  return emu::exchange(m, v, (src & ~(width - 1)) == (l & ~(width - 1)) ? src : l);
}
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  const int l = emu::lane_id();
  const int src = l + (int)d;
  return emu::exchange(m, v, (src & ~(width - 1)) == (l & ~(width - 1)) ? src : l);
}
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  const int l = emu::lane_id();
  const int src = l - (int)d;
  return emu::exchange(m, v, src >= (l & ~(width - 1)) ? src : l);
}
static inline unsigned __ballot_sync(unsigned m, int p) { return emu::ballot(m, p); }
static inline int __any_sync(unsigned m, int p) { return emu::ballot(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return emu::ballot(m, !p) == 0; }
static inline unsigned __activemask() { return emu::live_mask(); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

// blocks run one after the other and a block's fibers switch only at rendezvous points: plain read-modify-write
template <typename T, typename U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U, typename V> static inline T atomicCAS(T *p, U c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// CUDA's global min / max accept mixed arithmetic types
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)b < (C)a ? (C)b : (C)a; }
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)b : (C)a; }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

// ---- runtime API ---------------------------------------------------------------------------------------------------
enum cudaError_t {
  cudaSuccess = 0,
  cudaErrorInvalidValue = 1,
  cudaErrorMemoryAllocation = 2,
  cudaErrorNotSupported = 801,
  cudaErrorStreamCaptureUnsupported = 900,
  cudaErrorStreamCaptureInvalidated = 901,
  cudaErrorUnknown = 999
};
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaStreamNonBlocking = 1, cudaHostRegisterDefault = 0, cudaEventDisableTiming = 2, cudaEventDefault = 0 };

namespace emu {
struct Stream;
struct Event;
struct Graph;
struct GraphExec;
}  // namespace emu
typedef emu::Stream *cudaStream_t;
typedef emu::Event *cudaEvent_t;
typedef emu::Graph *cudaGraph_t;
typedef emu::GraphExec *cudaGraphExec_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned flags);
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode m);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long flags = 0);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- kernel launch -------------------------------------------------------------------------------------------------
namespace emu {
struct Cfg {
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
};
static inline Cfg cfg(dim3 g, dim3 b, size_t sm = 0, cudaStream_t s = nullptr) { return Cfg{g, b, sm, s}; }
void submit(const Cfg &c, const char *name, std::function<void()> thread_body);

template <typename... P>
struct Launcher {
  Cfg c;
  void (*k)(P...);
  const char *name;
  template <typename... A>
  void operator()(A &&...a) const {
    // kernel arguments are evaluated and copied NOW (launch time), like the parameter buffer of a real launch
    std::tuple<std::decay_t<P>...> args(std::forward<A>(a)...);
    void (*kk)(P...) = k;
    submit(c, name, [kk, args]() { std::apply(kk, args); });
  }
};
template <typename... P>
static inline Launcher<P...> launch(const Cfg &c, void (*k)(P...), const char *name = "kernel") {
  return Launcher<P...>{c, k, name};
}
}  // namespace emu
