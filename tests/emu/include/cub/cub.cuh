// TEST INFRASTRUCTURE: host stand-ins for the few cub device-wide primitives the product calls (see ../cuda_runtime.h).
// Same call contract: a first call with d_temp_storage == nullptr only reports the temporary size.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace emu { bool capturing(cudaStream_t s); }

namespace cub {

template <typename T>
struct CastOp {
  template <typename U> T operator()(const U &u) const { return (T)u; }
};
template <typename V, typename Op, typename It>
struct TransformInputIterator {
  It it;
  Op op;
  TransformInputIterator(It i, Op o) : it(i), op(o) {}
  V operator[](int64_t i) const { return op(it[i]); }
};

struct DeviceScan {
  template <typename In, typename Out, typename N>
  static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, In in, Out out, N n, cudaStream_t s = nullptr) {
    if (!tmp) { bytes = 256; return cudaSuccess; }
    if (emu::capturing(s)) return cudaErrorStreamCaptureUnsupported;
    using T = std::decay_t<decltype(out[0])>;
    T acc = 0;
    for (int64_t i = 0; i < (int64_t)n; ++i) { T v = (T)in[i]; out[i] = acc; acc += v; }   // in-place safe
    return cudaSuccess;
  }
  template <typename In, typename Out, typename N>
  static cudaError_t InclusiveSum(void *tmp, size_t &bytes, In in, Out out, N n, cudaStream_t s = nullptr) {
    if (!tmp) { bytes = 256; return cudaSuccess; }
    if (emu::capturing(s)) return cudaErrorStreamCaptureUnsupported;
    using T = std::decay_t<decltype(out[0])>;
    T acc = 0;
    for (int64_t i = 0; i < (int64_t)n; ++i) { acc += (T)in[i]; out[i] = acc; }
    return cudaSuccess;
  }
};

struct DeviceRadixSort {
  // stable LSD sort on the key bits [begin_bit, end_bit), like the device implementation
  template <typename K, typename V, typename N>
  static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, N n, int begin_bit = 0,
                               int end_bit = sizeof(K) * 8, cudaStream_t s = nullptr) {
    if (!tmp) { bytes = 256; return cudaSuccess; }
    if (emu::capturing(s)) return cudaErrorStreamCaptureUnsupported;
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    const int nb = end_bit - begin_bit;
    const K m = nb >= (int)sizeof(K) * 8 ? ~(K)0 : (K)((((K)1) << nb) - 1);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      return ((kin[a] >> begin_bit) & m) < ((kin[b] >> begin_bit) & m);
    });
    for (int64_t i = 0; i < (int64_t)n; ++i) { kout[i] = kin[order[(size_t)i]]; vout[i] = vin[order[(size_t)i]]; }
    return cudaSuccess;
  }
};

}  // namespace cub
