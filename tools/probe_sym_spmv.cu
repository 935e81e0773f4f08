// Probe: can the L2 serve the mirrored half of a symmetric sliced-ELL matrix?
//
// Access pattern of a hypothetical symmetric SpMV on the 256^3 Q1 Poisson matrix: slice s (64 rows) streams its
// NU = 14 "upper" value rows (512 B each) from HBM and re-reads NL = 13 value rows that EARLIER slices streamed:
// 1 from itself, 3 from the previous mesh line (4 slices back), 9 from the previous mesh plane (1020 +- {0,4}
// slices back, i.e. ~7.3 MB earlier in the stream).  Compared with MODE 0 = the full-storage pattern (27 rows
// streamed).  Prints ms and effective GB/s; tells whether halving the stored values pays on B200.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_sym tools/probe_sym_spmv.cu && /tmp/probe_sym
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE, bool CS>
__global__ void __launch_bounds__(256, 2) k_probe(const double *__restrict__ val, const double *__restrict__ x,
                                                  double *__restrict__ y, int64_t n_slices, int n_cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int NU = MODE == 0 ? 27 : 14;
  const int lineoff[9] = {-1020 - 4, -1020, -1020 + 4, -4, 0, 4, 1020 - 4, 1020, 1020 + 4};
  for (int64_t s = warp0; s < n_slices; s += nwarps) {
    const double2 *vp = reinterpret_cast<const double2 *>(val + s * (int64_t)NU * 64) + lane;
    const int r0 = (int)(s * 64 + 2 * lane);
    double a0 = 0.0, a1 = 0.0;
    // streamed part
#pragma unroll
    for (int jb = 0; jb < NU; jb += (NU == 27 ? 9 : 7)) {
      constexpr int U = (NU == 27 ? 9 : 7);
      double2 v[U];
      double xa[U], xb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = CS ? __ldcs(vp + (jb + u) * 32) : __ldg(vp + (jb + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jb + u;                       // column of x: line offset + (-1,0,1)
        const int jj = MODE == 0 ? j : j + 13;      // upper half = entries 13..26 of the 27-point stencil
        const int off = lineoff[jj / 3] * 64 + (jj % 3 - 1);   // a mesh line = 4 slices = 256 rows (approx.)
        const int c0 = min(max(r0 + off, 0), n_cols - 1), c1 = min(max(r0 + 1 + off, 0), n_cols - 1);
        xa[u] = __ldg(x + c0);
        xb[u] = __ldg(x + c1);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { a0 += v[u].x * xa[u]; a1 += v[u].y * xb[u]; }
    }
    if (MODE == 1) {
      // mirrored part: entry jj in 0..12 is stored by the slice `dist` slices back, in its column 26 - jj - 13,
      // rows shifted by the in-line offset (unaligned -> two 8-byte loads per lane)
      double m0[13], m1[13], xa[13], xb[13];
#pragma unroll
      for (int jj = 0; jj < 13; ++jj) {
        const int dist = -lineoff[jj / 3];                 // slices back (0 for the same line)
        const int sh = -(jj % 3 - 1);                      // row shift inside the line
        int64_t sp = s - dist;
        if (sp < 0) sp = s;
        const double *base = val + sp * (int64_t)NU * 64 + (int64_t)(13 - jj) * 64;
        int k0 = 2 * lane + sh, k1 = k0 + 1;
        k0 = min(max(k0, 0), 63); k1 = min(max(k1, 0), 63);
        m0[jj] = __ldg(base + k0);
        m1[jj] = __ldg(base + k1);
        const int off = lineoff[jj / 3] * 64 + (jj % 3 - 1);
        xa[jj] = __ldg(x + min(max(r0 + off, 0), n_cols - 1));
        xb[jj] = __ldg(x + min(max(r0 + 1 + off, 0), n_cols - 1));
      }
#pragma unroll
      for (int jj = 0; jj < 13; ++jj) { a0 += m0[jj] * xa[jj]; a1 += m1[jj] * xb[jj]; }
    }
    y[r0] = a0;
    y[r0 + 1] = a1;
  }
}

template <int MODE, bool CS>
static int run(const char *name, const double *val, const double *x, double *y, int64_t n_slices, int n_cols, double gb_alg) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) k_probe<MODE, CS><<<148 * 32, 256>>>(val, x, y, n_slices, n_cols);
  CK(cudaEventRecord(e0));
  const int reps = 30;
  for (int i = 0; i < reps; ++i) k_probe<MODE, CS><<<148 * 32, 256>>>(val, x, y, n_slices, n_cols);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  printf("%-34s %8.4f ms   stored %.2f GB -> %7.1f GB/s of stored bytes; full-matrix equivalent %7.1f GB/s\n", name, ms,
         gb_alg, gb_alg / ms * 1e3, 27.0 * 64 * 8 * n_slices * 1e-9 / ms * 1e3);
  return 0;
}

int main() {
  const int64_t n_slices = 259084;   // 16 581 375 rows / 64
  const int n_cols = (int)(n_slices * 64);
  double *val, *x, *y;
  CK(cudaMalloc(&val, n_slices * 27 * 64 * sizeof(double)));
  CK(cudaMalloc(&x, (size_t)n_cols * sizeof(double)));
  CK(cudaMalloc(&y, (size_t)n_cols * sizeof(double)));
  CK(cudaMemset(val, 0, n_slices * 27 * 64 * sizeof(double)));
  CK(cudaMemset(x, 0, (size_t)n_cols * sizeof(double)));
  if (run<0, true>("full storage, evict-first stream", val, x, y, n_slices, n_cols, 27.0 * 512 * n_slices * 1e-9)) return 1;
  if (run<0, false>("full storage, default policy", val, x, y, n_slices, n_cols, 27.0 * 512 * n_slices * 1e-9)) return 1;
  if (run<1, false>("half storage + mirrored L2 reads", val, x, y, n_slices, n_cols, 14.0 * 512 * n_slices * 1e-9)) return 1;
  if (run<1, true>("half storage (stream evict-first)", val, x, y, n_slices, n_cols, 14.0 * 512 * n_slices * 1e-9)) return 1;
  return 0;
}
