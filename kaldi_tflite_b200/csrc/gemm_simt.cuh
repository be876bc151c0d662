// Tiled SIMT GEMM (C = A * B^T, both operands K-contiguous) used for the exact-arithmetic
// paths: fp32 TDNN reference precision, PLDA transform / scoring in fp32 or fp64.
// 64x64 output tile, 16-wide K slab, 256 threads, 4x4 outputs per thread.
#pragma once

#include <cuda_runtime.h>

namespace ktf {

constexpr int kTileM = 64, kTileN = 64, kTileK = 16, kGemmThreads = 256;

// ALoad:  T operator()(long long row, int k) const      -- row < M, k < K guaranteed
// BLoad:  T operator()(long long col, int k) const
// Epi:    void operator()(long long row, long long col, T acc) const
template <typename T, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(kGemmThreads)
gemm_nt_kernel(long long M, long long N, int K, ALoad aload, BLoad bload, Epi epi) {
  __shared__ T As[kTileK][kTileM + 4];
  __shared__ T Bs[kTileK][kTileN + 4];
  const long long m0 = (long long)blockIdx.x * kTileM;   // M tiles on x (2^31 limit)
  const long long n0 = (long long)blockIdx.y * kTileN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  for (int k0 = 0; k0 < K; k0 += kTileK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread; k fastest for coalescing
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * kGemmThreads;
      const int kk = idx & 15, r = idx >> 4;
      const int k = k0 + kk;
      const long long row = m0 + r, col = n0 + r;
      As[kk][r] = (row < M && k < K) ? aload(row, k) : T(0);
      Bs[kk][r] = (col < N && k < K) ? bload(col, k) : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTileK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long col = n0 + tx * 4 + j;
      if (col < N) epi(row, col, acc[i][j]);
    }
  }
}

template <typename T>
struct DenseLoad {
  const T* p;
  long long ld;
  __device__ __forceinline__ T operator()(long long r, int k) const { return p[r * ld + k]; }
};

}  // namespace ktf
