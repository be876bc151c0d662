// Shared host-side helpers for libktf_b200.so (error channel, launch accounting).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ktf_b200.h"

namespace ktf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define KTF_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ktf::set_error(__VA_ARGS__);          \
      return KTF_EINVAL;                    \
    }                                       \
  } while (0)

#define KTF_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ktf::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                     __FILE__, __LINE__);                                           \
      return KTF_ECUDA;                                                             \
    }                                                                               \
  } while (0)

// Checks the launch that was just enqueued.
#define KTF_LAUNCH_OK()                                                             \
  do {                                                                              \
    ktf::count_launch();                                                            \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ktf::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                           \
      return KTF_ECUDA;                                                             \
    }                                                                               \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
inline int upload(T** dst, const T* src, size_t n) {
  cudaError_t e = cudaMalloc((void**)dst, n * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return KTF_ENOMEM;
  }
  e = cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("cudaMemcpy H2D failed: %s", cudaGetErrorString(e));
    return KTF_ECUDA;
  }
  return KTF_OK;
}

}  // namespace ktf
