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
uint64_t next_dither_stream();   // seed + number of dithered forward calls so far (ktf_set_dither_seed)

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

// SM count of the CURRENT device (cached per device: a process may drive several GPUs).
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// True the first time it is called for (tag, current device): function attributes such as the dynamic shared-memory
// limit are per device, so "set once" guards must be keyed on the device as well.
inline bool first_use_on_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

// Stream-ordered scratch that is released on every exit path of the call that took it (early error returns included).
struct Scratch {
  cudaStream_t st;
  void* ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
  int n = 0;
  explicit Scratch(cudaStream_t s) : st(s) {}
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  ~Scratch() {
    for (int i = 0; i < n; ++i)
      if (ptrs[i]) cudaFreeAsync(ptrs[i], st);
  }
  template <typename T>
  cudaError_t take(T** p, size_t bytes);
};

// Stream-ordered scratch allocation.  The device's default memory pool is told to keep freed
// blocks (release threshold = max) on first use: with the driver default (0) every stream / event
// synchronisation hands the pool's memory back to the OS and the next call pays to map it again.
inline cudaError_t malloc_async(void** p, size_t bytes, cudaStream_t st) {
  static int pooled_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (pooled_dev != dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pooled_dev = dev;
  }
  return cudaMallocAsync(p, bytes, st);
}

template <typename T>
inline cudaError_t Scratch::take(T** p, size_t bytes) {
  void* q = nullptr;
  const cudaError_t e = malloc_async(&q, bytes, st);
  if (e == cudaSuccess && n < 4) ptrs[n++] = q;
  *p = static_cast<T*>(q);
  return e;
}

// Grow-only device workspace owned by a handle (handles own packed weights, tables and a
// workspace).  Growth is a cudaFree + cudaMalloc (device-synchronising); steady state is free.
struct Workspace {
  void* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return KTF_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      e = cudaMalloc(&ptr, bytes);
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMalloc(%zu) for a workspace failed: %s", bytes, cudaGetErrorString(e));
        return KTF_ENOMEM;
      }
      cap = bytes;
      return KTF_OK;
    }
    cap = want;
    return KTF_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// Carves 256-byte aligned regions out of a workspace (sizes first, pointers after ensure()).
struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  }
};

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
