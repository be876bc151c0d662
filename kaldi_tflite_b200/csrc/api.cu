// Error channel, version and launch accounting of libktf_b200.so.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace ktf {

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<uint64_t> g_dither_stream{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

uint64_t next_dither_stream() { return g_dither_stream.fetch_add(1, std::memory_order_relaxed); }

}  // namespace ktf

extern "C" {

const char* ktf_last_error(void) { return ktf::g_err; }

int ktf_version(void) { return 101; }

int ktf_set_dither_seed(uint64_t seed) {
  ktf::g_dither_stream.store(seed, std::memory_order_relaxed);
  return KTF_OK;
}

int ktf_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  KTF_CUDA(cudaGetDevice(&dev));
  KTF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  KTF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

int64_t ktf_launch_count(void) { return ktf::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
