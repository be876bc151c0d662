// Context handle and the one collective of the hot path for hosts that are not PyTorch.
//
// The reference has no multi-device story (tf.keras layers on one CPU); SURVEY.md 8e derives the only exchange the
// B200 path needs: PLDA all-vs-all scoring shards the ENROLLED rows over the GPUs and all-gathers the transformed test
// x-vectors (layers/plda/plda.py:247-263 scores one set against itself; 25.6 MB at 50 k x 128).  The Python host uses
// torch.distributed for it (kaldi_tflite_b200/parallel.py); this file gives a C / Go / Java host the same step through
// the C-ABI: a context (device + stream + optional NCCL communicator) and ktf_nccl_allgather_xvec.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library must still load on a box without NCCL, and a host
// that never calls ktf_nccl_* never touches it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

struct NcclId { char bytes[KTF_NCCL_UNIQUE_ID_BYTES]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES == 128), passed by value
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclId, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllGatherFn)(const void*, void*, size_t, int /*ncclDataType_t*/, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclUint8 = 1;   // ncclUint8: the exchange is typeless, counts are bytes

struct NcclApi {
  void* lib = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllGatherFn all_gather = nullptr;
  GetErrorStringFn error_string = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* override_path = getenv("KTF_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (n == nullptr) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.get_unique_id = (GetUniqueIdFn)dlsym(api.lib, "ncclGetUniqueId");
      api.comm_init_rank = (CommInitRankFn)dlsym(api.lib, "ncclCommInitRank");
      api.comm_destroy = (CommDestroyFn)dlsym(api.lib, "ncclCommDestroy");
      api.all_gather = (AllGatherFn)dlsym(api.lib, "ncclAllGather");
      api.error_string = (GetErrorStringFn)dlsym(api.lib, "ncclGetErrorString");
      if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_gather) {
        dlclose(api.lib);
        api = NcclApi();
      }
    }
  }
  return api.lib ? &api : nullptr;
}

int nccl_fail(const NcclApi* api, const char* what, int code) {
  ktf::set_error("%s failed: %s (ncclResult %d)", what, (api && api->error_string) ? api->error_string(code) : "?", code);
  return KTF_ECUDA;
}

}  // namespace

struct ktf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  NcclComm comm = nullptr;
  int nranks = 1, rank = 0;
};

extern "C" {

int ktf_ctx_create(int32_t device, ktf_ctx** out) {
  KTF_CHECK_ARG(out != nullptr, "ktf_ctx_create: null argument");
  int count = 0;
  KTF_CUDA(cudaGetDeviceCount(&count));
  KTF_CHECK_ARG(device >= 0 && device < count, "device %d out of range (%d CUDA devices)", device, count);
  KTF_CUDA(cudaSetDevice(device));
  ktf_ctx* c = new ktf_ctx();
  c->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    ktf::set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    return KTF_ECUDA;
  }
  *out = c;
  return KTF_OK;
}

void ktf_ctx_destroy(ktf_ctx* c) {
  if (!c) return;
  if (c->comm) {
    NcclApi* api = nccl_api();
    if (api) api->comm_destroy(c->comm);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int32_t ktf_ctx_device(const ktf_ctx* c) { return c ? c->device : -1; }
void* ktf_ctx_stream(const ktf_ctx* c) { return c ? (void*)c->stream : nullptr; }

int ktf_ctx_synchronize(ktf_ctx* c) {
  KTF_CHECK_ARG(c != nullptr, "ktf_ctx_synchronize: null context");
  KTF_CUDA(cudaSetDevice(c->device));
  KTF_CUDA(cudaStreamSynchronize(c->stream));
  return KTF_OK;
}

int ktf_ctx_malloc(ktf_ctx* c, int64_t bytes, void** dev_out) {
  KTF_CHECK_ARG(c && dev_out && bytes >= 0, "ktf_ctx_malloc: bad argument");
  KTF_CUDA(cudaSetDevice(c->device));
  *dev_out = nullptr;
  if (bytes == 0) return KTF_OK;
  cudaError_t e = cudaMalloc(dev_out, (size_t)bytes);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ktf::set_error("cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return KTF_ENOMEM;
  }
  return KTF_OK;
}

int ktf_ctx_free(ktf_ctx* c, void* dev) {
  KTF_CHECK_ARG(c != nullptr, "ktf_ctx_free: null context");
  KTF_CUDA(cudaSetDevice(c->device));
  if (dev) KTF_CUDA(cudaFree(dev));
  return KTF_OK;
}

int ktf_ctx_memcpy_h2d(ktf_ctx* c, void* dst_dev, const void* src_host, int64_t bytes) {
  KTF_CHECK_ARG(c && (bytes == 0 || (dst_dev && src_host)) && bytes >= 0, "ktf_ctx_memcpy_h2d: bad argument");
  KTF_CUDA(cudaSetDevice(c->device));
  if (bytes) KTF_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
  return KTF_OK;
}

int ktf_ctx_memcpy_d2h(ktf_ctx* c, void* dst_host, const void* src_dev, int64_t bytes) {
  KTF_CHECK_ARG(c && (bytes == 0 || (dst_host && src_dev)) && bytes >= 0, "ktf_ctx_memcpy_d2h: bad argument");
  KTF_CUDA(cudaSetDevice(c->device));
  if (bytes) {
    KTF_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    KTF_CUDA(cudaStreamSynchronize(c->stream));     // the host buffer is valid when this returns
  }
  return KTF_OK;
}

int ktf_nccl_available(void) { return nccl_api() != nullptr ? 1 : 0; }

int ktf_nccl_unique_id(void* id_out_host) {
  KTF_CHECK_ARG(id_out_host != nullptr, "ktf_nccl_unique_id: null argument");
  NcclApi* api = nccl_api();
  KTF_CHECK_ARG(api != nullptr, "libnccl.so.2 could not be loaded (set KTF_NCCL_LIB to its path)");
  NcclId id;
  memset(&id, 0, sizeof(id));
  const int r = api->get_unique_id(&id);
  if (r != 0) return nccl_fail(api, "ncclGetUniqueId", r);
  memcpy(id_out_host, &id, sizeof(id));
  return KTF_OK;
}

int ktf_nccl_comm_init(ktf_ctx* c, int32_t nranks, int32_t rank, const void* id_host) {
  KTF_CHECK_ARG(c && id_host, "ktf_nccl_comm_init: null argument");
  KTF_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
  KTF_CHECK_ARG(c->comm == nullptr, "the context already has a communicator");
  NcclApi* api = nccl_api();
  KTF_CHECK_ARG(api != nullptr, "libnccl.so.2 could not be loaded (set KTF_NCCL_LIB to its path)");
  KTF_CUDA(cudaSetDevice(c->device));
  NcclId id;
  memcpy(&id, id_host, sizeof(id));
  const int r = api->comm_init_rank(&c->comm, nranks, id, rank);
  if (r != 0) {
    c->comm = nullptr;
    return nccl_fail(api, "ncclCommInitRank", r);
  }
  c->nranks = nranks;
  c->rank = rank;
  return KTF_OK;
}

int ktf_nccl_comm_destroy(ktf_ctx* c) {
  KTF_CHECK_ARG(c != nullptr, "ktf_nccl_comm_destroy: null context");
  if (c->comm) {
    NcclApi* api = nccl_api();
    if (api) {
      const int r = api->comm_destroy(c->comm);
      c->comm = nullptr;
      if (r != 0) return nccl_fail(api, "ncclCommDestroy", r);
    }
    c->comm = nullptr;
  }
  c->nranks = 1;
  c->rank = 0;
  return KTF_OK;
}

int ktf_nccl_allgather_xvec(ktf_ctx* c, const void* send_dev, void* recv_dev, int64_t rows_per_rank, int32_t dim,
                            int32_t elem_bytes, void* stream) {
  KTF_CHECK_ARG(c && send_dev && recv_dev, "ktf_nccl_allgather_xvec: null argument");
  KTF_CHECK_ARG(rows_per_rank >= 0 && dim > 0 && (elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8),
                "bad block shape");
  KTF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const size_t bytes = (size_t)rows_per_rank * dim * elem_bytes;
  if (bytes == 0) return KTF_OK;
  if (c->comm == nullptr) {
    // a context without a communicator is a world of one rank: the gather is a copy
    KTF_CHECK_ARG(c->nranks == 1, "no communicator");
    if (send_dev != recv_dev) KTF_CUDA(cudaMemcpyAsync(recv_dev, send_dev, bytes, cudaMemcpyDeviceToDevice, st));
    return KTF_OK;
  }
  NcclApi* api = nccl_api();
  const int r = api->all_gather(send_dev, recv_dev, bytes, kNcclUint8, c->comm, st);
  if (r != 0) return nccl_fail(api, "ncclAllGather", r);
  return KTF_OK;
}

}  // extern "C"
