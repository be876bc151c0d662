// Microbenchmark behind DESIGN.md section 4 (K5): how fast can 148 persistent CTAs write a row-major fp32 matrix in
// the tile pattern of the PLDA score GEMM (128 rows x 256 columns per CTA tile, 200 KB row pitch), by mechanism?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern_bench scripts/store_pattern_bench.cu -lcuda
//   ./store_pattern_bench [rows] [cols]
// Variants: 0 sequential fill (grid-stride float4)            1 tile walk, warp writes 512 B row pieces (st.global.v4)
//           2 tile walk, lane owns a row (32 rows x 16 B per instruction)
//           3 tile walk, TMA boxes 32 rows x 128 B (swizzle 128B)   4 tile walk, TMA boxes 16 rows x 1 KB (no swizzle)
//           5 tile walk, 1-D bulk copies of 1 KB rows          6, 7 as 3, 1 but tiles handed out n-fastest
//           8, 9, 10 as 3, 1, 4 but every CTA sweeps the columns of its own row blocks (contiguous m-major ranges)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int TM = 128, TN = 256, GROUP = 32;

__device__ __forceinline__ void tile_of(long long t, long long m_tiles, int n_tiles, int nfast, long long& mt, int& nt) {
  if (nfast) { mt = t / n_tiles; nt = (int)(t % n_tiles); return; }
  const long long per_group = (long long)GROUP * n_tiles;
  const long long g = t / per_group, r = t % per_group;
  const long long g0 = g * GROUP;
  const long long gsz = (m_tiles - g0) < GROUP ? (m_tiles - g0) : GROUP;
  nt = (int)(r / gsz);
  mt = g0 + r % gsz;
}

__global__ void fill_seq(float4* out, long long n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) out[i] = v;
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) tile_store(float* out, long long rows, long long cols, long long ld, int nfast,
                                                     const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m_tiles = rows / TM;
  const int n_tiles = (int)(cols / TN);
  const long long total = m_tiles * n_tiles;
  const int quarter = warp & 3, colq = warp >> 2;
  // nfast == 2: every CTA takes a contiguous range of the m-major order (it sweeps the columns of its own row blocks)
  const long long t0 = nfast == 2 ? blockIdx.x * total / gridDim.x : blockIdx.x;
  const long long t1 = nfast == 2 ? (blockIdx.x + 1) * total / gridDim.x : total;
  const long long ts = nfast == 2 ? 1 : gridDim.x;
  for (long long t = t0; t < t1; t += ts) {
    long long mt;
    int nt;
    tile_of(t, m_tiles, n_tiles, nfast, mt, nt);
    float* base = out + (mt * TM) * ld + (long long)nt * TN;
    if (VARIANT == 1) {
      const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll 4
      for (int r = 0; r < 32; ++r)
        *reinterpret_cast<float4*>(base + (long long)(quarter * 32 + r) * ld + colq * 128 + lane * 4) = v;
    } else if (VARIANT == 2) {
      const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll 4
      for (int c = 0; c < 32; ++c)
        *reinterpret_cast<float4*>(base + (long long)(quarter * 32 + lane) * ld + colq * 128 + c * 4) = v;
    } else if (VARIANT == 3) {
      unsigned char* box = smem + warp * 2 * 4096;
      for (int c = 0; c < 4; ++c) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
        __syncwarp();
        reinterpret_cast<float4*>(box + (c & 1) * 4096)[lane * 8 + (c & 7)] = make_float4(1.f, 2.f, 3.f, 4.f);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const int c0 = nt * TN + colq * 128 + c * 32, c1 = (int)(mt * TM) + quarter * 32;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(&tm),
                       "r"((unsigned)__cvta_generic_to_shared(box + (c & 1) * 4096)), "r"(c0), "r"(c1)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      }
    } else if (VARIANT == 4) {
      unsigned char* box = smem + warp * 16384;                  // 16 rows x 1 KB
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      __syncwarp();
      reinterpret_cast<float4*>(box)[lane] = make_float4(1.f, 2.f, 3.f, 4.f);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        const int c0 = nt * TN, c1 = (int)(mt * TM) + warp * 16;
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(&tm),
                     "r"((unsigned)__cvta_generic_to_shared(box)), "r"(c0), "r"(c1)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    } else if (VARIANT == 5) {
      unsigned char* rowbuf = smem + warp * 16384;               // 16 rows x 1 KB per warp
      if (lane < 16) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      __syncwarp();
      reinterpret_cast<float4*>(rowbuf)[lane] = make_float4(1.f, 2.f, 3.f, 4.f);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane < 16) {
        float* dst = base + (long long)(warp * 16 + lane) * ld;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst),
                     "r"((unsigned)__cvta_generic_to_shared(rowbuf + lane * 1024)), "r"(1024)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? atoll(argv[1]) : 20480, cols = argc > 2 ? atoll(argv[2]) : 49920, ld = 50000;
  float* out;
  CK(cudaMalloc(&out, rows * ld * 4));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fnp;
  CUtensorMap tm128, tm1k;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t box128[2] = {32, 32}, box1k[2] = {256, 16};
  if (enc(&tm128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&tm1k, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box1k, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 1;
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int smem = 8 * 16384;
  CK(cudaFuncSetAttribute(tile_store<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(tile_store<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(tile_store<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double bytes = (double)rows * cols * 4;
  const char* names[] = {"sequential fill", "tile, 512 B row pieces st.global", "tile, lane owns a row st.global",
                         "tile, TMA 32 x 128 B boxes", "tile, TMA 16 x 1 KB boxes", "tile, bulk 1 KB rows",
                         "n-fastest, TMA 32 x 128 B boxes", "n-fastest, 512 B row pieces st.global",
                         "own row block, TMA 32 x 128 B boxes", "own row block, 512 B row pieces st.global",
                         "own row block, TMA 16 x 1 KB boxes"};
  for (int v = 0; v < 11; ++v) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      switch (v) {
        case 0: fill_seq<<<sms * 8, 256>>>(reinterpret_cast<float4*>(out), rows * ld / 4); break;
        case 1: tile_store<1><<<sms, 256, 0>>>(out, rows, cols, ld, 0, tm128); break;
        case 2: tile_store<2><<<sms, 256, 0>>>(out, rows, cols, ld, 0, tm128); break;
        case 3: tile_store<3><<<sms, 256, smem>>>(out, rows, cols, ld, 0, tm128); break;
        case 4: tile_store<4><<<sms, 256, smem>>>(out, rows, cols, ld, 0, tm1k); break;
        case 5: tile_store<5><<<sms, 256, smem>>>(out, rows, cols, ld, 0, tm128); break;
        case 6: tile_store<3><<<sms, 256, smem>>>(out, rows, cols, ld, 1, tm128); break;
        case 7: tile_store<1><<<sms, 256, 0>>>(out, rows, cols, ld, 1, tm128); break;
        case 8: tile_store<3><<<sms, 256, smem>>>(out, rows, cols, ld, 2, tm128); break;
        case 9: tile_store<1><<<sms, 256, 0>>>(out, rows, cols, ld, 2, tm128); break;
        case 10: tile_store<4><<<sms, 256, smem>>>(out, rows, cols, ld, 2, tm1k); break;
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const double b = v == 0 ? (double)rows * ld * 4 : bytes;
    printf("%-40s %8.3f ms  %7.1f GB/s\n", names[v], best, b / best * 1e-6);
  }
  return 0;
}
