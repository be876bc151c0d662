// tcgen05 / TMEM implicit-GEMM engine for the TDNN stack (sm_100a only).
//
// Replaces the tf.gather + tf.nn.conv2d formulation of layers/tdnn/tdnn.py:251-280 and the keras
// ReLU / BatchNormalization passes that follow it (models/kaldi/sequential.py:71-76,
// layers/normalization/batchnorm.py:81-88) with ONE warp-specialised kernel per layer:
//
//   y[r, u] = scale[u] * relu( sum_k sum_d x[r + ctx_k, d] * W[u, k*D + d] + bias[u] ) + offset[u]
//
//   * operands bf16, accumulation fp32 in TMEM (tcgen05.mma.cta_group::1.kind::f16, M128 x N256 x K16);
//   * the frame splice is IMPLICIT: tap k is a TMA box load of the activation matrix shifted by ctx_k
//     rows -- the (B, T, K, D) gathered tensor of the reference is never built;
//   * edge clamping (tdnn.py:244-247) is provided by the activation layout: every utterance carries
//     kHalo replicated rows on both sides ("padded rows"); the epilogue of each layer writes the halo
//     replicas the next layer's taps need and skips the halo rows of its own tile;
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//     warps 2..5 = epilogue (TMEM -> registers -> bias/ReLU/BN -> bf16/fp32 -> global);
//   * 4-stage smem ring (48 KB per stage) between TMA and MMA, 2 accumulator stages of 256 TMEM columns
//     between MMA and epilogue, persistent CTAs (one per SM) walking output tiles n-fastest so that the
//     n-tiles of one row block run concurrently and share the activation rows through L2.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tdnn_internal.cuh"

namespace {

constexpr int kHalo = 4;          // replicated rows on each side of every utterance (>= max |context|)
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kThreadsTc = 192;   // 6 warps
constexpr int kABytes = BM * BK * 2;   // 16 KB
constexpr int kBBytes = BN * BK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemTc = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                        kAccStages * 3 * BN * 4 /*epilogue vectors*/;

// rowmap flags
constexpr int kRowStore = 1, kRowFirst = 2, kRowLast = 4;

struct TcArgs {
  int num_taps;
  int ctx[KTF_MAX_CONTEXT];
  int kblocks_per_tap;      // ceil(D / 64)
  int tap_cols;             // column distance between taps in the weight matrix (= D)
  long long m_rows;         // rows of the A / output matrices (padded rows)
  int n_cols;               // U
  const int* rowmap;        // per output row flags, or nullptr = store every row < m_rows
  const float* bias;
  const float* scale;
  const float* offset;
  int relu;
  void* out;                // bf16 or fp32, row-major
  long long out_ld;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);
  d |= (unsigned long long)(1024 >> 4) << 32;   // stride byte offset
  d |= (unsigned long long)1 << 46;             // descriptor version (sm_100)
  d |= (unsigned long long)2 << 61;             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool OUT_BF16>
__device__ __forceinline__ void store_row32(void* out, long long ld, long long row, int col0, int n_cols,
                                            const float (&v)[32]) {
  if (OUT_BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(out) + row * ld + col0;
    if (col0 + 32 <= n_cols && ((reinterpret_cast<unsigned long long>(p) & 15ull) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
        pk.x = *reinterpret_cast<unsigned*>(&h0);
        pk.y = *reinterpret_cast<unsigned*>(&h1);
        pk.z = *reinterpret_cast<unsigned*>(&h2);
        pk.w = *reinterpret_cast<unsigned*>(&h3);
        reinterpret_cast<uint4*>(p)[q] = pk;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_cols) p[i] = __float2bfloat16_rn(v[i]);
    }
  } else {
    float* p = reinterpret_cast<float*>(out) + row * ld + col0;
    if (col0 + 32 <= n_cols && ((reinterpret_cast<unsigned long long>(p) & 15ull) == 0)) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_cols) p[i] = v[i];
    }
  }
}

template <bool OUT_BF16>
__global__ void __launch_bounds__(kThreadsTc, 1)
tdnn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<unsigned long long>(smem_raw) + 1023ull) &
                                                         ~1023ull);
  unsigned char* sA = smem;
  unsigned char* sB = smem + kStages * kABytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kStages * kStageBytes);
  unsigned long long* full_bar = bars;                       // [kStages]
  unsigned long long* empty_bar = bars + kStages;            // [kStages]
  unsigned long long* tfull_bar = bars + 2 * kStages;        // [kAccStages]
  unsigned long long* tempty_bar = tfull_bar + kAccStages;   // [kAccStages]
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty_bar + kAccStages);
  float* s_vec = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);  // [kAccStages][3][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m_tiles = (a.m_rows + BM - 1) / BM;
  const int n_tiles = (a.n_cols + BN - 1) / BN;
  const long long total_tiles = m_tiles * n_tiles;
  const int num_kb = a.num_taps * a.kblocks_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(kAccStages * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
      int stage = 0;
      unsigned phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long mt = tile / n_tiles;
        const int nt = (int)(tile - mt * n_tiles);
        const int row0 = (int)(mt * BM), col0 = nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / a.kblocks_per_tap;
          const int d0 = (kb - tap * a.kblocks_per_tap) * BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sA + stage * kABytes, &tmA, &full_bar[stage], d0, row0 + a.ctx[tap]);
          tma_load_2d(sB + stage * kBBytes, &tmB, &full_bar[stage], tap * a.tap_cols + d0, col0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(BN >> 3) << 17) |
                             ((unsigned)(BM >> 4) << 24);
      int stage = 0;
      unsigned phase = 0;
      int it = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const unsigned acc_phase = (unsigned)(it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const unsigned tmem_d = tmem_base + (unsigned)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const unsigned long long adesc = umma_desc(smem_u32(sA + stage * kABytes));
          const unsigned long long bdesc = umma_desc(smem_u32(sB + stage * kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the swizzled row: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + (unsigned long long)(2 * k), bdesc + (unsigned long long)(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);     // frees the smem slot once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);         // accumulator ready for the epilogue
      }
    }
  } else {
    // ================= epilogue warps (2..5) =================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 64;          // 0..127
    int it = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const long long mt = tile / n_tiles;
      const int nt = (int)(tile - mt * n_tiles);
      const int col_base = nt * BN;
      const int acc = it & 1;
      const unsigned acc_phase = (unsigned)(it >> 1) & 1u;
      float* vb = s_vec + acc * 3 * BN;
      // per-column epilogue vectors for this n-tile (double buffered with the accumulator stage)
      for (int c = et; c < BN; c += 128) {
        const int col = col_base + c;
        const bool ok = col < a.n_cols;
        vb[c] = (ok && a.bias) ? a.bias[col] : 0.0f;
        vb[BN + c] = (ok && a.scale) ? a.scale[col] : 1.0f;
        vb[2 * BN + c] = (ok && a.offset) ? a.offset[col] : 0.0f;
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();

      const long long row = mt * BM + quarter * 32 + lane;
      int flags = 0;
      if (row < a.m_rows) flags = a.rowmap ? a.rowmap[row] : kRowStore;
      const unsigned taddr0 = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(acc * BN);
      for (int cc = 0; cc < BN; cc += 32) {
        if (col_base + cc >= a.n_cols) break;           // warp-uniform
        float v[32];
        tmem_ld32(taddr0 + (unsigned)cc, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = v[i] + vb[cc + i];
          if (a.relu) t = fmaxf(t, 0.0f);
          v[i] = fmaf(t, vb[BN + cc + i], vb[2 * BN + cc + i]);
        }
        if (flags & kRowStore) {
          store_row32<OUT_BF16>(a.out, a.out_ld, row, col_base + cc, a.n_cols, v);
          if (flags & kRowFirst)
            for (int h = 1; h <= kHalo; ++h)
              store_row32<OUT_BF16>(a.out, a.out_ld, row - h, col_base + cc, a.n_cols, v);
          if (flags & kRowLast)
            for (int h = 1; h <= kHalo; ++h)
              store_row32<OUT_BF16>(a.out, a.out_ld, row + h, col_base + cc, a.n_cols, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kAccStages * BN));
  }
}

// ---------------------------------------------------------------------------------------------------
// Layout kernels around the GEMM
// ---------------------------------------------------------------------------------------------------

// Padded-row bookkeeping of a ragged batch: utterance b owns padded rows
// [poffs[b], poffs[b+1]) = kHalo + T_b + kHalo rows.  rowmap flags per padded row.
__global__ void build_padded_kernel(const long long* __restrict__ offs, long long batch,
                                    long long* __restrict__ poffs, int* __restrict__ rowmap) {
  // one CTA per utterance
  const long long b = blockIdx.x;
  const long long r0 = offs[b], T = offs[b + 1] - r0;
  const long long p0 = r0 + 2LL * kHalo * b;
  if (threadIdx.x == 0) {
    poffs[b] = p0;
    if (b == batch - 1) poffs[batch] = p0 + T + 2 * kHalo;
  }
  for (long long i = threadIdx.x; i < T + 2 * kHalo; i += blockDim.x) {
    int f = 0;
    const long long t = i - kHalo;
    if (t >= 0 && t < T) {
      f = kRowStore;
      if (t == 0) f |= kRowFirst;
      if (t == T - 1) f |= kRowLast;
    }
    rowmap[p0 + i] = f;
  }
}

// Materialised splice ("im2col") for layers whose feature dimension is not a multiple of 64:
// out[p, k*D + d] = x[clamp(t + ctx_k)] for every padded row p (halo rows clamp to the edge frames),
// bf16, row stride ld (zero padded).  x is fp32 (rows, D) or bf16 padded-row (prow, D).
template <typename TIn>
__global__ void splice_kernel(const TIn* __restrict__ x, int D, long long x_ld, int x_is_padded,
                              const long long* __restrict__ offs, const long long* __restrict__ poffs,
                              long long batch, int num_taps, const int* __restrict__ ctx_dev,
                              __nv_bfloat16* __restrict__ out, long long ld) {
  // grid: (padded rows), block: 128 threads over the K*D columns
  __shared__ int s_ctx[KTF_MAX_CONTEXT];
  if (threadIdx.x < num_taps) s_ctx[threadIdx.x] = ctx_dev[threadIdx.x];
  __syncthreads();
  const long long p = blockIdx.x;
  long long lo = 0, hi = batch;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (poffs[mid] <= p) lo = mid; else hi = mid;
  }
  const long long T = offs[lo + 1] - offs[lo];
  long long t = p - poffs[lo] - kHalo;
  t = min(max(t, 0LL), T - 1);
  const long long base = x_is_padded ? (poffs[lo] + kHalo) : offs[lo];
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float v = 0.0f;
    if (c < num_taps * D) {
      const int k = c / D, d = c - k * D;
      const long long tt = min(max(t + s_ctx[k], 0LL), T - 1);
      v = (float)x[(base + tt) * x_ld + d];
    }
    out[p * ld + c] = __float2bfloat16_rn(v);
  }
}

// Per-utterance sum / sum of squares over the real (non-halo) rows of a padded bf16 activation matrix,
// then mean || std (stats_pooling.py:228-240) written as bf16 (next GEMM operand) and/or fp32.
__global__ void stats_padded_kernel(const __nv_bfloat16* __restrict__ y, long long ld, int dim,
                                    const long long* __restrict__ poffs, int include_std, float eps,
                                    __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                    long long out_ld) {
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  const long long b = blockIdx.x;
  const long long r0 = poffs[b] + kHalo, r1 = poffs[b + 1] - kHalo;
  float s = 0.0f, s2 = 0.0f, cs = 0.0f, cs2 = 0.0f;          // Kahan-compensated fp32 sums
  for (long long r = r0; r < r1; ++r) {
    const float v = __bfloat162float(y[r * ld + d]);
    float yk = v - cs, tk = s + yk;
    cs = (tk - s) - yk;
    s = tk;
    yk = v * v - cs2;
    tk = s2 + yk;
    cs2 = (tk - s2) - yk;
    s2 = tk;
  }
  const float n = (float)(r1 - r0);
  const float mean = s / n;
  const float var = s2 / n - __fmul_rn(mean, mean);
  const float sd = sqrtf(fmaxf(var, 0.0f) + eps);
  if (out_bf16) {
    out_bf16[b * out_ld + d] = __float2bfloat16_rn(mean);
    if (include_std) out_bf16[b * out_ld + dim + d] = __float2bfloat16_rn(sd);
  }
  if (out_f32) {
    out_f32[b * out_ld + d] = mean;
    if (include_std) out_f32[b * out_ld + dim + d] = sd;
  }
}

__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ x, long long rows, int dim, long long ld,
                                        __nv_bfloat16* __restrict__ out) {
  const long long total = rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld;
    const int c = (int)(i - r * ld);
    out[i] = __float2bfloat16_rn(c < dim ? x[r * dim + c] : 0.0f);
  }
}

// Gathers the real rows of a padded fp32 matrix back into the caller's ragged (rows, dim) layout.
__global__ void unpad_rows_kernel(const float* __restrict__ yp, long long ld, int dim,
                                  const long long* __restrict__ offs, const long long* __restrict__ poffs,
                                  long long batch, float* __restrict__ y) {
  const long long b = blockIdx.x;
  const long long r0 = offs[b], T = offs[b + 1] - r0, p0 = poffs[b] + kHalo;
  for (long long i = threadIdx.x; i < T * dim; i += blockDim.x) {
    const long long t = i / dim;
    const int d = (int)(i - t * dim);
    y[(r0 + t) * dim + d] = yp[(p0 + t) * ld + d];
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------

struct TcLayer {
  int D = 0, U = 0, K = 0;
  int ctx[KTF_MAX_CONTEXT] = {0};
  bool implicit = false;            // taps via shifted TMA loads (D % 64 == 0, SAME, no subsampling)
  long long w_ld = 0;               // bf16 weight row stride (elements)
  __nv_bfloat16* d_w = nullptr;     // (U, w_ld)
  int* d_ctx = nullptr;
  CUtensorMap tmB;
};

// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so that the
// library has no link-time dependency on libcuda (it must still load on a box without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_map(CUtensorMap* map, const void* base, unsigned long long inner, unsigned long long rows,
               unsigned long long ld_elems, unsigned box_inner, unsigned box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) {
    ktf::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return KTF_ECUDA;
  }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ktf::set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner %llu rows %llu ld %llu)", (int)r, inner,
                   rows, ld_elems);
    return KTF_ECUDA;
  }
  return KTF_OK;
}

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

int check_arch() {
  static int arch = 0;
  if (arch == 0) arch = ktf_device_arch();
  if (arch < 100) {
    ktf::set_error("the tcgen05 TDNN engine needs an sm_100 device (found sm_%d)", arch);
    return KTF_EINVAL;
  }
  return KTF_OK;
}

int prepare_layer(TcLayer* L, const ktf_affine_cfg& c, const float* w_host) {
  L->D = c.in_dim;
  L->U = c.out_dim;
  L->K = c.num_context;
  for (int k = 0; k < c.num_context; ++k) L->ctx[k] = c.context[k];
  int maxabs = 0;
  for (int k = 0; k < c.num_context; ++k) maxabs = std::max(maxabs, std::abs(c.context[k]));
  L->implicit = (c.in_dim % 64 == 0) && !c.padding_valid && c.subsampling_factor == 1 && maxabs <= kHalo;
  const long long cols = (long long)c.num_context * c.in_dim;
  L->w_ld = round_up(cols, 8);
  std::vector<__nv_bfloat16> wb((size_t)c.out_dim * L->w_ld, __float2bfloat16(0.0f));
  for (int u = 0; u < c.out_dim; ++u)
    for (long long j = 0; j < cols; ++j) wb[(size_t)u * L->w_ld + j] = __float2bfloat16(w_host[(size_t)u * cols + j]);
  int rc;
  if ((rc = ktf::upload(&L->d_w, wb.data(), wb.size())) != KTF_OK) return rc;
  if ((rc = ktf::upload(&L->d_ctx, L->ctx, (size_t)KTF_MAX_CONTEXT)) != KTF_OK) return rc;
  return encode_map(&L->tmB, L->d_w, (unsigned long long)cols, (unsigned long long)c.out_dim,
                    (unsigned long long)L->w_ld, BK, BN);
}

void release_layer(TcLayer* L) {
  if (!L) return;
  cudaFree(L->d_w);
  cudaFree(L->d_ctx);
}

template <bool OUT_BF16>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& args, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    KTF_CUDA(cudaFuncSetAttribute(tdnn_tc_kernel<OUT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc));
    attr_done = true;
  }
  const long long tiles = ((args.m_rows + BM - 1) / BM) * ((args.n_cols + BN - 1) / BN);
  if (tiles <= 0) return KTF_OK;
  const unsigned grid = (unsigned)std::min<long long>(tiles, ktf::num_sms());
  tdnn_tc_kernel<OUT_BF16><<<grid, kThreadsTc, kSmemTc, st>>>(tmA, tmB, args);
  KTF_LAUNCH_OK();
  return KTF_OK;
}

// One affine layer on padded rows.  in: bf16 (prow, in_ld) padded-row activations, or (for !implicit)
// any source handled by the caller through `spliced`.
int run_layer(const TcLayer& L, const ktf_affine* a, const __nv_bfloat16* A, long long a_ld, long long a_cols,
              long long m_rows, const int* rowmap, void* out, long long out_ld, bool out_bf16, bool a_is_spliced,
              cudaStream_t st) {
  CUtensorMap tmA;
  int rc = encode_map(&tmA, A, (unsigned long long)a_cols, (unsigned long long)m_rows, (unsigned long long)a_ld, BK, BM);
  if (rc != KTF_OK) return rc;
  TcArgs args{};
  if (a_is_spliced) {
    args.num_taps = 1;
    args.ctx[0] = 0;
    args.kblocks_per_tap = (int)((a_cols + BK - 1) / BK);
    args.tap_cols = 0;
  } else {
    args.num_taps = L.K;
    for (int k = 0; k < L.K; ++k) args.ctx[k] = L.ctx[k];
    args.kblocks_per_tap = L.D / BK;
    args.tap_cols = L.D;
  }
  args.m_rows = m_rows;
  args.n_cols = L.U;
  args.rowmap = rowmap;
  args.bias = a->d_bias;
  args.scale = a->d_scale;
  args.offset = a->d_offset;
  args.relu = a->cfg.activation == KTF_ACT_RELU;
  args.out = out;
  args.out_ld = out_ld;
  return out_bf16 ? launch_gemm<true>(tmA, L.tmB, args, st) : launch_gemm<false>(tmA, L.tmB, args, st);
}

}  // namespace

namespace ktf {

int affine_tc_prepare(ktf_affine* a, const float* weights_host) {
  int rc = check_arch();
  if (rc != KTF_OK) return rc;
  TcLayer* L = new TcLayer();
  rc = prepare_layer(L, a->cfg, weights_host);
  if (rc != KTF_OK) {
    release_layer(L);
    delete L;
    return rc;
  }
  a->tc = L;
  return KTF_OK;
}

void affine_tc_release(ktf_affine* a) {
  if (a && a->tc) {
    release_layer(static_cast<TcLayer*>(a->tc));
    delete static_cast<TcLayer*>(a->tc);
    a->tc = nullptr;
  }
}

// Stand-alone layer call (fp32 in / fp32 out): splice to bf16 once, then the tensor-core GEMM.
// Only SAME padding without subsampling is offered on this engine (the x-vector networks use nothing else).
int affine_tc_forward(const ktf_affine* a, const float* x_dev, const int64_t* in_offsets_dev,
                      const int64_t* out_offsets_dev, int64_t batch, int64_t total_in_rows,
                      int64_t total_out_rows, float* y_dev, float* stats_dev, cudaStream_t st) {
  const TcLayer& L = *static_cast<const TcLayer*>(a->tc);
  KTF_CHECK_ARG(!a->cfg.padding_valid && a->cfg.subsampling_factor == 1,
                "KTF_PREC_BF16 supports padding=SAME, subsampling_factor=1 (use KTF_PREC_F32 otherwise)");
  KTF_CHECK_ARG(total_in_rows == total_out_rows, "row count mismatch");
  (void)out_offsets_dev;
  const long long prow = total_in_rows + 2LL * kHalo * batch;
  const long long cols = (long long)L.K * L.D, ld = round_up(cols, 8);
  long long* poffs = nullptr;
  int* rowmap = nullptr;
  __nv_bfloat16* spliced = nullptr;
  float* yp = nullptr;
  KTF_CUDA(cudaMallocAsync((void**)&poffs, (batch + 1) * sizeof(long long), st));
  KTF_CUDA(cudaMallocAsync((void**)&rowmap, prow * sizeof(int), st));
  KTF_CUDA(cudaMallocAsync((void**)&spliced, (size_t)prow * ld * sizeof(__nv_bfloat16), st));
  KTF_CUDA(cudaMallocAsync((void**)&yp, (size_t)prow * L.U * sizeof(float), st));
  build_padded_kernel<<<(unsigned)batch, 128, 0, st>>>((const long long*)in_offsets_dev, batch, poffs, rowmap);
  KTF_LAUNCH_OK();
  splice_kernel<float><<<(unsigned)prow, 128, 0, st>>>(x_dev, L.D, L.D, 0, (const long long*)in_offsets_dev, poffs,
                                                       batch, L.K, L.d_ctx, spliced, ld);
  KTF_LAUNCH_OK();
  int rc = run_layer(L, a, spliced, ld, cols, prow, rowmap, yp, L.U, /*out_bf16=*/false, /*spliced=*/true, st);
  if (rc != KTF_OK) return rc;
  float* y = y_dev;
  if (y == nullptr) KTF_CUDA(cudaMallocAsync((void**)&y, (size_t)total_out_rows * L.U * sizeof(float), st));
  unpad_rows_kernel<<<(unsigned)batch, 256, 0, st>>>(yp, L.U, L.U, (const long long*)in_offsets_dev, poffs, batch, y);
  KTF_LAUNCH_OK();
  if (stats_dev) {
    rc = ktf::stats_sums_f32(y, in_offsets_dev, batch, L.U, stats_dev, st);
    if (rc != KTF_OK) return rc;
  }
  if (y_dev == nullptr) KTF_CUDA(cudaFreeAsync(y, st));
  KTF_CUDA(cudaFreeAsync(yp, st));
  KTF_CUDA(cudaFreeAsync(spliced, st));
  KTF_CUDA(cudaFreeAsync(rowmap, st));
  KTF_CUDA(cudaFreeAsync(poffs, st));
  return KTF_OK;
}

}  // namespace ktf

// ---------------------------------------------------------------------------------------------------
// Whole-stack API: [affine(+ReLU+BN)] x n1 -> StatsPooling(reduce-all) -> [affine(+ReLU+BN)] x n2
// with bf16 activations that never leave the padded-row layout between layers.
// ---------------------------------------------------------------------------------------------------

struct ktf_tdnn_stack {
  std::vector<ktf_affine*> layers;       // borrowed handles (must be KTF_PREC_BF16)
  int stats_after = -1;                  // index of the layer followed by stats pooling (-1 = none)
  int include_std = 1;
  float stats_eps = 1e-10f;
};

extern "C" {

int ktf_tdnn_stack_create(ktf_affine* const* layers, int32_t num_layers, int32_t stats_after_layer,
                          int32_t include_std, float stats_epsilon, ktf_tdnn_stack** out) {
  KTF_CHECK_ARG(layers && out && num_layers > 0, "ktf_tdnn_stack_create: bad arguments");
  KTF_CHECK_ARG(stats_after_layer >= -1 && stats_after_layer < num_layers, "stats_after_layer out of range");
  int rc = check_arch();
  if (rc != KTF_OK) return rc;
  int dim = -1;
  for (int i = 0; i < num_layers; ++i) {
    KTF_CHECK_ARG(layers[i] && layers[i]->tc, "layer %d was not created with KTF_PREC_BF16", i);
    const ktf_affine_cfg& c = layers[i]->cfg;
    KTF_CHECK_ARG(!c.padding_valid && c.subsampling_factor == 1,
                  "layer %d: the tcgen05 stack supports padding=SAME, subsampling_factor=1", i);
    KTF_CHECK_ARG(dim < 0 || c.in_dim == dim, "layer %d expects in_dim %d, previous layer produces %d", i, c.in_dim,
                  dim);
    dim = c.out_dim;
    if (i == stats_after_layer) dim = include_std ? 2 * dim : dim;
    if (i > stats_after_layer && stats_after_layer >= 0)
      KTF_CHECK_ARG(c.num_context == 1 && c.context[0] == 0, "layer %d (after stats pooling) must have context [0]", i);
  }
  ktf_tdnn_stack* s = new ktf_tdnn_stack();
  s->layers.assign(layers, layers + num_layers);
  s->stats_after = stats_after_layer;
  s->include_std = include_std;
  s->stats_eps = stats_epsilon;
  *out = s;
  return KTF_OK;
}

void ktf_tdnn_stack_destroy(ktf_tdnn_stack* s) { delete s; }

int32_t ktf_tdnn_stack_out_dim(const ktf_tdnn_stack* s) {
  if (!s) return 0;
  int dim = s->layers.back()->cfg.out_dim;
  if (s->stats_after == (int)s->layers.size() - 1) dim = s->include_std ? 2 * dim : dim;
  return dim;
}

int ktf_tdnn_stack_forward(const ktf_tdnn_stack* s, const float* feats_dev, const int64_t* offsets_dev,
                           int64_t batch, int64_t total_rows, float* out_dev, void* stream) {
  KTF_CHECK_ARG(s && feats_dev && offsets_dev && out_dev, "ktf_tdnn_stack_forward: null argument");
  if (batch <= 0 || total_rows <= 0) return KTF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long prow = total_rows + 2LL * kHalo * batch;
  const int nl = (int)s->layers.size();

  long long* poffs = nullptr;
  int* rowmap = nullptr;
  KTF_CUDA(cudaMallocAsync((void**)&poffs, (batch + 1) * sizeof(long long), st));
  KTF_CUDA(cudaMallocAsync((void**)&rowmap, prow * sizeof(int), st));
  build_padded_kernel<<<(unsigned)batch, 128, 0, st>>>((const long long*)offsets_dev, batch, poffs, rowmap);
  KTF_LAUNCH_OK();

  // activation ping-pong buffers sized for the widest layer
  long long max_ld = 8;
  for (int i = 0; i < nl; ++i) {
    const ktf_affine_cfg& c = s->layers[i]->cfg;
    max_ld = std::max(max_ld, round_up((long long)c.num_context * c.in_dim, 8));
    max_ld = std::max(max_ld, round_up(c.out_dim, 8));
  }
  __nv_bfloat16* buf[2] = {nullptr, nullptr};
  KTF_CUDA(cudaMallocAsync((void**)&buf[0], (size_t)prow * max_ld * sizeof(__nv_bfloat16), st));
  KTF_CUDA(cudaMallocAsync((void**)&buf[1], (size_t)prow * max_ld * sizeof(__nv_bfloat16), st));
  __nv_bfloat16* pooled = nullptr;

  const __nv_bfloat16* cur = nullptr;   // current activations (bf16) and their geometry
  long long cur_ld = 0, cur_rows = prow;
  const int* cur_rowmap = rowmap;
  bool per_frame = true;                // false once stats pooling collapsed the time axis
  int which = 0;
  int rc = KTF_OK;

  for (int i = 0; i < nl && rc == KTF_OK; ++i) {
    const ktf_affine* a = s->layers[i];
    const TcLayer& L = *static_cast<const TcLayer*>(a->tc);
    const long long cols = (long long)L.K * L.D;
    const bool last = (i == nl - 1) && (s->stats_after != i);
    const __nv_bfloat16* A = cur;
    long long a_ld = cur_ld, a_cols = L.D;
    bool spliced = false;
    if (i == 0) {
      // first layer: fp32 ragged features -> bf16 spliced padded rows (also covers D % 64 != 0)
      __nv_bfloat16* sp = buf[which];
      const long long ld = round_up(cols, 8);
      splice_kernel<float><<<(unsigned)prow, 128, 0, st>>>(feats_dev, L.D, L.D, 0, (const long long*)offsets_dev,
                                                           poffs, batch, L.K, L.d_ctx, sp, ld);
      KTF_LAUNCH_OK();
      A = sp;
      a_ld = ld;
      a_cols = cols;
      spliced = true;
      which ^= 1;
    } else if (per_frame && !L.implicit) {
      __nv_bfloat16* sp = buf[which];
      const long long ld = round_up(cols, 8);
      splice_kernel<__nv_bfloat16><<<(unsigned)prow, 128, 0, st>>>(cur, L.D, cur_ld, 1, (const long long*)offsets_dev,
                                                                   poffs, batch, L.K, L.d_ctx, sp, ld);
      KTF_LAUNCH_OK();
      A = sp;
      a_ld = ld;
      a_cols = cols;
      spliced = true;
      which ^= 1;
    } else if (!per_frame) {
      a_cols = L.D;          // context [0] on pooled rows
      spliced = true;
    }
    if (last) {
      // final layer writes fp32; per-frame outputs are un-padded into the caller's layout
      if (per_frame) {
        float* yp = nullptr;
        KTF_CUDA(cudaMallocAsync((void**)&yp, (size_t)prow * L.U * sizeof(float), st));
        rc = run_layer(L, a, A, a_ld, a_cols, cur_rows, cur_rowmap, yp, L.U, false, spliced, st);
        if (rc == KTF_OK) {
          unpad_rows_kernel<<<(unsigned)batch, 256, 0, st>>>(yp, L.U, L.U, (const long long*)offsets_dev, poffs,
                                                            batch, out_dev);
          KTF_LAUNCH_OK();
        }
        KTF_CUDA(cudaFreeAsync(yp, st));
      } else {
        rc = run_layer(L, a, A, a_ld, a_cols, cur_rows, nullptr, out_dev, L.U, false, spliced, st);
      }
      break;
    }
    __nv_bfloat16* y = buf[which];
    const long long y_ld = round_up(L.U, 8);
    rc = run_layer(L, a, A, a_ld, a_cols, cur_rows, per_frame ? cur_rowmap : nullptr, y, y_ld, true, spliced, st);
    if (rc != KTF_OK) break;
    cur = y;
    cur_ld = y_ld;
    which ^= 1;
    if (i == s->stats_after) {
      const int od = s->include_std ? 2 * L.U : L.U;
      const long long p_ld = round_up(od, 8);
      const bool final_stats = (i == nl - 1);
      if (!final_stats) KTF_CUDA(cudaMallocAsync((void**)&pooled, (size_t)batch * p_ld * sizeof(__nv_bfloat16), st));
      dim3 grid((unsigned)batch, (unsigned)((L.U + 127) / 128));
      stats_padded_kernel<<<grid, 128, 0, st>>>(cur, cur_ld, L.U, poffs, s->include_std, s->stats_eps,
                                                final_stats ? nullptr : pooled, final_stats ? out_dev : nullptr,
                                                final_stats ? od : p_ld);
      KTF_LAUNCH_OK();
      cur = pooled;
      cur_ld = p_ld;
      cur_rows = batch;
      per_frame = false;
    }
  }

  if (pooled) cudaFreeAsync(pooled, st);
  cudaFreeAsync(buf[1], st);
  cudaFreeAsync(buf[0], st);
  cudaFreeAsync(rowmap, st);
  cudaFreeAsync(poffs, st);
  return rc;
}

}  // extern "C"
