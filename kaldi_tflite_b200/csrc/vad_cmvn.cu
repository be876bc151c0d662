// Energy VAD (mask + stable compaction) and sliding-window CMVN for ragged batches.
//
// Replaces (file:line under /root/reference/kaldi_tflite/lib/):
//   layers/dsp/vad.py:156-203, the tf.gather_nd compaction of
//   models/kaldi/xvector_extractor.py:163-165, layers/normalization/cmvn.py:186-250.
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ double block_sum_double(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? scratch[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[0] = t;
  }
  __syncthreads();
  const double r = scratch[0];
  __syncthreads();
  return r;
}

// One CTA per utterance.  The energy column is strided by the feature dimension in global memory: utterances of up to
// kVadStage frames copy it into shared memory once (the mean pass), so the 2 * ctx + 1 reads per vote are LDS.
constexpr int kVadStage = 4096;

__global__ void vad_mask_kernel(const float* __restrict__ feats, int dim, int coeff,
                                const long long* __restrict__ offs, float thr0, float mean_scale,
                                float prop_thr, int ctx, float* __restrict__ mask) {
  __shared__ double scratch[32];
  __shared__ float s_e[kVadStage];
  const long long r0 = offs[blockIdx.x];
  const int T = (int)(offs[blockIdx.x + 1] - r0);
  if (T <= 0) return;
  const float* eg = feats + r0 * dim + coeff;
  const bool staged = T <= kVadStage;
  auto energy = [&](int t) { return staged ? s_e[t] : eg[(long long)t * dim]; };

  double s = 0.0;                 // vad.py:162-166 -- mean accumulated exactly (fp64), rounded once
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float v = eg[(long long)t * dim];
    if (staged) s_e[t] = v;
    s += (double)v;
  }
  s = block_sum_double(s, scratch);         // (also the barrier that publishes s_e)
  float thr = thr0;
  if (mean_scale > 0.0f) {
    const float mean = (float)(s / (double)T);
    thr = thr0 + mean_scale * mean;
  }
  const int N = 2 * ctx + 1;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    bool keep;
    if (ctx == 0) {
      keep = energy(t) > thr;                                 // vad.py:168-174
    } else {
      float count = 0.0f;                                     // conv1d, SAME zero padding (:180-182)
      for (int k = -ctx; k <= ctx; ++k) {
        const int u = t + k;
        if (u >= 0 && u < T && energy(u) > thr) count += 1.0f;
      }
      float size = (float)N;                                  // edge window sizes (:124-135,187-193)
      if (t < ctx || t >= T - ctx) {
        // only the first / last `ctx` frames can be named by the reference's edge lists (python indices i and
        // -ctx + i, wrapping for utterances shorter than the context); interior frames skip the modulo loops
        for (int i = 0; i < ctx; ++i) {                       // left edge: index i, size ctx+1+i
          if (((i % T) + T) % T == t) size = (float)(ctx + 1 + i);
        }
        for (int i = 0; i < ctx; ++i) {                       // right edge: index -ctx+i, size 2ctx-i
          const int idx = -ctx + i;
          if ((((idx + T) % T) + T) % T == t) size = (float)(2 * ctx - i);
        }
      }
      keep = __fdiv_rn(count, size) >= prop_thr;              // :197-199
    }
    mask[r0 + t] = keep ? 1.0f : 0.0f;
  }
}

__global__ void vad_count_kernel(const float* __restrict__ mask, const long long* __restrict__ offs,
                                 long long* __restrict__ counts) {
  __shared__ double scratch[32];
  const long long r0 = offs[blockIdx.x];
  const int T = (int)(offs[blockIdx.x + 1] - r0);
  double c = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) c += (mask[r0 + t] != 0.0f) ? 1.0 : 0.0;
  c = block_sum_double(c, scratch);
  if (threadIdx.x == 0) counts[blockIdx.x] = (long long)(c + 0.5);
}

// Single CTA: exclusive scan of counts -> out_offsets (batch + 1).
__global__ void scan_counts_kernel(const long long* __restrict__ counts, long long batch,
                                   long long* __restrict__ out_offsets) {
  __shared__ long long s_part[1024];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < batch; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    const long long v = i < batch ? counts[i] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
      long long add = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    const long long incl = s_part[threadIdx.x];
    if (i < batch) out_offsets[i] = s_carry + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_offsets[batch] = s_carry;
}

// One CTA per utterance: stable compaction of kept rows.
// out[j, :] = feats[index[j], :] for the kept rows (count read from the device: out_offs[batch]).
__global__ void vad_gather_kernel(const float* __restrict__ feats, int dim, const long long* __restrict__ index,
                                  const long long* __restrict__ out_offs, long long batch, long long max_rows,
                                  float* __restrict__ out) {
  const long long kept = min(out_offs[batch], max_rows);
  if ((dim & 1) == 0 && kept * (dim >> 1) < 0x7fffffffLL) {
    // even feature dimension: 8-byte pieces (rows start 8-byte aligned), 32-bit index arithmetic
    const unsigned half = (unsigned)dim >> 1, total = (unsigned)kept * half;
    const float2* f2 = reinterpret_cast<const float2*>(feats);
    float2* o2 = reinterpret_cast<float2*>(out);
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const unsigned r = e / half, d = e - r * half;
      o2[e] = f2[index[r] * half + d];
    }
    return;
  }
  const long long total = kept * dim;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / dim;
    const int d = (int)(e - r * dim);
    out[e] = feats[index[r] * dim + d];
  }
}

__global__ void vad_compact_kernel(const float* __restrict__ feats, int dim,
                                   const float* __restrict__ mask,
                                   const long long* __restrict__ offs,
                                   const long long* __restrict__ out_offs,
                                   long long* __restrict__ index, float* __restrict__ out_feats) {
  // one CTA (256 threads) per utterance; chunks of 256 frames: ballot / popc prefix inside a warp, warp totals
  // through shared memory.  Writes the index list only; the rows are gathered by vad_gather_kernel.
  __shared__ int s_warp[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r0 = offs[blockIdx.x];
  const int T = (int)(offs[blockIdx.x + 1] - r0);
  const long long o0 = out_offs[blockIdx.x];
  int done = 0;                                   // rows kept in earlier chunks (same value in every thread)
  for (int base = 0; base < T; base += 256) {
    const int t = base + threadIdx.x;
    const bool keep = t < T && mask[r0 + t] != 0.0f;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = s_warp[w];
      before += (w < warp) ? c : 0;
      total += c;
    }
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (keep) {
      index[o0 + done + pos] = r0 + t;
    }
    __syncthreads();
    done += total;
    __syncthreads();
  }
}

// Sliding CMVN.  blockDim = (32 feature lanes, 8 sub-chunks); each y-slice owns kSub consecutive
// output frames of one utterance and carries the window sum along them.
constexpr int kCmvnSub = 128;   // frames per warp: the window sum is rebuilt once per sub-chunk, then slides
constexpr int kCmvnY = 8;

__global__ void cmvn_kernel(const float* __restrict__ in, int dim, const long long* __restrict__ offs,
                            int window, int norm_vars, int padding_valid,
                            const long long* __restrict__ out_offs, float* __restrict__ out) {
  // grid (batch, chunk groups); block (32 feature lanes, kCmvnY sub-chunks of kCmvnSub frames)
  const long long b = blockIdx.x;
  const long long r0 = offs[b];
  const int T = (int)(offs[b + 1] - r0);
  const int N = window;
  // output frame range of this utterance (cmvn.py:230-237)
  int ta = 0, tb = T;
  if (padding_valid) {
    ta = N / 2;
    tb = T - (N - 1) / 2;
    if (tb < 0) tb += T;            // python slice semantics for a negative stop
    if (tb < 0) tb = 0;
    if (tb > T) tb = T;
    if (ta > tb) ta = tb;
  }
  const long long o0 = out_offs ? out_offs[b] : r0;
  const int t0 = ta + (blockIdx.y * kCmvnY + threadIdx.y) * kCmvnSub;
  const int t1 = min(t0 + kCmvnSub, tb);
  if (t0 >= t1) return;
  const float* x = in + r0 * dim;

  for (int d = threadIdx.x; d < dim; d += 32) {
    if (T <= N) {                                            // global stats (cmvn.py:214-222)
      float s = 0.0f, s2 = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float v = x[(long long)t * dim + d];
        s += v;
        s2 += __fmul_rn(v, v);
      }
      const float mean = s / (float)T;
      float sd = 1.0f;
      if (norm_vars) sd = sqrtf(s2 / (float)T - __fmul_rn(mean, mean));
      for (int t = t0; t < t1; ++t) {
        float v = x[(long long)t * dim + d] - mean;
        if (norm_vars) v = v / sd;
        out[(o0 + (t - ta)) * dim + d] = v;
      }
    } else {                                                 // sliding window (cmvn.py:172-204)
      int ws = min(max(t0 - N / 2, 0), T - N);
      // window sum of the first frame of the sub-chunk: four independent partial sums (the loads are
      // independent; a single accumulator would serialise N dependent adds)
      float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
      int t = ws;
      for (; t + 4 <= ws + N; t += 4) {
        const float v0 = x[(long long)t * dim + d], v1 = x[(long long)(t + 1) * dim + d];
        const float v2 = x[(long long)(t + 2) * dim + d], v3 = x[(long long)(t + 3) * dim + d];
        p0 += v0; p1 += v1; p2 += v2; p3 += v3;
        q0 += __fmul_rn(v0, v0); q1 += __fmul_rn(v1, v1); q2 += __fmul_rn(v2, v2); q3 += __fmul_rn(v3, v3);
      }
      for (; t < ws + N; ++t) {
        const float v = x[(long long)t * dim + d];
        p0 += v;
        q0 += __fmul_rn(v, v);
      }
      float s = (p0 + p1) + (p2 + p3), s2 = (q0 + q1) + (q2 + q3);
      const float inv_n = 1.0f / (float)N;
#pragma unroll 4
      for (int tt = t0; tt < t1; ++tt) {
        const int want = min(max(tt - N / 2, 0), T - N);
        const float xc = x[(long long)tt * dim + d];
        if (want != ws) {  // advances by exactly one
          const float vo = x[(long long)ws * dim + d];
          const float vn = x[(long long)(ws + N) * dim + d];
          s += vn - vo;
          s2 += __fmul_rn(vn, vn) - __fmul_rn(vo, vo);
          ws = want;
        }
        const float mean = s * inv_n;
        float v = xc - mean;
        if (norm_vars) v = v / sqrtf(s2 * inv_n - __fmul_rn(mean, mean));
        out[(o0 + (tt - ta)) * dim + d] = v;
      }
    }
  }
}


// Staged variant for utterances longer than the window: a CTA owns `tc` output frames of one utterance, copies the
// tc + N - 1 input rows they depend on into shared memory with 16-byte asynchronous copies (one pass over HBM/L2,
// every load in flight at once), and then slides the window out of shared memory.  The arithmetic (four partial sums
// for the first window, add-new / subtract-old afterwards) is the one of cmvn_kernel.
constexpr int kCmvnStagedWarps = 8;
constexpr int kCmvnBlk = 32;   // rows per block sum

template <bool NORM_VARS>
__global__ void __launch_bounds__(kCmvnStagedWarps * 32)
cmvn_staged_kernel(const float* __restrict__ in, int dim, const long long* __restrict__ offs, int window,
                   int padding_valid, const long long* __restrict__ out_offs,
                   float* __restrict__ out, int tc) {
  constexpr int norm_vars = NORM_VARS ? 1 : 0;
  extern __shared__ __align__(16) float sx[];
  const long long b = blockIdx.x;
  const long long r0 = offs[b];
  const int T = (int)(offs[b + 1] - r0);
  const int N = window;
  int ta = 0, tb = T;
  if (padding_valid) {
    ta = N / 2;
    tb = T - (N - 1) / 2;
    if (tb < 0) tb += T;
    if (tb < 0) tb = 0;
    if (tb > T) tb = T;
    if (ta > tb) ta = tb;
  }
  const long long o0 = out_offs ? out_offs[b] : r0;
  const int c0 = ta + blockIdx.y * tc;
  const int c1 = min(c0 + tc, tb);
  if (c0 >= c1) return;
  const float* x = in + r0 * dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (T <= N) {   // global statistics (cmvn.py:214-222): short utterances, read straight from global memory
    const int per = (c1 - c0 + kCmvnStagedWarps - 1) / kCmvnStagedWarps;
    const int t0 = c0 + warp * per, t1 = min(t0 + per, c1);
    for (int d = lane; d < dim && t0 < t1; d += 32) {
      float s = 0.0f, s2 = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float v = x[(long long)t * dim + d];
        s += v;
        s2 += __fmul_rn(v, v);
      }
      const float mean = s / (float)T;
      float sd = 1.0f;
      if (norm_vars) sd = sqrtf(s2 / (float)T - __fmul_rn(mean, mean));
      for (int t = t0; t < t1; ++t) {
        float v = x[(long long)t * dim + d] - mean;
        if (norm_vars) v = v / sd;
        out[(o0 + (t - ta)) * dim + d] = v;
      }
    }
    return;
  }

  // rows [lo, hi) cover every window of the CTA's frames (and the frames themselves)
  const int lo = min(max(c0 - N / 2, 0), T - N);
  const int hi = min(max(c1 - 1 - N / 2, 0), T - N) + N;
  const float* src = x + (long long)lo * dim;
  const int count = (hi - lo) * dim;
  const int mis = (int)((reinterpret_cast<unsigned long long>(src) & 15ull) >> 2);   // floats past a 16-byte boundary
  float* sbase = sx + mis;                                                            // same misalignment in smem
  {
    const int head = mis ? min(4 - mis, count) : 0;
    for (int i = threadIdx.x; i < head; i += blockDim.x) sbase[i] = src[i];
    const int body4 = (count - head) >> 2;
    const unsigned sdst = (unsigned)__cvta_generic_to_shared(sbase + head);
    const float* gsrc = src + head;
    for (int i = threadIdx.x; i < body4; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst + 16u * i), "l"(gsrc + 4 * i) : "memory");
    for (int i = head + (body4 << 2) + threadIdx.x; i < count; i += blockDim.x) sbase[i] = src[i];
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
  }
  __syncthreads();

  // Column sums of the staged rows in blocks of kCmvnBlk: a warp's first window (N rows) is then a handful of block
  // sums plus the rows that stick out on both sides, instead of N dependent loads and adds per column.
  const float* xs = sbase - (long long)lo * dim;   // xs[t * dim + d] == x[t * dim + d] for lo <= t < hi
  const int nblk = (hi - lo) / kCmvnBlk;
  float* bsum = sx + (((size_t)(tc + N) * dim + 4 + 3) & ~(size_t)3);   // [nblk][dim] (+ [nblk][dim] of squares)
  float* bsq = bsum + (size_t)((tc + N) / kCmvnBlk + 1) * dim;
  for (int k = warp; k < nblk; k += kCmvnStagedWarps) {
    for (int d = lane; d < dim; d += 32) {
      const float* col = xs + (long long)(lo + k * kCmvnBlk) * dim + d;
      float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
      for (int r = 0; r < kCmvnBlk; r += 4) {
        const float v0 = col[r * dim], v1 = col[(r + 1) * dim], v2 = col[(r + 2) * dim], v3 = col[(r + 3) * dim];
        p0 += v0; p1 += v1; p2 += v2; p3 += v3;
        if (NORM_VARS) {
          q0 += __fmul_rn(v0, v0); q1 += __fmul_rn(v1, v1); q2 += __fmul_rn(v2, v2); q3 += __fmul_rn(v3, v3);
        }
      }
      bsum[k * dim + d] = (p0 + p1) + (p2 + p3);
      if (NORM_VARS) bsq[k * dim + d] = (q0 + q1) + (q2 + q3);
    }
  }
  __syncthreads();

  const int per = (c1 - c0 + kCmvnStagedWarps - 1) / kCmvnStagedWarps;
  const int t0 = c0 + warp * per, t1 = min(t0 + per, c1);
  if (t0 >= t1) return;
  const float inv_n = 1.0f / (float)N;
  const int H = N / 2;
  for (int d = lane; d < dim; d += 32) {
    int ws = min(max(t0 - H, 0), T - N);
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
    {
      // rows [ws, ws + N) = leading rows up to the next block boundary, whole blocks, trailing rows
      const int a0 = ws - lo;
      const int kb = (a0 + kCmvnBlk - 1) / kCmvnBlk;
      const int ke = min((a0 + N) / kCmvnBlk, nblk);
      int r_lead_end = ws + N, r_trail = ws + N;       // no whole block inside: everything is "leading"
      if (ke > kb) {
        r_lead_end = lo + kb * kCmvnBlk;
        r_trail = lo + ke * kCmvnBlk;
        int k = kb;
        for (; k + 2 <= ke; k += 2) {
          p0 += bsum[k * dim + d]; p1 += bsum[(k + 1) * dim + d];
          if (NORM_VARS) { q0 += bsq[k * dim + d]; q1 += bsq[(k + 1) * dim + d]; }
        }
        if (k < ke) {
          p0 += bsum[k * dim + d];
          if (NORM_VARS) q0 += bsq[k * dim + d];
        }
      }
      for (int r = ws; r < r_lead_end; ++r) {
        const float v = xs[(long long)r * dim + d];
        p2 += v;
        if (NORM_VARS) q2 += __fmul_rn(v, v);
      }
      for (int r = r_trail; r < ws + N; ++r) {
        const float v = xs[(long long)r * dim + d];
        p3 += v;
        if (NORM_VARS) q3 += __fmul_rn(v, v);
      }
    }
    float s = (p0 + p1) + (p2 + p3), s2 = (q0 + q1) + (q2 + q3);
    float* orow = out + (o0 + (t0 - ta)) * dim + d;
    int tt = t0;
    auto emit = [&](int t_) {
      const float mean = s * inv_n;
      float v = xs[t_ * dim + d] - mean;
      if (NORM_VARS) v = v / sqrtf(s2 * inv_n - __fmul_rn(mean, mean));
      *orow = v;
      orow += dim;
    };
    // head: frames whose window is still clamped at the start of the utterance (ws == 0)
    for (; tt < t1 && tt - H <= ws; ++tt) emit(tt);
    // interior: the window advances by exactly one row per frame
    const int t_mid = min(t1, T - N + H + 1);   // last frame + 1 with an unclamped window start
    const float* po = xs + ws * dim + d;        // row leaving the window
    const float* pn = po + N * dim;             // row entering it
    const float* px = xs + tt * dim + d;
#pragma unroll 4
    for (; tt < t_mid; ++tt) {
      const float vo = *po, vn = *pn;
      s += vn - vo;
      if (NORM_VARS) s2 += __fmul_rn(vn, vn) - __fmul_rn(vo, vo);
      po += dim; pn += dim;
      const float mean = s * inv_n;
      float v = *px - mean;
      px += dim;
      if (NORM_VARS) v = v / sqrtf(s2 * inv_n - __fmul_rn(mean, mean));
      *orow = v;
      orow += dim;
    }
    // tail: window clamped at the end of the utterance
    for (; tt < t1; ++tt) emit(tt);
  }
}

}  // namespace

extern "C" {

int ktf_vad_mask(const ktf_vad_cfg* cfg, const float* feats_dev, int32_t dim,
                 const int64_t* frame_offsets_dev, int64_t batch, int64_t total_frames,
                 float* mask_dev, void* stream) {
  KTF_CHECK_ARG(cfg && feats_dev && frame_offsets_dev && mask_dev, "ktf_vad_mask: null argument");
  KTF_CHECK_ARG(cfg->energy_mean_scale >= 0.0f, "`energy_mean_scale` must be >= 0");
  KTF_CHECK_ARG(cfg->frames_context >= 0, "`frames_context` must be >= 0");
  KTF_CHECK_ARG(cfg->proportion_threshold > 0.0f && cfg->proportion_threshold < 1.0f,
                "`proportion_threshold` must be between 0 and 1 (exlcusive)");
  KTF_CHECK_ARG(cfg->energy_coeff >= 0 && cfg->energy_coeff < dim, "energy_coeff out of range");
  (void)total_frames;
  if (batch <= 0) return KTF_OK;
  vad_mask_kernel<<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>(
      feats_dev, dim, cfg->energy_coeff, (const long long*)frame_offsets_dev, cfg->energy_threshold,
      cfg->energy_mean_scale, cfg->proportion_threshold, cfg->frames_context, mask_dev);
  KTF_LAUNCH_OK();
  return KTF_OK;
}

int64_t ktf_vad_compact_workspace(int64_t batch, int64_t total_frames) {
  (void)total_frames;
  return (batch + 1) * (int64_t)sizeof(long long);
}

int ktf_vad_compact(const float* feats_dev, int32_t dim, const float* mask_dev,
                    const int64_t* frame_offsets_dev, int64_t batch, int64_t total_frames,
                    int64_t* out_offsets_dev, int64_t* index_dev, float* out_feats_dev,
                    void* workspace_dev, void* stream) {
  KTF_CHECK_ARG(mask_dev && frame_offsets_dev && out_offsets_dev && index_dev && workspace_dev,
                "ktf_vad_compact: null argument");
  KTF_CHECK_ARG(out_feats_dev == nullptr || feats_dev != nullptr, "feats_dev required for the gather");
  if (batch <= 0) return KTF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long* counts = (long long*)workspace_dev;
  vad_count_kernel<<<(unsigned)batch, 256, 0, st>>>(mask_dev, (const long long*)frame_offsets_dev, counts);
  KTF_LAUNCH_OK();
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts, batch, (long long*)out_offsets_dev);
  KTF_LAUNCH_OK();
  vad_compact_kernel<<<(unsigned)batch, 256, 0, st>>>(feats_dev, dim, mask_dev,
                                                      (const long long*)frame_offsets_dev,
                                                      (const long long*)out_offsets_dev,
                                                      (long long*)index_dev, out_feats_dev);
  KTF_LAUNCH_OK();
  if (out_feats_dev != nullptr && total_frames > 0) {
    const long long elems = (long long)total_frames * dim;
    const long long blocks = std::min<long long>((elems + 255) / 256, (long long)ktf::num_sms() * 32);
    vad_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(feats_dev, dim, (const long long*)index_dev,
                                                        (const long long*)out_offsets_dev, batch, total_frames,
                                                        out_feats_dev);
    KTF_LAUNCH_OK();
  }
  return KTF_OK;
}

int ktf_cmvn_forward(const float* in_dev, int32_t dim, const int64_t* frame_offsets_dev,
                     int64_t batch, int64_t total_frames, int64_t max_frames, int32_t window,
                     int32_t norm_vars, int32_t padding_valid, const int64_t* out_offsets_dev,
                     float* out_dev, void* stream) {
  KTF_CHECK_ARG(in_dev && frame_offsets_dev && out_dev, "ktf_cmvn_forward: null argument");
  KTF_CHECK_ARG(window > 0, "`window` and `min_window` must be > 0");
  KTF_CHECK_ARG(!padding_valid || out_offsets_dev, "out_offsets_dev is required for VALID padding");
  if (batch <= 0 || total_frames <= 0 || max_frames <= 0) return KTF_OK;
  {
    // staged kernel when a CTA's rows fit in shared memory with >= 2 CTAs per SM
    // up to 256 frames per CTA, balanced over the longest utterance (998 frames -> 4 x 250: one CTA more per SM
    // than 3 x 256 + 230 because the shared memory of 250 + window rows fits four times)
    const long long gys = (max_frames + 255) / 256;
    const int tc = (int)((((max_frames + gys - 1) / gys) + 1) & ~1LL);
    // staged rows, then the block sums (and, for norm_vars, the block sums of squares)
    const size_t smem = ((((size_t)(tc + window) * dim + 4 + 3) & ~(size_t)3) +
                         (norm_vars ? 2 : 1) * (size_t)((tc + window) / kCmvnBlk + 1) * dim) * sizeof(float);
    if (smem <= 113 * 1024 && gys <= 65535) {
      static unsigned long long attr_set = 0;           // per device (function attributes are per context)
      if (ktf::first_use_on_device(&attr_set)) {
        KTF_CUDA(cudaFuncSetAttribute(cmvn_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        KTF_CUDA(cudaFuncSetAttribute(cmvn_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
      }
      auto kern = norm_vars ? cmvn_staged_kernel<true> : cmvn_staged_kernel<false>;
      kern<<<dim3((unsigned)batch, (unsigned)gys), kCmvnStagedWarps * 32, smem, (cudaStream_t)stream>>>(
          in_dev, dim, (const long long*)frame_offsets_dev, window, padding_valid,
          (const long long*)out_offsets_dev, out_dev, tc);
      KTF_LAUNCH_OK();
      return KTF_OK;
    }
  }
  const long long per_block = (long long)kCmvnY * kCmvnSub;
  const long long gy = (max_frames + per_block - 1) / per_block;
  KTF_CHECK_ARG(gy <= 65535, "max_frames too large for ktf_cmvn_forward");
  cmvn_kernel<<<dim3((unsigned)batch, (unsigned)gy), dim3(32, kCmvnY), 0, (cudaStream_t)stream>>>(
      in_dev, dim, (const long long*)frame_offsets_dev, window, norm_vars, padding_valid,
      (const long long*)out_offsets_dev, out_dev);
  KTF_LAUNCH_OK();
  return KTF_OK;
}

}  // extern "C"
