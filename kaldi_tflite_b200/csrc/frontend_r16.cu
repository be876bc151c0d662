// Fast path of the fused front-end for the common Kaldi geometry: 400-sample frames, 512-point FFT
// (25 ms @ 16 kHz), power spectrum, MFCC or log-mel output.  Same math and the same reference lines
// as frontend.cu (framing.py:243-265, windowing.py:180-209, filterbank.py:225-242, dct.py:175-176,
// mfcc.py:197-244); what changes is the instruction budget per frame:
//
//   * the real FFT of length 512 is a 256-point complex FFT factored 16 x 16.  A frame is owned by 8
//     lanes (4 frames per warp); in both stages a lane runs TWO 16-point FFTs held entirely in
//     registers (radix-4 x radix-4, compile-time twiddles):
//       stage 1  lane j owns n2 = 2j, 2j+1 : one LDS.128 per 32-sample row feeds both FFTs
//       twiddle  W_256^(n2 k1) from a per-lane table, written as [k1][n2] rows with STS.128
//       stage 2  lane j owns the columns k1 = j and 16-j (lane 0: 0 and 8), read back with LDS.128.
//     Columns j and 16-j hold exactly the bin pairs (k, 256-k) the real-FFT untangling needs, so the
//     power spectrum is produced without any further exchange.  Lane 0's two self-paired columns use
//     the same instruction stream with a handful of selects (no divergent path), the Nyquist bin is
//     never computed (its mel weight is identically zero, filterbank.py:176-187).
//   * all complex arithmetic runs on packed FP32 pairs (add/sub/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2, sm_100):
//     one issue slot per complex add, two per complex multiply; the half-swap / per-half-negate / broadcast operand
//     modifiers of the packed SASS forms absorb the shuffles a complex product needs.
//   * every shared-memory address is "per-lane base + immediate"; runtime configuration that the old
//     kernel tested per frame is a template parameter or folded into the tables (lifter into the DCT
//     matrix, the 1/4 of the untangling into the mel weights).
//   * mel bank: the phase re-maps lanes so that a quarter-warp (the conflict domain of LDS.128) holds schedule rows
//     2q, 2q+1 of all 4 frames; even rows only read even 4-bin chunks, odd rows odd ones (host schedule), and the frame
//     tiles sit 2 bank groups apart, so every power-tile load is bank-conflict free by construction.  Four units per
//     trip, all loads in flight before the first product.
//   * DCT: for the mirror-symmetric DCT-II (checked on the host) lane c keeps 16 coefficients and reads the sums /
//     differences of mirrored log-mel pairs; any other matrix uses 32 registers, > 32 mel bins a shared-memory table.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "frontend_internal.cuh"

using namespace ktf_fe;

// the per-mode tails of the item loop end in `continue`; the code after them is dead for that instantiation only
#pragma nv_diag_suppress 128

namespace {

constexpr int kW = 400;               // frame width of the fast path
// 4 warps per CTA, 3 CTAs per SM (168 registers): registers are allocated in units of 4 warps, so the next step up
// would be 16 resident warps at 128 registers, which spills the two 16-point FFTs (measured: 14 warps as 2 x 7 end
// up as ONE resident CTA)
#ifndef KTF_R16_MINB
#define KTF_R16_MINB 3
#endif
constexpr int kR16Warps = 4;
constexpr int kR16Threads = kR16Warps * 32;
constexpr int kRows = 13;             // ceil(400 / 32) rows of 32 samples
constexpr int kTailLanes = 4;         // lanes j < 4 own valid samples in row 12 (400 = 12 * 32 + 16)
constexpr int kWinPad = 416;          // window table padded with zeros to 13 * 32
constexpr int kRowStride = 36;        // floats per [k1] row of the exchange tile (32 + 4: LDS.128 conflict-free)
constexpr int kTile = 584;            // floats per frame tile (16 * 36 = 576, +8 so that frames shift by 8 banks)
constexpr int kOffWin = 0;
// per-lane table rows are padded to a stride of 4 (mod 32) floats: the 8 lanes of a frame (one LDS.128
// quarter-warp) then read 8 different bank groups.
// Twiddles: read from exactly rounded per-lane tables (0), or derived in the kernel from a smaller table and
// compile-time constants (1: fewer LDS bytes and registers, but two more roundings per twiddle, which shows on mel
// bins 60 dB below the frame peak).
#ifndef KTF_R16_TW1_DERIVED
#define KTF_R16_TW1_DERIVED 0
#endif
#ifndef KTF_R16_TW2_DERIVED
#define KTF_R16_TW2_DERIVED 0
#endif
#if KTF_R16_TW1_DERIVED
constexpr int kTw1Stride = 36;                  // floats per lane: [16 k1][2] + 4   (W_256^(2j k1))
#else
constexpr int kTw1Stride = 68;                  // floats per lane: [16 k1][4] + 4   (W_256^(2j k1), W_256^((2j+1) k1))
#endif
#if KTF_R16_TW2_DERIVED
constexpr int kTw2Stride = 4;                   // floats per lane: untangling base twiddles for e < 8 and e >= 8
#else
constexpr int kTw2Stride = 36;                  // floats per lane: [16 e][2] + 4
#endif
constexpr int kOffTw1 = kOffWin + kWinPad;
constexpr int kOffTw2 = kOffTw1 + 8 * kTw1Stride;
constexpr int kOffUnits = kOffTw2 + 8 * kTw2Stride;  // [8 lanes][SD] unit descriptors, then [8 lanes][SW] unit weights
__host__ __device__ constexpr int pad4mod32(int x) { return x + (((4 - x) % 32) + 32) % 32; }

// Packed FP32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): one issue slot per complex add / half a complex multiply.  The
// operand modifiers of the SASS forms (half swap, per-half negation, scalar broadcast) absorb the shuffles that a
// complex product needs, so the packing costs no extra moves.
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};"
      " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
// cos / sin of 2 pi k / 256 and 2 pi e / 32 for k, e < 16 (folded to immediates by the unrolled loops)
__device__ constexpr float kCos256[16] = {1.0f, 0.99969881869620425f, 0.99879545620517241f, 0.99729045667869021f, 0.99518472667219693f, 0.99247953459870997f, 0.98917650996478101f, 0.98527764238894122f, 0.98078528040323043f, 0.97570213003852857f, 0.97003125319454397f, 0.96377606579543984f, 0.95694033573220882f, 0.94952818059303667f, 0.94154406518302081f, 0.93299279883473896f};
__device__ constexpr float kSin256[16] = {0.0f, 0.024541228522912288f, 0.049067674327418015f, 0.073564563599667426f, 0.098017140329560604f, 0.1224106751992162f, 0.14673047445536175f, 0.17096188876030122f, 0.19509032201612825f, 0.2191012401568698f, 0.24298017990326387f, 0.26671275747489837f, 0.29028467725446233f, 0.31368174039889152f, 0.33688985339222005f, 0.35989503653498811f};
__device__ constexpr float kCos32[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
__device__ constexpr float kSin32[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f, 0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f, 0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }
// (a.x b.x - a.y b.y, a.y b.x + a.x b.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return fma2(a, make_float2(b.x, b.x), mul2(make_float2(a.y, a.x), make_float2(-b.y, b.y)));
}
// a + (-i) b = (a.x + b.y, a.y - b.x)  and  a - (-i) b
__device__ __forceinline__ float2 add_mi(float2 a, float2 b) {
  return fma2(make_float2(b.y, b.x), make_float2(1.0f, -1.0f), a);
}
__device__ __forceinline__ float2 sub_mi(float2 a, float2 b) {
  return fma2(make_float2(b.y, b.x), make_float2(-1.0f, 1.0f), a);
}

// x * W_16^M with compile-time M (W_16 = exp(-2 pi i / 16)).
template <int M>
__device__ __forceinline__ float2 mul_w16(float2 x) {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  if constexpr (M == 0) return x;
  else if constexpr (M == 1) return cmul(x, make_float2(c1, -s1));
  else if constexpr (M == 2) return cmul(x, make_float2(h, -h));
  else if constexpr (M == 3) return cmul(x, make_float2(s1, -c1));
  else if constexpr (M == 4) return make_float2(x.y, -x.x);
  else if constexpr (M == 6) return cmul(x, make_float2(-h, -h));
  else if constexpr (M == 9) return cmul(x, make_float2(-c1, s1));
  else return x;
}

// 4-point DFT of (a, b, c, d) in place: a <- y0, b <- y1, c <- y2, d <- y3.
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
  a = cadd(s0, s2);
  c = csub(s0, s2);
  b = add_mi(s1, s3);
  d = sub_mi(s1, s3);
}

// dft4 with d == 0 on input (x + 0 is not folded by the compiler: -0 + 0 = +0 under IEEE rules).
__device__ __forceinline__ void dft4_d0(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = b;
  a = cadd(s0, s2);
  c = csub(s0, s2);
  b = add_mi(s1, s2);
  d = sub_mi(s1, s2);
}

// 16-point complex FFT in registers (radix 4 x 4, decimation in frequency).  Input natural order,
// output X[k] at x[pos16(k)].
__host__ __device__ constexpr int pos16(int k) { return 4 * (k & 3) + (k >> 2); }

// TAIL0: x[13], x[14], x[15] are zero on input (400-sample frames fill 13 of the 16 rows).
template <bool TAIL0>
__device__ __forceinline__ void fft16(float2 (&x)[16]) {
  // layer A: butterflies over (i, i+4, i+8, i+12); output q of butterfly i, times W_16^(i q), lands at 4q + i
  dft4(x[0], x[4], x[8], x[12]);
  if constexpr (TAIL0) {
    dft4_d0(x[1], x[5], x[9], x[13]);
    dft4_d0(x[2], x[6], x[10], x[14]);
    dft4_d0(x[3], x[7], x[11], x[15]);
  } else {
    dft4(x[1], x[5], x[9], x[13]);
    dft4(x[2], x[6], x[10], x[14]);
    dft4(x[3], x[7], x[11], x[15]);
  }
  x[5] = mul_w16<1>(x[5]);   x[9] = mul_w16<2>(x[9]);   x[13] = mul_w16<3>(x[13]);
  x[6] = mul_w16<2>(x[6]);   x[10] = mul_w16<4>(x[10]); x[14] = mul_w16<6>(x[14]);
  x[7] = mul_w16<3>(x[7]);   x[11] = mul_w16<6>(x[11]); x[15] = mul_w16<9>(x[15]);
  // layer B: 4-point DFT over i inside each group q; output r of group q is X[4r + q] at 4q + r
  dft4(x[0], x[1], x[2], x[3]);
  dft4(x[4], x[5], x[6], x[7]);
  dft4(x[8], x[9], x[10], x[11]);
  dft4(x[12], x[13], x[14], x[15]);
}

// K consecutive mel units of one schedule row: descriptors (first bin | filter slot << 16, keep flag), then the 2 K
// power chunks and 2 K weight chunks, all loaded before the first product.  `u` is a multiple of 4.
template <int K>
__device__ __forceinline__ void mel_trip(int u, const int4* udesc4, const float4* uwts, const float* mel_tile,
                                         float* mel_acc, float& run, float& keep) {
  const int4 da = udesc4[u >> 1];
  int4 db = make_int4(0, 0, 0, 0);
  if (K > 2) db = udesc4[(u >> 1) + 1];
  const int dx[4] = {da.x, da.z, db.x, db.z};
  const float kp[4] = {__int_as_float(da.y), __int_as_float(da.w), __int_as_float(db.y), __int_as_float(db.w)};
  float4 q0[K], q1[K], w0[K], w1[K];
#pragma unroll
  for (int t = 0; t < K; ++t) {
    const float* q = mel_tile + (dx[t] & 0xffff);
    q0[t] = *reinterpret_cast<const float4*>(q);
    q1[t] = *reinterpret_cast<const float4*>(q + 4);
    w0[t] = uwts[2 * (u + t)];
    w1[t] = uwts[2 * (u + t) + 1];
  }
#pragma unroll
  for (int t = 0; t < K; ++t) {
    float2 acc0 = mul2(make_float2(w0[t].x, w0[t].y), make_float2(q0[t].x, q0[t].y));
    float2 acc1 = mul2(make_float2(w1[t].x, w1[t].y), make_float2(q1[t].x, q1[t].y));
    acc0 = fma2(make_float2(w0[t].z, w0[t].w), make_float2(q0[t].z, q0[t].w), acc0);
    acc1 = fma2(make_float2(w1[t].z, w1[t].w), make_float2(q1[t].z, q1[t].w), acc1);
    acc0 = add2(acc0, acc1);
    // the only serial dependency between units: keep = 1 inside a filter, 0 after its last unit
    run = fmaf(run, keep, acc0.x + acc0.y);
    keep = kp[t];
    mel_acc[dx[t] >> 16] = run;   // partial sums are overwritten by the filter's last unit (same lane, program order)
  }
}

// DCT_REG 1 / 2 (MFCC with <= 32 mel bins): lane c keeps column c of the DCT matrix in registers and produces cepstrum
// c of all 4 frames, so the DCT reads no table at all (only broadcast loads of the log-mel rows).
// PCM16: the input is int16 PCM; the span buffer holds the raw 16-bit samples (half the HBM / L2 / smem bytes) and
// they are converted when the window is applied.
// DCT_REG == 2: the matrix has the DCT-II mirror symmetry D[M-1-i][c] = (-1)^c D[i][c] (checked on the host), so
// even cepstra are dot products with s[i] = lm[i] + lm[M-1-i] and odd ones with d[i] = lm[i] - lm[M-1-i]: half the
// coefficient registers, half the broadcast loads and half the FMAs.
template <int OUTPUT, bool RAW_ENERGY, int DCT_REG, bool PCM16>
__global__ void __launch_bounds__(kR16Threads, KTF_R16_MINB) frontend_r16_kernel(const FrontendArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int NU = a.r16_nf;                        // mel units (8 bins each) per lane
  const int SD = pad4mod32(2 * ((NU + 3) & ~3)), SW = pad4mod32(8 * NU);
  const int M = a.M;
  float* s_win = smem + kOffWin;
#if KTF_R16_TW1_DERIVED
  const float2* s_tw1 = reinterpret_cast<const float2*>(smem + kOffTw1);
#else
  const float4* s_tw1 = reinterpret_cast<const float4*>(smem + kOffTw1);
#endif
  const float4* s_tw2 = reinterpret_cast<const float4*>(smem + kOffTw2);
  const float* s_udesc = smem + kOffUnits;
  const float* s_uwts = s_udesc + 8 * SD;
  const float* s_dct = s_uwts + 8 * SW;
  float* s_warp0 = smem + a.r16_blob_floats;

  const int span_p = ((a.span + 3) & ~3) + 16;     // row 12 of the last frame reads 16 floats past the span
  const int LMS = DCT_REG ? 36 : ((M + 3) & ~3) + 4;   // log-mel row stride (one spare slot for padding units)
  const int out_row = (OUTPUT == KTF_OUT_MFCC) ? a.Kc : M;
  const int out_sz = (4 * out_row + 3) & ~3;
  const int warp_floats = (span_p + 4 * kTile + 4 * LMS + out_sz + 2 + 31) & ~31;   // 128-byte aligned per-warp regions
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_span = s_warp0 + warp * warp_floats;
  float* s_T = s_span + span_p;
  float* s_LM = s_T + 4 * kTile;             // mel sums, then log-mel [f][LMS]
  float* s_out = s_LM + 4 * LMS;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_out + ((out_sz + 1) & ~1));   // span mbarrier
  unsigned bar_parity = 0;                   // phase of the NEXT TMA completion to wait for
  bool cur_tma = false;                      // the span of `cur` was requested with a bulk copy (else cp.async)
  if (lane == 0) span_mbar_init(s_bar);
  __syncwarp();
  // Stages the span of `it`: ONE bulk asynchronous copy (TMA) when it lies inside the utterance and is 16-byte aligned --
  // every interior item of the 16 kHz geometry -- else the per-lane cp.async / mirrored path.
  auto stage = [&](const Item& it) {
    const long long s0 = it.frame0 * a.shift - a.edge_off;
    const unsigned bytes = (unsigned)a.span * (PCM16 ? 2u : 4u);
    const char* src = PCM16 ? reinterpret_cast<const char*>(a.wav16 + it.utt_base + s0)
                            : reinterpret_cast<const char*>(a.wav + it.utt_base + s0);
    const bool bulk = s0 >= 0 && it.utt_len - s0 >= a.span && (bytes & 15u) == 0 &&
                      (reinterpret_cast<unsigned long long>(src) & 15ull) == 0;
    if (bulk) {
      if (lane == 0) span_bulk_load(s_span, src, bytes, s_bar);
    } else if (PCM16) {
      stage_span16(a, it, reinterpret_cast<short*>(s_span), lane);
    } else {
      stage_span(a, it, s_span, lane);
    }
    return bulk;
  };

  const long long warp_global = (long long)blockIdx.x * kR16Warps + warp;
  const long long warp_stride = (long long)gridDim.x * kR16Warps;

  Item cur;
  long long item = warp_global;
  if (item < a.total_groups) {
    cur = decode_item(a, item);
    cur_tma = stage(cur);
  }
  for (int i = threadIdx.x; i < a.r16_blob_floats; i += kR16Threads) smem[i] = a.r16_blob[i];
  for (int i = lane; i < 4 * LMS; i += 32) s_LM[i] = 0.0f;   // slots >= M stay finite
  __syncthreads();

  const int f = lane >> 3;
  const int j = lane & 7;
  const bool j0 = (j == 0);
  const bool tail_ok = j < kTailLanes;

  // per-lane bases: everything below is base + immediate
  const float* fr = s_span + f * a.shift + 4 * j;
  const short* fr16 = reinterpret_cast<const short*>(s_span) + f * a.shift + 4 * j;
  const float* wn = s_win + 4 * j;
#if KTF_R16_TW1_DERIVED
  const float2* tw1 = s_tw1 + j * (kTw1Stride / 2);
#else
  const float4* tw1 = s_tw1 + j * (kTw1Stride / 4);
#endif
#if KTF_R16_TW2_DERIVED
  const float4 tw2 = s_tw2[j];   // (base for e < 8, base for e >= 8): W_512^j, lane 0: 1 and W_512^8
#else
  const float4* tw2 = s_tw2 + j * (kTw2Stride / 4);
#endif
  float* tile = s_T + f * kTile;
  float* tile_w = tile + 4 * j;
  const int col_a = j, col_b = j0 ? 8 : 16 - j;
  const float* row_a = tile + col_a * kRowStride;
  const float* row_b = tile + col_b * kRowStride;
  // power-spectrum bins written by evaluation e: A + 16 e and B + 16 (15 - e)
  float* qa_lo = tile + j;                       // e < 8
  float* qa_hi = tile + (j0 ? 8 : j);            // e >= 8
  float* qb_lo = tile + 16 - j;                  // e < 8  (lane 0: 256 - 16 e)
  float* qb_hi = tile + (j0 ? 8 : 16 - j);       // e >= 8
  // mel phase: a quarter-warp (the conflict domain of LDS.128) holds schedule rows 2q, 2q+1 of all 4 frames.  The
  // frame tiles are 2 bank groups apart and rows of opposite parity only ever read chunks of opposite parity (host
  // schedule), so the 8 lanes always hit 8 different bank groups.
  const int mf = (lane >> 1) & 3, mj = (lane & 1) | ((lane >> 3) << 1);
  const int4* udesc4 = reinterpret_cast<const int4*>(s_udesc + mj * SD);
  const float4* uwts = reinterpret_cast<const float4*>(s_uwts + mj * SW);
  const float* mel_tile = s_T + mf * kTile;
  float* mel_acc = s_LM + mf * LMS;
  float* lm_acc = s_LM + f * LMS;
  float* lm_row = (OUTPUT == KTF_OUT_MFCC) ? s_LM + f * LMS : s_out + f * M;
  const float pc = a.preemph > 0.0f ? a.preemph : 0.0f;
  const int prev_lane = (lane & 24) | ((j + 7) & 7);   // the lane that owns the 4 samples before mine
  constexpr int kDregs = DCT_REG == 2 ? 16 : (DCT_REG == 1 ? 32 : 1);
  float dreg[kDregs];
  if (DCT_REG) {
    const int rows = DCT_REG == 2 ? (M + 1) / 2 : M;
#pragma unroll
    for (int i = 0; i < kDregs; ++i) dreg[i] = (i < rows && lane < a.Kc) ? s_dct[i * 32 + lane] : 0.0f;
  }

  for (; item < a.total_groups; item += warp_stride) {
    if (cur_tma) {
      span_mbar_wait(s_bar, bar_parity);
      bar_parity ^= 1u;
    } else {
      cp_async_wait_all();
      __syncwarp();
    }

    // ---- windowing (windowing.py:180-209): rows of 32 samples, lane j owns samples 4j .. 4j+3 of a row
    float2 ze[16], zo[16];
    float esum = 0.0f;
    {
      float4 xv[kRows];
      float xm[kRows];
      float2 sum2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (PCM16) {
          // int16 -> float without I2F (quarter-rate pipe): x ^ 0x8000 = x + 32768 as an unsigned 16-bit value u,
          // placed in the mantissa of 2^23 it reads 8388608 + u, and subtracting 8421376 is exact
          const uint2 raw = *reinterpret_cast<const uint2*>(fr16 + 32 * r);
          const unsigned ua = raw.x ^ 0x80008000u, ub = raw.y ^ 0x80008000u;
          const float2 bias2 = make_float2(8421376.0f, 8421376.0f);
          const float2 lo = sub2(make_float2(__uint_as_float(__byte_perm(ua, 0x4B000000u, 0x7610)),
                                             __uint_as_float(__byte_perm(ua, 0x4B000000u, 0x7632))), bias2);
          const float2 hi = sub2(make_float2(__uint_as_float(__byte_perm(ub, 0x4B000000u, 0x7610)),
                                             __uint_as_float(__byte_perm(ub, 0x4B000000u, 0x7632))), bias2);
          xv[r] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          xv[r] = *reinterpret_cast<const float4*>(fr + 32 * r);
        }
        if (r == kRows - 1 && !tail_ok) xv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        // sample 32r + 4j - 1 is the last of the previous lane's four (lane 0: lane 7's four of the previous row)
        const float give = (r > 0 && j == 7) ? xv[r - 1].w : xv[r].w;
        xm[r] = __shfl_sync(0xffffffffu, give, prev_lane);
        sum2 = add2(sum2, add2(make_float2(xv[r].x, xv[r].y), make_float2(xv[r].z, xv[r].w)));
      }
      float mean = 0.0f;
      if (a.remove_dc) mean = group_sum8(sum2.x + sum2.y) / (float)kW;
      const float2 mean2 = make_float2(mean, mean);
      float2 e2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(wn + 32 * r);
        float2 d01 = sub2(make_float2(xv[r].x, xv[r].y), mean2);
        float2 d23 = sub2(make_float2(xv[r].z, xv[r].w), mean2);
        float dm = xm[r] - mean;
        if (r == 0) dm = j0 ? d01.x : dm;
        if (r == kRows - 1 && !tail_ok) { d01 = make_float2(0.f, 0.f); d23 = make_float2(0.f, 0.f); dm = 0.f; }
        if (RAW_ENERGY) { e2 = fma2(d01, d01, e2); e2 = fma2(d23, d23, e2); }
        // pre-emphasis pairs (d[i-1], d[i]) straddle the packed pairs, so that step stays scalar
        const float2 t01 = make_float2(fmaf(-pc, dm, d01.x), fmaf(-pc, d01.x, d01.y));
        const float2 t23 = make_float2(fmaf(-pc, d01.y, d23.x), fmaf(-pc, d23.x, d23.y));
        const float2 y01 = mul2(t01, make_float2(w.x, w.y));
        const float2 y23 = mul2(t23, make_float2(w.z, w.w));
        if (!RAW_ENERGY) { e2 = fma2(y01, y01, e2); e2 = fma2(y23, y23, e2); }
        ze[r] = y01;
        zo[r] = y23;
      }
      esum = e2.x + e2.y;
#pragma unroll
      for (int r = kRows; r < 16; ++r) { ze[r] = make_float2(0.f, 0.f); zo[r] = make_float2(0.f, 0.f); }
    }
    __syncwarp();  // every lane is done with the span buffer

    // ---- prefetch the next item's span into the same buffer ----------------------------
    const Item me = cur;
    {
      const long long nxt = item + warp_stride;
      if (nxt < a.total_groups) {
        cur = decode_item(a, nxt);
        cur_tma = stage(cur);
      }
    }

    float log_e = 0.0f;
    if (a.use_energy) {
      esum = group_sum8(esum);
      log_e = logf(fmaxf(esum, 0.0f) + a.eps);
      log_e = fminf(fmaxf(log_e, a.energy_floor), 3.402823466e+38f);
    }

    // ---- stage 1: two 16-point FFTs over n1, twiddle W_256^(n2 k1), rows [k1][n2] of the tile ---------
#if KTF_R16_TW1_DERIVED
    // Twiddles: the table holds W_256^(2j k1) for the even column n2 = 2j; the odd column's W_256^((2j+1) k1) is that
    // times the compile-time constant W_256^k1 (two packed instructions instead of 8 more bytes of LDS per k1).
    float2 t1[16];   // issued ahead of the FFTs so that their latency is covered by arithmetic
#else
    float4 t1[16];   // issued ahead of the FFTs so that their latency is covered by arithmetic
#endif
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) t1[k1] = tw1[k1];
    fft16<true>(ze);
#pragma unroll
    for (int k1 = 8; k1 < 16; ++k1) t1[k1] = tw1[k1];
    fft16<true>(zo);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      float2 e = ze[pos16(k1)], o = zo[pos16(k1)];
      if (k1 > 0) {
#if KTF_R16_TW1_DERIVED
        const float2 te = t1[k1];
        const float2 to = cmul(te, make_float2(kCos256[k1], -kSin256[k1]));
#else
        const float2 te = make_float2(t1[k1].x, t1[k1].y), to = make_float2(t1[k1].z, t1[k1].w);
#endif
        e = cmul(e, te);
        o = cmul(o, to);
      }
      *reinterpret_cast<float4*>(tile_w + k1 * kRowStride) = make_float4(e.x, e.y, o.x, o.y);
    }
    __syncwarp();

    // ---- stage 2: columns a and b over n2 ---------------------------------------------------------------
    float2 va[16], vb[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 ra = *reinterpret_cast<const float4*>(row_a + 4 * c);
      const float4 rb = *reinterpret_cast<const float4*>(row_b + 4 * c);
      va[2 * c] = make_float2(ra.x, ra.y);
      va[2 * c + 1] = make_float2(ra.z, ra.w);
      vb[2 * c] = make_float2(rb.x, rb.y);
      vb[2 * c + 1] = make_float2(rb.z, rb.w);
    }
#if !KTF_R16_TW2_DERIVED
    float4 t2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) t2[e] = tw2[e];
#endif
    __syncwarp();  // the tile is consumed; it is reused as the power buffer below
    fft16<false>(va);
    fft16<false>(vb);

    // ---- real-FFT untangling + |X|^2 (x4; the 1/4 sits in the mel weights) ----------------------------------
    // evaluation e pairs zk = Z[k] with zp = Z[256 - k]:  4|X[k]|^2 = |S + G|^2, 4|X[256-k]|^2 = |S - G|^2,
    // S = zk + conj(zp), G = (-i W_512^k) (zk - conj(zp)).
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float2 zk, zp;
      if (e < 8) {
        zk = va[pos16(e)];
        const float2 p_self = va[pos16((16 - e) & 15)], p_reg = vb[pos16(15 - e)];
        zp = make_float2(j0 ? p_self.x : p_reg.x, j0 ? p_self.y : p_reg.y);
      } else {
        const float2 k_self = vb[pos16(e)], k_reg = va[pos16(e)];
        zk = make_float2(j0 ? k_self.x : k_reg.x, j0 ? k_self.y : k_reg.y);
        zp = vb[pos16(15 - e)];
      }
#if KTF_R16_TW2_DERIVED
      // t = -i W_512^k, k = j + 16 e (lane 0: 16 e, then 8 + 16 e): the lane's base twiddle times the constant -i W_32^e
      const float2 t = cmul(e < 8 ? make_float2(tw2.x, tw2.y) : make_float2(tw2.z, tw2.w),
                            make_float2(-kSin32[e], -kCos32[e]));
#else
      const float4 t4 = t2[e >> 1];
      const float2 t = (e & 1) ? make_float2(t4.z, t4.w) : make_float2(t4.x, t4.y);
#endif
      const float2 S = fma2(zp, make_float2(1.0f, -1.0f), zk);    // zk + conj(zp)
      const float2 D = fma2(zp, make_float2(-1.0f, 1.0f), zk);    // zk - conj(zp)
      const float2 G = cmul(D, t);
      const float2 u = cadd(S, G), v = csub(S, G);
      const float p1 = fmaf(u.x, u.x, u.y * u.y);
      float p2 = fmaf(v.x, v.x, v.y * v.y);
      float* dst1 = (e < 8 ? qa_lo : qa_hi) + 16 * e;
      float* dst2 = (e < 8 ? qb_lo : qb_hi) + 16 * (15 - e);
      if (e == 0) {   // lane 0: the partner of bin 0 is the Nyquist bin (unused); store bin 128 = conj(Z[128]) instead
        const float2 z8 = va[pos16(8)];
        const float p128 = 4.0f * fmaf(z8.x, z8.x, z8.y * z8.y);
        p2 = j0 ? p128 : p2;
        dst2 = j0 ? tile + 128 : dst2;
      }
      *dst1 = p1;
      *dst2 = p2;
    }
    __syncwarp();

    // ---- mel bank (filterbank.py:238-240): lane (mf, mj) walks the NU units (8 bins each) of schedule row mj on frame
    //      mf; the units of a filter are consecutive on one row, the running sum is flushed at the filter's last unit.
    //      Four units per trip: all 18 loads of a trip are in flight before the first product, instead of a
    //      descriptor -> address -> data chain per unit.
    float run = 0.0f, keep = 0.0f;
    {
      int u = 0;
      for (; u + 4 <= NU; u += 4) mel_trip<4>(u, udesc4, uwts, mel_tile, mel_acc, run, keep);
      switch (NU - u) {   // warp-uniform remainder, each size fully unrolled (predicating a 4-unit trip spills)
        case 3: mel_trip<3>(u, udesc4, uwts, mel_tile, mel_acc, run, keep); break;
        case 2: mel_trip<2>(u, udesc4, uwts, mel_tile, mel_acc, run, keep); break;
        case 1: mel_trip<1>(u, udesc4, uwts, mel_tile, mel_acc, run, keep); break;
        default: break;
      }
    }
    __syncwarp();
    if (OUTPUT == KTF_OUT_MFCC && DCT_REG == 2) {
      // log (filterbank.py:240) of the mirrored pairs (i, M-1-i), i = 2j, 2j+1; their sum and difference go to the
      // (now dead) power tile of the frame: [0, 16) = s, [16, 32) = d
      const int i0 = 2 * j, half = (M + 1) >> 1;
      float2 lo = *reinterpret_cast<const float2*>(lm_acc + i0);
      const int m0 = M - 1 - i0, m1 = M - 2 - i0;
      float2 hi = make_float2(lm_acc[m0 < 0 ? 0 : m0], lm_acc[m1 < 0 ? 0 : m1]);
      if (a.use_log) {
        lo.x = __logf(fmaxf(lo.x, 0.0f) + a.eps);
        lo.y = __logf(fmaxf(lo.y, 0.0f) + a.eps);
        hi.x = __logf(fmaxf(hi.x, 0.0f) + a.eps);
        hi.y = __logf(fmaxf(hi.y, 0.0f) + a.eps);
      }
      float2 sv = add2(lo, hi), dv = sub2(lo, hi);
      if (i0 == m0) { sv.x = lo.x; dv.x = 0.0f; }               // the self-paired middle bin of an odd M
      if (i0 + 1 == m1) { sv.y = lo.y; dv.y = 0.0f; }
      if (i0 >= half) { sv.x = 0.0f; dv.x = 0.0f; }
      if (i0 + 1 >= half) { sv.y = 0.0f; dv.y = 0.0f; }
      *reinterpret_cast<float2*>(tile + i0) = sv;
      *reinterpret_cast<float2*>(tile + 16 + i0) = dv;
      __syncwarp();
      float2 o2[4];   // even-i and odd-i partial sums
#pragma unroll
      for (int ff = 0; ff < 4; ++ff) o2[ff] = make_float2(0.0f, 0.0f);
      const float* vrow = s_T + 16 * (lane & 1);
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
#pragma unroll
        for (int ff = 0; ff < 4; ++ff) {
          const float4 v = *reinterpret_cast<const float4*>(vrow + ff * kTile + 4 * i4);
          o2[ff] = fma2(make_float2(v.x, v.y), make_float2(dreg[4 * i4], dreg[4 * i4 + 1]), o2[ff]);
          o2[ff] = fma2(make_float2(v.z, v.w), make_float2(dreg[4 * i4 + 2], dreg[4 * i4 + 3]), o2[ff]);
        }
      }
      // C0 <- log-energy (mfcc.py:219-228): frame ff's value lives in lanes 8 ff .. 8 ff + 7
      float o[4];
#pragma unroll
      for (int ff = 0; ff < 4; ++ff) {
        o[ff] = o2[ff].x + o2[ff].y;
        const float le = __shfl_sync(0xffffffffu, log_e, 8 * ff);
        if (lane == 0 && a.use_energy) o[ff] = le;
      }
      if (lane < a.Kc) {
        float* dst = a.out + (me.out_row0 + me.frame0) * (long long)a.Kc + lane;
#pragma unroll
        for (int ff = 0; ff < 4; ++ff)
          if (ff < me.nvalid) dst[ff * a.Kc] = o[ff];
      }
      continue;
    }

    // log (filterbank.py:240); each lane finishes mel bins 4j .. 4j+3 (+32, +64, ...) of its frame
    for (int i0 = 4 * j; i0 < M; i0 += 32) {
      float4 v = *reinterpret_cast<const float4*>(lm_acc + i0);
      // __logf (MUFU.LG2 * ln2): absolute error <= ~2e-6 on log-mel values of magnitude ~20, two orders below the
      // float32 noise of the spectrum itself; the frame log-energy (VAD input) keeps the exact logf
      if (a.use_log) {
        v.x = __logf(fmaxf(v.x, 0.0f) + a.eps);
        v.y = __logf(fmaxf(v.y, 0.0f) + a.eps);
        v.z = __logf(fmaxf(v.z, 0.0f) + a.eps);
        v.w = __logf(fmaxf(v.w, 0.0f) + a.eps);
      }
      if (OUTPUT == KTF_OUT_MFCC) {
        *reinterpret_cast<float4*>(lm_acc + i0) = v;
      } else {
        if (i0 < M) lm_row[i0] = v.x;
        if (i0 + 1 < M) lm_row[i0 + 1] = v.y;
        if (i0 + 2 < M) lm_row[i0 + 2] = v.z;
        if (i0 + 3 < M) lm_row[i0 + 3] = v.w;
      }
    }
    __syncwarp();

    if (OUTPUT == KTF_OUT_MFCC && DCT_REG == 1) {
      // ---- DCT (dct.py:176) with the lifter (mfcc.py:212) folded into the matrix: lane c = cepstrum c of all 4 frames
      float2 o2[4];   // even-i and odd-i partial sums
#pragma unroll
      for (int ff = 0; ff < 4; ++ff) o2[ff] = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        if (4 * i4 < M) {
#pragma unroll
          for (int ff = 0; ff < 4; ++ff) {
            const float4 lm = *reinterpret_cast<const float4*>(s_LM + ff * LMS + 4 * i4);   // broadcast
            o2[ff] = fma2(make_float2(lm.x, lm.y), make_float2(dreg[4 * i4], dreg[4 * i4 + 1]), o2[ff]);
            o2[ff] = fma2(make_float2(lm.z, lm.w), make_float2(dreg[4 * i4 + 2], dreg[4 * i4 + 3]), o2[ff]);
          }
        }
      }
      float o[4];
#pragma unroll
      for (int ff = 0; ff < 4; ++ff) o[ff] = o2[ff].x + o2[ff].y;
      // C0 <- log-energy (mfcc.py:219-228): frame ff's value lives in lanes 8 ff .. 8 ff + 7
#pragma unroll
      for (int ff = 0; ff < 4; ++ff) {
        const float le = __shfl_sync(0xffffffffu, log_e, 8 * ff);
        if (lane == 0 && a.use_energy) o[ff] = le;
      }
      if (lane < a.Kc) {
        float* dst = a.out + (me.out_row0 + me.frame0) * (long long)a.Kc + lane;
#pragma unroll
        for (int ff = 0; ff < 4; ++ff)
          if (ff < me.nvalid) dst[ff * a.Kc] = o[ff];
      }
      continue;
    }

    if (OUTPUT == KTF_OUT_MFCC) {
      // ---- DCT from the shared-memory table (more than 32 mel bins): lane (f, j) = cepstra 4j .. 4j+3 of frame f
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const float* dcol = s_dct + 4 * j;
      const int M4 = M & ~3;
      for (int i = 0; i < M4; i += 4) {
        const float4 lm = *reinterpret_cast<const float4*>(lm_row + i);
        const float lmv[4] = {lm.x, lm.y, lm.z, lm.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 d = *reinterpret_cast<const float4*>(dcol + (i + u) * 32);
          acc[0] = fmaf(lmv[u], d.x, acc[0]);
          acc[1] = fmaf(lmv[u], d.y, acc[1]);
          acc[2] = fmaf(lmv[u], d.z, acc[2]);
          acc[3] = fmaf(lmv[u], d.w, acc[3]);
        }
      }
      for (int i = M4; i < M; ++i) {
        const float lm = lm_row[i];
        const float4 d = *reinterpret_cast<const float4*>(dcol + i * 32);
        acc[0] = fmaf(lm, d.x, acc[0]);
        acc[1] = fmaf(lm, d.y, acc[1]);
        acc[2] = fmaf(lm, d.z, acc[2]);
        acc[3] = fmaf(lm, d.w, acc[3]);
      }
      if (j0 && a.use_energy) acc[0] = log_e;
      float* orow = s_out + f * a.Kc + 4 * j;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (4 * j + r < a.Kc) orow[r] = acc[r];
      __syncwarp();
    }

    // ---- coalesced store of the group's nvalid x out_row tile ----------------------------------------------
    {
      float* dst = a.out + (me.out_row0 + me.frame0) * (long long)out_row;
      const int n = me.nvalid * out_row;
      if (((n & 3) == 0) && ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0)) {
        const int n4 = n >> 2;
        for (int i = lane; i < n4; i += 32)
          reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_out)[i];
      } else {
        for (int i = lane; i < n; i += 32) dst[i] = s_out[i];
      }
    }
  }
}

__host__ int r16_lms(const ktf_frontend_cfg& c) {
  return (c.output == KTF_OUT_MFCC && c.num_mels <= 32) ? 36 : ((c.num_mels + 3) & ~3) + 4;
}

size_t r16_smem_bytes(const ktf_frontend* fe) {
  const int span_p = ((fe->span + 3) & ~3) + 16;
  const int LMS = r16_lms(fe->cfg);
  const int out_row = fe->out_dim;
  const int out_sz = (4 * out_row + 3) & ~3;
  const size_t warp_floats = ((size_t)span_p + 4 * kTile + 4 * LMS + out_sz + 2 + 31) & ~(size_t)31;
  return ((size_t)fe->r16_blob_floats + kR16Warps * warp_floats) * sizeof(float);
}

template <int OUTPUT, bool RAW, int DCT_REG, bool PCM16>
int launch_r16(const ktf_frontend* fe, FrontendArgs& a, cudaStream_t st) {
  const size_t smem = r16_smem_bytes(fe);
  auto kern = frontend_r16_kernel<OUTPUT, RAW, DCT_REG, PCM16>;
  KTF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)std::max<size_t>(smem, 48 * 1024)));
  int occ = 0;
  KTF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kR16Threads, smem));
  if (occ < 1) occ = 1;
  const long long ctas_needed = (a.total_groups + kR16Warps - 1) / kR16Warps;
  const long long grid = std::min<long long>(ctas_needed, (long long)ktf::num_sms() * occ);
  if (grid <= 0) return KTF_OK;
  kern<<<(unsigned)grid, kR16Threads, smem, st>>>(a);
  KTF_LAUNCH_OK();
  return KTF_OK;
}

}  // namespace

namespace ktf_fe {

int r16_build(ktf_frontend* fe, const float* window_host, const float* mel_bank_host, const float* dct_host,
              const float* lifter_host) {
  const ktf_frontend_cfg& c = fe->cfg;
  const int M = c.num_mels, Kc = c.num_ceps, C = 256;
  const bool shape_ok = c.fft_length == 512 && c.frame_width == kW && (c.frame_shift % 4) == 0 &&
                        c.frame_shift >= 4 && c.frame_shift <= 1024;
  const bool kind_ok = (c.output == KTF_OUT_MFCC || c.output == KTF_OUT_FBANK) && c.use_power != 0;
  if (!shape_ok || !kind_ok || mel_bank_host == nullptr || M < 1 || M > kMaxMels) return KTF_OK;
  if (c.output == KTF_OUT_MFCC && (dct_host == nullptr || Kc < 1 || Kc > kMaxCeps)) return KTF_OK;
  for (int i = 0; i < M; ++i)   // the fast path never computes the Nyquist bin
    if (mel_bank_host[(size_t)C * M + i] != 0.0f) return KTF_OK;

  // ---- mel filters: chunk ranges (4 bins), cut into units of 2 chunks, then a longest-first assignment of whole
  //      filters to the 8 lanes of a frame ---------------------------------------------------------------------
  struct Filt { int c0, n; };
  std::vector<Filt> filt((size_t)M);
  for (int i = 0; i < M; ++i) {
    int lo = -1, hi = -1;
    for (int k = 0; k < C; ++k)
      if (mel_bank_host[(size_t)k * M + i] != 0.0f) { if (lo < 0) lo = k; hi = k; }
    int c0 = 0, n = 0;
    if (lo >= 0) { c0 = lo >> 2; n = (hi >> 2) - c0 + 1; }
    filt[i] = {c0, n};
  }
  // Rows of even index take filters whose first chunk is even, odd rows odd ones (units advance by two chunks, so the
  // parity holds for every unit of the filter): together with the kernel's lane mapping this makes every LDS.128 of
  // the mel phase bank-conflict free.  A filter with an odd number of chunks may start one (all-zero) chunk early at
  // no cost in units, which lets it go to either parity; filters go longest-first to the least loaded allowed row.
  std::vector<int> order((size_t)M);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return filt[x].n > filt[y].n; });
  std::vector<std::vector<int>> lanes(8);
  int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i : order) {
    if (filt[i].n == 0) { filt[i].c0 = 0; }
    const bool flexible = (filt[i].n & 1) && filt[i].c0 > 0;
    int best = -1;
    for (int l = 0; l < 8; ++l) {
      if (!flexible && filt[i].n > 0 && (l & 1) != (filt[i].c0 & 1)) continue;
      if (best < 0 || load[l] < load[best]) best = l;
    }
    if (filt[i].n > 0 && (best & 1) != (filt[i].c0 & 1)) { filt[i].c0 -= 1; filt[i].n += 1; }
    lanes[best].push_back(i);
    load[best] += (filt[i].n + 1) / 2;
  }
  int NU = 1;
  for (int l = 0; l < 8; ++l) NU = std::max(NU, load[l]);

  // descriptors are read four at a time (the last trip may be partial: pad them to a multiple of four)
  const int SD = pad4mod32(2 * ((NU + 3) & ~3)), SW = pad4mod32(8 * NU);
  const int LMS = r16_lms(c);
  const float one = 1.0f;
  int one_bits;
  memcpy(&one_bits, &one, sizeof(one_bits));
  const int dct_floats = c.output == KTF_OUT_MFCC ? M * 32 : 0;
  const int blob_floats = (kOffUnits + 8 * SD + 8 * SW + dct_floats + 31) & ~31;
  std::vector<float> blob((size_t)blob_floats, 0.0f);
  for (int i = 0; i < kW; ++i) blob[kOffWin + i] = window_host[i];
  for (int l = 0; l < 8; ++l) {
    int* ud = reinterpret_cast<int*>(blob.data() + kOffUnits + l * SD);
    float* uw = blob.data() + kOffUnits + 8 * SD + l * SW;
    int u = 0;
    for (int i : lanes[l])
      for (int cc = 0; cc < filt[i].n; cc += 2, ++u) {
        ud[2 * u] = (4 * (filt[i].c0 + cc)) | (i << 16);
        ud[2 * u + 1] = (cc + 2 >= filt[i].n) ? 0 : one_bits;   // running sum restarts after the filter's last unit
        for (int b8 = 0; b8 < 8; ++b8) {
          const int k = 4 * (filt[i].c0 + cc) + b8;
          const bool in = (cc + b8 / 4) < filt[i].n && k < C;
          uw[8 * u + b8] = in ? mel_bank_host[(size_t)k * M + i] * 0.25f : 0.0f;   // the kernel stores 4|X|^2
        }
      }
    // padding: zero weights, spare slot, chunk l (the row's own parity)
    for (; u < ((NU + 3) & ~3); ++u) { ud[2 * u] = (4 * l) | ((LMS - 1) << 16); ud[2 * u + 1] = 0; }
  }

  const double PI = 3.14159265358979323846;
#if KTF_R16_TW1_DERIVED
  for (int j = 0; j < 8; ++j)
    for (int k1 = 0; k1 < 16; ++k1) {
      const double th = -2.0 * PI * (double)((2 * j * k1) % 256) / 256.0;
      blob[kOffTw1 + j * kTw1Stride + k1 * 2] = (float)cos(th);
      blob[kOffTw1 + j * kTw1Stride + k1 * 2 + 1] = (float)sin(th);
    }
#else
  for (int j = 0; j < 8; ++j)
    for (int k1 = 0; k1 < 16; ++k1)
      for (int h = 0; h < 2; ++h) {
        const int n2 = 2 * j + h;
        const double th = -2.0 * PI * (double)((n2 * k1) % 256) / 256.0;
        blob[kOffTw1 + j * kTw1Stride + k1 * 4 + 2 * h] = (float)cos(th);
        blob[kOffTw1 + j * kTw1Stride + k1 * 4 + 2 * h + 1] = (float)sin(th);
      }
#endif
#if KTF_R16_TW2_DERIVED
  for (int j = 0; j < 8; ++j) {   // W_512^k of the lane's first column for e < 8 and e >= 8
    const int k_lo = j, k_hi = (j == 0) ? 8 : j;
    blob[kOffTw2 + j * kTw2Stride + 0] = (float)cos(2.0 * PI * k_lo / 512.0);
    blob[kOffTw2 + j * kTw2Stride + 1] = (float)(-sin(2.0 * PI * k_lo / 512.0));
    blob[kOffTw2 + j * kTw2Stride + 2] = (float)cos(2.0 * PI * k_hi / 512.0);
    blob[kOffTw2 + j * kTw2Stride + 3] = (float)(-sin(2.0 * PI * k_hi / 512.0));
  }
#else
  for (int j = 0; j < 8; ++j)
    for (int e = 0; e < 16; ++e) {
      int k = j + 16 * e;
      if (j == 0) k = e < 8 ? 16 * e : 8 + 16 * e;
      const double th = 2.0 * PI * (double)k / 512.0;   // -i * exp(-i th) = (-sin th, -cos th)
      blob[kOffTw2 + j * kTw2Stride + e * 2] = (float)(-sin(th));
      blob[kOffTw2 + j * kTw2Stride + e * 2 + 1] = (float)(-cos(th));
    }
#endif
  fe->r16_dct_sym = 0;
  if (c.output == KTF_OUT_MFCC) {
    float* dp = blob.data() + kOffUnits + 8 * SD + 8 * SW;
    for (int i = 0; i < M; ++i)
      for (int cc = 0; cc < Kc; ++cc) {
        const float lf = (c.apply_lifter && lifter_host) ? lifter_host[cc] : 1.0f;
        dp[(size_t)i * 32 + cc] = dct_host[(size_t)i * Kc + cc] * lf;
      }
    // DCT-II mirror symmetry (dct.py:98-143 builds cos(pi/M (i + 1/2) c), so row M-1-i is (-1)^c times row i; the
    // two are rounded from float64 separately, hence the 1-ulp tolerance): enables the half-size register DCT
    float dmax = 0.0f;
    for (int i = 0; i < M * Kc; ++i) dmax = std::max(dmax, fabsf(dct_host[i]));
    bool sym = M <= 32;
    for (int i = 0; i < M && sym; ++i)
      for (int cc = 0; cc < Kc && sym; ++cc) {
        const float sgn = (cc & 1) ? -1.0f : 1.0f;
        sym = fabsf(dct_host[(size_t)(M - 1 - i) * Kc + cc] - sgn * dct_host[(size_t)i * Kc + cc]) <= 2.5e-7f * dmax;
      }
    fe->r16_dct_sym = sym ? 1 : 0;
  }

  fe->r16_blob_floats = blob_floats;
  fe->r16_nf = NU;
  fe->r16_melw_floats = 8 * (SD + SW);
  if (r16_smem_bytes(fe) > 227 * 1024) return KTF_OK;   // generic kernel instead
  return ktf::upload(&fe->d_r16, blob.data(), blob.size());
}

template <int OUTPUT, int DCT_REG>
int launch_r16_sel(const ktf_frontend* fe, FrontendArgs& a, cudaStream_t st) {
  const bool raw = fe->cfg.raw_energy != 0, pcm = a.wav16 != nullptr;
  if (raw) return pcm ? launch_r16<OUTPUT, true, DCT_REG, true>(fe, a, st) : launch_r16<OUTPUT, true, DCT_REG, false>(fe, a, st);
  return pcm ? launch_r16<OUTPUT, false, DCT_REG, true>(fe, a, st) : launch_r16<OUTPUT, false, DCT_REG, false>(fe, a, st);
}

int r16_launch(const ktf_frontend* fe, FrontendArgs& a, cudaStream_t st) {
  if (fe->cfg.output == KTF_OUT_MFCC) {
    if (fe->cfg.num_mels <= 32) {
      if (fe->r16_dct_sym) return launch_r16_sel<KTF_OUT_MFCC, 2>(fe, a, st);
      return launch_r16_sel<KTF_OUT_MFCC, 1>(fe, a, st);
    }
    return launch_r16_sel<KTF_OUT_MFCC, 0>(fe, a, st);
  }
  return launch_r16_sel<KTF_OUT_FBANK, 0>(fe, a, st);
}

}  // namespace ktf_fe
