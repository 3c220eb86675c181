// k_idct.cu -- decoder block path: ParseBlock + DecodeBlock of
// the reference decoder (libs/decoder.cpp:102-149) over the record loop of
// Decoder::operator() (:191-213), i.e. the inverse of K3.
//
// Per 772-byte record: quantisation step q = 1 inside the gaze rectangle, bg_q for
// BLOCK_TYPE_BACKGROUND, fg_q otherwise; c = round(c / q) * q in float (exactly the
// reference's three float statements); inverse orthonormal 8x8 DCT (cv::idct) per
// channel; cv::merge into the interleaved float BGR frame.
//
// Mirror image of dct8x8_stream_kernel: a CTA = 3 warps owns 32 consecutive records
// of one block row; the contiguous 24.7 KB record span is pulled into shared memory
// with one TMA bulk copy (cp.async.bulk global -> shared on an mbarrier); warp c
// dequantises and inverse-transforms channel c of all 32 blocks (lane = block,
// bank-conflict free at the 193-word record stride); the pixels are exchanged through
// shared memory so that every output row of the tile leaves as contiguous 128-bit
// stores (32 blocks x 8 px x 3 channels = 3 KB per row).  HBM-bound: 772 B in,
// 768 B out per block.  idct16x16_decode_kernel / idct4x4_decode_kernel (further down) are the
// mirror images of dct16x16_stream_kernel / dct4x4_stream_kernel for 3076-byte and 196-byte records.
#include "common.cuh"

namespace svc {

// ---- dequantiser: round(c / q) * q, libs/decoder.cpp:140-142 ----------------------------------------
// The reference divides a float coefficient by the (integer) step in IEEE single precision, rounds half
// away from zero and multiplies back.  __fdiv_rn is a ~15-instruction subroutine per coefficient, as
// expensive as the inverse transform itself.  For steps up to 4096 and |c| <= 2^18 (a 16x16 block of
// 255s has a DC coefficient of 4080) the correctly rounded quotient is instead obtained with Markstein's
// sequence from the correctly rounded reciprocal rq = RN(1/q):
//     y = RN(c * rq);  e = RN(c - q * y) (one FMA);  d = RN(y + e * rq) (one FMA)
// svc_selftest_dequant() compares it with __fdiv_rn for EVERY float in [0, 2^18] and every step of a
// range (tests/test_decode.py runs q = 1 .. 4096: zero mismatches); anything outside that domain --
// larger steps, huge / non-finite coefficients of a corrupt stream -- takes the IEEE division.
constexpr float kDequantFastMaxAbs = 262144.0f;
constexpr uint32_t kDequantFastMaxStep = 4096u;

struct QuantStep {
  float q, rq;
  bool fast;
};

__device__ __forceinline__ QuantStep make_quant_step(const uint32_t step) {
  QuantStep s;
  s.q = (float)step;
  s.rq = __frcp_rn(s.q);
  s.fast = step <= kDequantFastMaxStep;
  return s;
}

__device__ __forceinline__ float quotient_rn(const float c, const QuantStep& s) {
  if (s.fast && fabsf(c) <= kDequantFastMaxAbs) {
    const float y = __fmul_rn(c, s.rq);
    const float e = __fmaf_rn(-s.q, y, c);
    return __fmaf_rn(e, s.rq, y);
  }
  return __fdiv_rn(c, s.q);
}

// round(c / q) * q.  The FMA sequence may differ from the division where the quotient underflows
// (|c / q| < 2^-126: double rounding in the subnormal range, and -0 comes out as +0); every such quotient
// rounds to zero, so only the sign of that zero needs restoring: the result always has the sign of c.
__device__ __forceinline__ float dequant_value(const float c, const QuantStep& s) {
  return copysignf(__fmul_rn(roundf(quotient_rn(c, s)), s.q), c);
}

__device__ __forceinline__ float dequant(const uint32_t w, const QuantStep& s) {
  return dequant_value(__uint_as_float(w), s);
}

// N coefficients at once: ONE (almost never taken) branch for the whole group instead of a
// divergence region per coefficient -- the straight path is 3 FMA-pipe instructions + the rounding.
template <int N>
__device__ __forceinline__ void dequant_group(const uint32_t (&w)[N], const QuantStep& s, float (&out)[N]) {
  // |c| <= 2^18 for the whole group <=> the largest magnitude bit pattern is at most that of 2^18
  // (infinities and NaNs have larger patterns than any finite float)
  uint32_t big = 0u;
#pragma unroll
  for (int i = 0; i < N; ++i) big = max(big, w[i] & 0x7fffffffu);
  const bool all_finite_small = s.fast && big <= 0x48800000u;
  if (all_finite_small) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float c = __uint_as_float(w[i]);
      const float y = __fmul_rn(c, s.rq);
      const float e = __fmaf_rn(-s.q, y, c);
      out[i] = copysignf(__fmul_rn(roundf(__fmaf_rn(e, s.rq, y)), s.q), c);
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < N; ++i) out[i] = dequant(w[i], s);
  }
}

// every float bit pattern in [-2^18, 2^18] x every step in [q_lo, q_hi]: dequantised value with the fast
// quotient vs with the IEEE division, bit for bit; mismatches[1] keeps one offending (c, step) pair
__global__ void __launch_bounds__(256) dequant_selftest_kernel(const uint32_t q_lo, const uint32_t q_hi,
                                                               unsigned long long* mismatches) {
  constexpr uint32_t kMaxBits = 0x48800000u;  // 2^18
  unsigned long long bad = 0;
  for (uint32_t step = q_lo + blockIdx.y; step <= q_hi; step += gridDim.y) {
    const QuantStep s = make_quant_step(step);
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= kMaxBits; b += (uint64_t)gridDim.x * blockDim.x) {
      const float c = __uint_as_float((uint32_t)b);
      const float ref = __fmul_rn(roundf(__fdiv_rn(c, s.q)), s.q);  // libs/decoder.cpp:140-142
      const float got = dequant_value(c, s);
      const float refn = __fmul_rn(roundf(__fdiv_rn(-c, s.q)), s.q);
      const float gotn = dequant_value(-c, s);
      const bool m = (__float_as_uint(ref) != __float_as_uint(got)) || (__float_as_uint(refn) != __float_as_uint(gotn));
      if (m) {
        ++bad;
        mismatches[1] = ((unsigned long long)step << 32) | (uint32_t)b;
      }
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

cudaError_t run_dequant_selftest(uint32_t q_lo, uint32_t q_hi, unsigned long long* d_mismatches, cudaStream_t st) {
  const uint32_t ny = q_hi - q_lo + 1u < 64u ? q_hi - q_lo + 1u : 64u;
  dequant_selftest_kernel<<<dim3(kNumSms * 8, ny), 256, 0, st>>>(q_lo, q_hi, d_mismatches);
  return cudaGetLastError();
}


#define SVC_C4 0.35355339059327376220f
#define SVC_A  0.49039264020161522456f
#define SVC_B2 0.46193976625564337806f
#define SVC_B  0.41573480615127261854f
#define SVC_C  0.27778511650980111237f
#define SVC_B6 0.19134171618254488586f
#define SVC_D  0.09754516100806413392f

// sqrt(1/8) cos(k pi / 16), k = 1..7 and 1/4: the 8-point factors times 1/sqrt(2), rounded once (see k_dct.cu)
#define SVC_H4 0.25f
#define SVC_HA  0.3467599613305369f
#define SVC_HB2 0.32664074121909414f
#define SVC_HB  0.2939689006048397f
#define SVC_HC  0.1964237395967756f
#define SVC_HB6 0.13529902503654928f
#define SVC_HD  0.06897484482073578f
// sqrt(1/2) cos(k pi / 16): the same factors times sqrt(2)
#define SVC_R4 0.5f
#define SVC_RA  0.6935199226610738f
#define SVC_RB2 0.6532814824381883f
#define SVC_RB  0.5879378012096794f
#define SVC_RC  0.3928474791935512f
#define SVC_RB6 0.27059805007309856f
#define SVC_RD  0.13794968964147156f

// x = C^T X for the orthonormal 8-point DCT-II matrix C (even/odd split, 34 flops).  The 2-D inverse
// applies one pass with every factor times sqrt(2) (kUp) and the other with every factor times
// 1/sqrt(2): the product is unchanged and both DC factors (1/2, 1/4) are exact, as in k_dct.cu.
template <bool kUp>
__device__ __forceinline__ void idct8(float& x0, float& x1, float& x2, float& x3,
                                      float& x4, float& x5, float& x6, float& x7) {
  constexpr float c4 = kUp ? SVC_R4 : SVC_H4, ca = kUp ? SVC_RA : SVC_HA, cb2 = kUp ? SVC_RB2 : SVC_HB2,
                  cb = kUp ? SVC_RB : SVC_HB, cc = kUp ? SVC_RC : SVC_HC, cb6 = kUp ? SVC_RB6 : SVC_HB6,
                  cd = kUp ? SVC_RD : SVC_HD;
  const float p = c4 * (x0 + x4), q = c4 * (x0 - x4);
  const float r = fmaf(cb2, x2, cb6 * x6), s = fmaf(cb6, x2, -cb2 * x6);
  const float e0 = p + r, e3 = p - r, e1 = q + s, e2 = q - s;
  const float o0 = fmaf(ca, x1, fmaf(cb, x3, fmaf(cc, x5, cd * x7)));
  const float o1 = fmaf(cb, x1, fmaf(-cd, x3, fmaf(-ca, x5, -cc * x7)));
  const float o2 = fmaf(cc, x1, fmaf(-ca, x3, fmaf(cd, x5, cb * x7)));
  const float o3 = fmaf(cd, x1, fmaf(-cc, x3, fmaf(cb, x5, -ca * x7)));
  x0 = e0 + o0; x7 = e0 - o0;
  x1 = e1 + o1; x6 = e1 - o1;
  x2 = e2 + o2; x5 = e2 - o2;
  x3 = e3 + o3; x4 = e3 - o3;
}

constexpr int kRecW = 193;              // words per record
constexpr int kOutBlk = 25;             // words per block per output row in smem (24 + 1 pad)
constexpr int kOutRow = 32 * kOutBlk;   // 800 words per output row

__global__ void __launch_bounds__(96)
idct8x8_decode_kernel(const DecodeParams p) {
  __shared__ __align__(128) uint32_t stage[8 * kOutRow];  // 25.6 KB >= 32 records (24.7 KB)
  __shared__ __align__(8) uint64_t bar;
  const uint32_t lane = threadIdx.x & 31u, c = threadIdx.x >> 5;
  const uint32_t nbx = p.pw / 8u, nby = p.ph / 8u;
  const uint32_t chunks_x = (nbx + 31u) / 32u;
  uint32_t u = blockIdx.x;
  const uint32_t cxi = u % chunks_x; u /= chunks_x;
  const uint32_t by = u % nby;
  const uint32_t f = u / nby;
  const uint32_t bx0 = cxi * 32u;
  const uint32_t n_act = min(32u, nbx - bx0);
  const uint32_t bytes = n_act * kRecW * 4u;
  const uint8_t* src = p.records + (uint64_t)f * p.frame_record_bytes +
                       ((uint64_t)by * nbx + bx0) * (kRecW * 4u);

  // ---- records -> shared memory --------------------------------------------------
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((bytes & 15u) == 0);
  if (bulk_ok) {
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes)
                   : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"((uint32_t)__cvta_generic_to_shared(stage)), "l"(src), "r"(bytes), "r"(bar_addr)
          : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
    }
  } else {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    for (uint32_t k = threadIdx.x; k < bytes / 4u; k += 96u) stage[k] = __ldg(s32 + k);
    __syncthreads();
  }

  // ---- dequantise + inverse transform channel c of block `lane` -------------------
  float v[8][8];
  const bool active = lane < n_act;
  if (active) {
    const uint32_t* rec = stage + lane * kRecW;
    const uint32_t type = rec[0];
    const uint32_t x0 = (bx0 + lane) * 8u, y0 = by * 8u;
    const bool gazed = p.has_gaze && x0 >= p.gaze_x && x0 < p.gaze_x + p.gaze_w && y0 >= p.gaze_y &&
                       y0 < p.gaze_y + p.gaze_h;
    // (this kernel keeps the IEEE division: with 64 coefficients per lane in registers the division-free
    // forms measured slower here -- 4K microbench 84 % of the copy peak with __fdiv_rn, 78-83 % with the
    // per-coefficient fast quotient, 68 % with the grouped one; the 4x4 and 16x16 kernels gain from it)
    const float q = (float)(gazed ? 1u : (type == 0u ? p.bg_q : p.fg_q));
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float cf = __uint_as_float(rec[1 + c * 64 + r * 8 + j]);
        v[r][j] = __fmul_rn(roundf(__fdiv_rn(cf, q)), q);  // libs/decoder.cpp:140-142
      }
#pragma unroll
    for (int r = 0; r < 8; ++r) idct8<true>(v[r][0], v[r][1], v[r][2], v[r][3], v[r][4], v[r][5], v[r][6], v[r][7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) idct8<false>(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
  }
  __syncthreads();  // every warp has consumed the records: the buffer becomes the pixel tile
  if (active) {
    // [row][block * 25 + px * 3 + channel]: bank = (25 lane + const) % 32, conflict free
    uint32_t* o = stage + lane * kOutBlk + c;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) o[r * kOutRow + j * 3] = __float_as_uint(v[r][j]);
  }
  __syncthreads();
  // ---- tile rows -> global, 128-bit stores (6 per block per row) --------------------
  float* out = p.out + ((uint64_t)f * p.ph + (uint64_t)by * 8u) * p.pw * 3u + (uint64_t)bx0 * 24u;
  const uint32_t vec_per_row = n_act * 6u;
  for (uint32_t k = threadIdx.x; k < 8u * vec_per_row; k += 96u) {
    const uint32_t r = k / vec_per_row, i = k - r * vec_per_row;
    const uint32_t b = i / 6u, q4 = i - b * 6u;
    const uint32_t* s = stage + r * kOutRow + b * kOutBlk + q4 * 4u;
    const float4 val = make_float4(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]),
                                   __uint_as_float(s[3]));
    *reinterpret_cast<float4*>(out + (uint64_t)r * p.pw * 3u + (uint64_t)i * 4u) = val;
  }
}

// ---- 4x4 and 16x16 transform blocks (streams of the fused dct4x4 / dct16x16 kernels) ----------

// x = C^T X for the orthonormal 4-point DCT-II
__device__ __forceinline__ void idct4(float& x0, float& x1, float& x2, float& x3) {
  constexpr float a = 0.6532814824381883f, b = 0.27059805007309856f;  // sqrt(1/2) cos(pi/8), cos(3 pi/8)
  const float p = 0.5f * (x0 + x2), q = 0.5f * (x0 - x2);
  const float o0 = fmaf(a, x1, b * x3), o1 = fmaf(b, x1, -a * x3);
  x0 = p + o0; x3 = p - o0;
  x1 = q + o1; x2 = q - o1;
}

__device__ __forceinline__ bool in_gaze(const DecodeParams& p, const uint32_t x0, const uint32_t y0) {
  return p.has_gaze && x0 >= p.gaze_x && x0 < p.gaze_x + p.gaze_w && y0 >= p.gaze_y && y0 < p.gaze_y + p.gaze_h;
}

// records -> shared memory: one TMA bulk copy per part on `bar` (phase 0), or a cooperative copy
// when the span is not 16-byte aligned.  All threads call it and return with the data visible.
__device__ __forceinline__ void fetch_records(uint32_t* dst0, const uint8_t* src, const uint32_t bytes0,
                                              uint32_t* dst1, const uint32_t bytes1, uint64_t* bar) {
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (((bytes0 | bytes1) & 15u) == 0);
  if (bulk_ok) {
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes0 + bytes1)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(dst0)), "l"(src), "r"(bytes0), "r"(bar_addr) : "memory");
      if (bytes1)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(dst1)), "l"(src + bytes0), "r"(bytes1), "r"(bar_addr)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
    }
  } else {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    for (uint32_t k = threadIdx.x; k < bytes0 / 4u; k += blockDim.x) dst0[k] = __ldg(s32 + k);
    for (uint32_t k = threadIdx.x; k < bytes1 / 4u; k += blockDim.x) dst1[k] = __ldg(s32 + bytes0 / 4u + k);
    __syncthreads();
  }
}

constexpr int kRecW4 = 49;               // words per 4x4 record
constexpr int kUnitD4 = 128;             // records (lanes) per CTA
constexpr int kOutBlk4 = 13;             // words per block per output row in smem (12 + 1 pad)
constexpr int kOutRow4 = kUnitD4 * kOutBlk4;

// One lane = one 4x4 block, three channels; the record buffer becomes the pixel tile.
__global__ void __launch_bounds__(kUnitD4)
idct4x4_decode_kernel(const DecodeParams p) {
  __shared__ __align__(128) uint32_t stage[4 * kOutRow4];  // 26.6 KB >= 128 records (25.1 KB)
  __shared__ __align__(8) uint64_t bar;
  const uint32_t t = threadIdx.x;
  const uint32_t nbx = p.pw / 4u, nby = p.ph / 4u;
  const uint32_t chunks_x = (nbx + kUnitD4 - 1u) / kUnitD4;
  uint32_t u = blockIdx.x;
  const uint32_t cxi = u % chunks_x; u /= chunks_x;
  const uint32_t by = u % nby, f = u / nby;
  const uint32_t bx0 = cxi * kUnitD4, n_act = min((uint32_t)kUnitD4, nbx - bx0);
  const uint8_t* src = p.records + (uint64_t)f * p.frame_record_bytes + ((uint64_t)by * nbx + bx0) * (kRecW4 * 4u);
  fetch_records(stage, src, n_act * kRecW4 * 4u, stage, 0u, &bar);
  float v[3][4][4];
  const bool active = t < n_act;
  if (active) {
    const uint32_t* rec = stage + t * kRecW4;  // 49 lane + const: conflict free
    const QuantStep q = make_quant_step(in_gaze(p, (bx0 + t) * 4u, by * 4u) ? 1u : (rec[0] == 0u ? p.bg_q : p.fg_q));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[c][r][j] = dequant(rec[1 + c * 16 + r * 4 + j], q);
        idct4(v[c][r][0], v[c][r][1], v[c][r][2], v[c][r][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) idct4(v[c][0][j], v[c][1][j], v[c][2][j], v[c][3][j]);
    }
  }
  __syncthreads();  // every lane has consumed its record
  if (active) {
    uint32_t* o = stage + t * kOutBlk4;  // [row][block * 13 + px * 3 + channel]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) o[r * kOutRow4 + j * 3 + c] = __float_as_uint(v[c][r][j]);
  }
  __syncthreads();
  float* out = p.out + ((uint64_t)f * p.ph + (uint64_t)by * 4u) * p.pw * 3u + (uint64_t)bx0 * 12u;
  const uint32_t vec_per_row = n_act * 3u;
  for (uint32_t k = t; k < 4u * vec_per_row; k += kUnitD4) {
    const uint32_t r = k / vec_per_row, i = k - r * vec_per_row;
    const uint32_t b = i / 3u, q4 = i - b * 3u;
    const uint32_t* s = stage + r * kOutRow4 + b * kOutBlk4 + q4 * 4u;
    *reinterpret_cast<float4*>(out + (uint64_t)r * p.pw * 3u + (uint64_t)i * 4u) =
        make_float4(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3]));
  }
}

// 16-point inverse: x_i = e_i + o_i, x_{15-i} = e_i - o_i with e = IDCT8(X_even) / sqrt 2 and
// o_i = sum_m sqrt(1/8) cos(pi (2i+1)(2m+1) / 32) X[2m+1] (transpose of the forward split of k_dct.cu).
__device__ __forceinline__ constexpr float dct16_odd_factor(int i, int m) {
  constexpr float K[8] = {0.35185093438159565f, 0.33832950029358816f, 0.31180625324666783f,
                          0.2733004667504394f,  0.2242918965856591f,  0.1666639146194367f,
                          0.10263113188058934f, 0.034654292299772925f};
  int q = ((2 * i + 1) * (2 * m + 1)) % 64;
  if (q > 32) q = 64 - q;
  const bool neg = q > 16;
  if (neg) q = 32 - q;
  return neg ? -K[(q - 1) / 2] : K[(q - 1) / 2];
}

__device__ __forceinline__ void idct16(const float (&X)[16], float (&x)[16]) {
  const float p = SVC_H4 * (X[0] + X[8]), q = SVC_H4 * (X[0] - X[8]);
  const float r = fmaf(SVC_HB2, X[4], SVC_HB6 * X[12]), s = fmaf(SVC_HB6, X[4], -SVC_HB2 * X[12]);
  const float e0 = p + r, e3 = p - r, e1 = q + s, e2 = q - s;
  const float o0 = fmaf(SVC_HA, X[2], fmaf(SVC_HB, X[6], fmaf(SVC_HC, X[10], SVC_HD * X[14])));
  const float o1 = fmaf(SVC_HB, X[2], fmaf(-SVC_HD, X[6], fmaf(-SVC_HA, X[10], -SVC_HC * X[14])));
  const float o2 = fmaf(SVC_HC, X[2], fmaf(-SVC_HA, X[6], fmaf(SVC_HD, X[10], SVC_HB * X[14])));
  const float o3 = fmaf(SVC_HD, X[2], fmaf(-SVC_HC, X[6], fmaf(SVC_HB, X[10], -SVC_HA * X[14])));
  const float e[8] = {e0 + o0, e1 + o1, e2 + o2, e3 + o3, e3 - o3, e2 - o2, e1 - o1, e0 - o0};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float o = dct16_odd_factor(i, 0) * X[1];
#pragma unroll
    for (int m = 1; m < 8; ++m) o = fmaf(dct16_odd_factor(i, m), X[2 * m + 1], o);
    x[i] = e[i] + o;
    x[15 - i] = e[i] - o;
  }
}

constexpr int kRecW16 = 769;                     // words per 16x16 record
constexpr int kUnitD16 = 8;                      // records per CTA
constexpr int kStageHalfD16 = 4 * kRecW16 + 12;  // blocks w and w + 4 of a warp sit 16 banks apart
constexpr int kTmpPitchD16 = 20, kTmpBlkD16 = 16 * kTmpPitchD16 + 4;
constexpr int kPixPitchD16 = 49;                 // pixel tile: 48 floats per block row + 1
constexpr int kStageD16 = kUnitD16 * 16 * kPixPitchD16 + 16;  // >= 2 * kStageHalfD16

// Mirror image of dct16x16_stream_kernel: 128 threads own 8 consecutive records of a block row.
// Per channel: thread (block, column) dequantises and inverse-transforms its column into a
// transposed scratch tile, barrier, thread (block, row) inverse-transforms its row; the three
// channels of a row are interleaved in registers, exchanged through the (then dead) record buffer
// and leave as contiguous 128-bit stores of whole tile rows.
__global__ void __launch_bounds__(128)
idct16x16_decode_kernel(const DecodeParams p) {
  static_assert(kStageD16 >= 2 * kStageHalfD16, "pixel tile must cover the records");
  __shared__ __align__(128) uint32_t stage[kStageD16];
  __shared__ __align__(16) float tmp[kUnitD16 * kTmpBlkD16];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t t = threadIdx.x, b = (t >> 5) + ((t >> 2) & 4u), r = t & 15u;
  const uint32_t nbx = p.pw / 16u, nby = p.ph / 16u;
  const uint32_t chunks_x = (nbx + kUnitD16 - 1u) / kUnitD16;
  uint32_t u = blockIdx.x;
  const uint32_t cxi = u % chunks_x; u /= chunks_x;
  const uint32_t by = u % nby, f = u / nby;
  const uint32_t bx0 = cxi * kUnitD16, n_act = min((uint32_t)kUnitD16, nbx - bx0);
  const uint8_t* src = p.records + (uint64_t)f * p.frame_record_bytes + ((uint64_t)by * nbx + bx0) * (kRecW16 * 4u);
  const uint32_t n0 = min(n_act, 4u);
  fetch_records(stage, src, n0 * kRecW16 * 4u, stage + kStageHalfD16, (n_act - n0) * kRecW16 * 4u, &bar);
  const bool active = b < n_act;
  const uint32_t* rec = stage + (b & 3u) * kRecW16 + (b >> 2) * kStageHalfD16;
  float* buf = tmp + b * kTmpBlkD16;
  const QuantStep q = make_quant_step(!active ? 1u : (in_gaze(p, (bx0 + b) * 16u, by * 16u) ? 1u : (rec[0] == 0u ? p.bg_q : p.fg_q)));
  float px[48];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    {
      float X[16], x[16];
      uint32_t w[16];
#pragma unroll
      for (int ky = 0; ky < 16; ++ky) w[ky] = active ? rec[1 + c * 256 + ky * 16 + r] : 0u;
      dequant_group<16>(w, q, X);
      idct16(X, x);
#pragma unroll
      for (int y = 0; y < 16; ++y) buf[y * kTmpPitchD16 + r] = x[y];
    }
    __syncthreads();
    {
      const float4* trow = reinterpret_cast<const float4*>(buf + r * kTmpPitchD16);
      const float4 a0 = trow[0], a1 = trow[1], a2 = trow[2], a3 = trow[3];
      const float X[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                           a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
      float x[16];
      idct16(X, x);
#pragma unroll
      for (int j = 0; j < 16; ++j) px[j * 3 + c] = x[j];
    }
    __syncthreads();
  }
  // The records are consumed (the barrier after the last column pass): the buffer becomes the pixel
  // tile, [block][row] rows of 48 floats at pitch 49 (+16 words for blocks 4..7: conflict free), so
  // that every output row of the unit (8 blocks x 192 B) leaves as contiguous 128-bit stores.
  uint32_t* tile = stage;
  {
    uint32_t* o = tile + (b * 16u + r) * kPixPitchD16 + (b >> 2) * 16u;
#pragma unroll
    for (int k = 0; k < 48; ++k) o[k] = __float_as_uint(px[k]);
  }
  __syncthreads();
  float* out = p.out + ((uint64_t)f * p.ph + (uint64_t)by * 16u) * p.pw * 3u + (uint64_t)bx0 * 48u;
  const uint32_t vec_per_row = n_act * 12u;
  for (uint32_t k = t; k < 16u * vec_per_row; k += 128u) {
    const uint32_t y = k / vec_per_row, i = k - y * vec_per_row;
    const uint32_t bb = i / 12u, k4 = i - bb * 12u;
    const uint32_t* sp = tile + (bb * 16u + y) * kPixPitchD16 + (bb >> 2) * 16u + k4 * 4u;
    *reinterpret_cast<float4*>(out + (uint64_t)y * p.pw * 3u + (uint64_t)i * 4u) =
        make_float4(__uint_as_float(sp[0]), __uint_as_float(sp[1]), __uint_as_float(sp[2]), __uint_as_float(sp[3]));
  }
}

cudaError_t launch_decode(const DecodeParams& p, cudaStream_t st) {
  if (p.n_frames == 0) return cudaSuccess;
  const uint32_t tb = p.tb ? p.tb : 8u;
  const uint32_t unit = tb == 8u ? 32u : (tb == 16u ? (uint32_t)kUnitD16 : (uint32_t)kUnitD4);
  const uint32_t nbx = p.pw / tb, nby = p.ph / tb;
  const uint64_t ctas = (uint64_t)((nbx + unit - 1u) / unit) * nby * p.n_frames;
  if (ctas > 0x7fffffffull) return cudaErrorInvalidValue;
  if (tb == 8u) idct8x8_decode_kernel<<<(uint32_t)ctas, 96, 0, st>>>(p);
  else if (tb == 16u) idct16x16_decode_kernel<<<(uint32_t)ctas, 128, 0, st>>>(p);
  else if (tb == 4u) idct4x4_decode_kernel<<<(uint32_t)ctas, kUnitD4, 0, st>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace svc
