// k_dct.cu -- K3: per-channel block DCT of the zero-padded BGR frame and the
// codec's stream layout.
//
// Replaces Mat::convertTo(CV_32FC3) + Dct (cv::split + in-place cv::dct per
// block; reference libs/encoder.cpp:638-640, 323-339) and
// SerializeEncodedFrame (libs/encoder.cpp:222-269).  cv::dct is the
// orthonormal 2-D DCT-II on raw 0..255 values (no level shift).
//
// Fast path (8x8 transform blocks -- the encoder default, apps/encoder.cpp:56-57):
// one lane owns one 8x8 block, walks the three channels in registers
// (even/odd-decomposed 8-point DCT, 36 flops per 1-D transform) and either
//   * stores planar rows straight from registers (128-bit, coalesced), or
//   * stages whole 772-byte records in shared memory in stream order and hands
//     each warp's contiguous span to the TMA engine (cp.async.bulk shared ->
//     global), so the record layout costs no scattered stores.
// 16x16 and 4x4 transform blocks have their own fused stream kernels with the same contract
// (dct16x16_stream_kernel: row pass / transposed scratch tile / column pass with an even-odd split
// 16-point transform; dct4x4_stream_kernel: one lane per block).
// Generic path (any other transform block up to 32x32, or W != padded W): separable matrix form with
// the basis in constant memory, plus a gather kernel that reproduces the
// reference serializer index arithmetic exactly (including its use of the
// unpadded width as row stride, libs/encoder.cpp:257-262).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace svc {

// ---- 8-point orthonormal DCT-II ---------------------------------------------
// C[k][n] = s_k cos(pi (2n+1) k / 16), s_0 = sqrt(1/8), s_k = 1/2.
#define SVC_C4 0.35355339059327376220f /* sqrt(1/8)        */
#define SVC_A  0.49039264020161522456f /* cos(1 pi/16) / 2 */
#define SVC_B2 0.46193976625564337806f /* cos(2 pi/16) / 2 */
#define SVC_B  0.41573480615127261854f /* cos(3 pi/16) / 2 */
#define SVC_C  0.27778511650980111237f /* cos(5 pi/16) / 2 */
#define SVC_B6 0.19134171618254488586f /* cos(6 pi/16) / 2 */
#define SVC_D  0.09754516100806413392f /* cos(7 pi/16) / 2 */

// sqrt(1/8) cos(k pi / 16), k = 1..7 and 1/4 for the DC term: the 8-point factors times 1/sqrt(2),
// rounded once from the exact value (a float product of two rounded factors is up to 2 ulp off,
// and 0.35355339f * 0.70710678f is 0.24999998, which alone costs 5e-4 on a DC of 4080)
#define SVC_H4 0.25f
#define SVC_HA  0.3467599613305369f
#define SVC_HB2 0.32664074121909414f
#define SVC_HB  0.2939689006048397f
#define SVC_HC  0.1964237395967756f
#define SVC_HB6 0.13529902503654928f
#define SVC_HD  0.06897484482073578f
// The 2-D 8x8 transform applies the row pass with every factor times sqrt(2) and the column pass with
// every factor times 1/sqrt(2): the product is unchanged, but both DC factors (1/2 and 1/4) are exact,
// which halves the worst-case error on bright blocks.  sqrt(1/2) cos(k pi / 16), k = 1..7:
#define SVC_R4 0.5f
#define SVC_RA  0.6935199226610738f
#define SVC_RB2 0.6532814824381883f
#define SVC_RB  0.5879378012096794f
#define SVC_RC  0.3928474791935512f
#define SVC_RB6 0.27059805007309856f
#define SVC_RD  0.13794968964147156f

__device__ __forceinline__ void dct8(float& x0, float& x1, float& x2, float& x3,
                                     float& x4, float& x5, float& x6, float& x7) {
  const float s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
  const float d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
  const float e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
  x0 = SVC_H4 * (e0 + e1);
  x4 = SVC_H4 * (e0 - e1);
  x2 = fmaf(SVC_HB2, e2, SVC_HB6 * e3);
  x6 = fmaf(SVC_HB6, e2, -SVC_HB2 * e3);
  x1 = fmaf(SVC_HA, d0, fmaf(SVC_HB, d1, fmaf(SVC_HC, d2, SVC_HD * d3)));
  x3 = fmaf(SVC_HB, d0, fmaf(-SVC_HD, d1, fmaf(-SVC_HA, d2, -SVC_HC * d3)));
  x5 = fmaf(SVC_HC, d0, fmaf(-SVC_HA, d1, fmaf(SVC_HD, d2, SVC_HB * d3)));
  x7 = fmaf(SVC_HD, d0, fmaf(-SVC_HC, d1, fmaf(SVC_HB, d2, -SVC_HA * d3)));
}

// Row pass on "magic" floats.  A byte b placed in bits [15:8] of 0x47000000 is
// the float 2^15 + b (one PRMT, no I2F).  Differences of two such values are
// exact and offset free, sums are exact too (2^16 + b0 + b7, ...), so the only
// place the offset survives is the DC term, where 8 * 2^15 is subtracted once:
// 8 PRMT + 1 FADD per row instead of 8 PRMT + 8 FADD.
__device__ __forceinline__ float byte_to_magic(uint32_t w, int b) {
  return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504u | ((uint32_t)b << 4)));
}

__device__ __forceinline__ void dct8_magic(float& x0, float& x1, float& x2, float& x3,
                                           float& x4, float& x5, float& x6, float& x7) {
  const float s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;   // 2^16 + ..
  const float d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;   // exact
  const float e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
  x0 = SVC_R4 * ((e0 + e1) - 262144.0f);
  x4 = SVC_R4 * (e0 - e1);
  x2 = fmaf(SVC_RB2, e2, SVC_RB6 * e3);
  x6 = fmaf(SVC_RB6, e2, -SVC_RB2 * e3);
  x1 = fmaf(SVC_RA, d0, fmaf(SVC_RB, d1, fmaf(SVC_RC, d2, SVC_RD * d3)));
  x3 = fmaf(SVC_RB, d0, fmaf(-SVC_RD, d1, fmaf(-SVC_RA, d2, -SVC_RC * d3)));
  x5 = fmaf(SVC_RC, d0, fmaf(-SVC_RA, d1, fmaf(SVC_RD, d2, SVC_RB * d3)));
  x7 = fmaf(SVC_RD, d0, fmaf(-SVC_RC, d1, fmaf(SVC_RB, d2, -SVC_RA * d3)));
}

// 8x8 block of channel C out of 8 rows x 24 interleaved bytes -> 64 coefficients
template <int C>
__device__ __forceinline__ void dct_block(const uint32_t (&raw)[8][6], float (&v)[8][8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[r][j] = byte_to_magic(raw[r][(3 * j + C) >> 2], (3 * j + C) & 3);
    dct8_magic(v[r][0], v[r][1], v[r][2], v[r][3], v[r][4], v[r][5], v[r][6], v[r][7]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
}

// 8 rows x 24 bytes of the lane's block; zero outside the unpadded frame
// (cv::copyMakeBorder with zeros, libs/encoder.cpp:459-461).
__device__ __forceinline__ void load_block_rows(const DctParams& p, const uint8_t* fr, bool active,
                                                uint32_t px, uint32_t py, uint32_t (&raw)[8][6]) {
  const bool inside = active && (px + 8u <= p.w) && (py + 8u <= p.h);
  const bool vec_ok = ((p.w & 7u) == 0) && ((reinterpret_cast<uintptr_t>(p.bgr) & 7u) == 0);
  if (inside && vec_ok) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint2* q = reinterpret_cast<const uint2*>(fr + ((uint64_t)(py + r) * p.w + px) * 3u);
      const uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
      raw[r][0] = a.x; raw[r][1] = a.y; raw[r][2] = b.x;
      raw[r][3] = b.y; raw[r][4] = c.x; raw[r][5] = c.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < 6; ++k) raw[r][k] = 0;
      if (active && py + r < p.h) {
        const uint8_t* q = fr + ((uint64_t)(py + r) * p.w + px) * 3u;
        const uint32_t nb = px < p.w ? min(8u, p.w - px) * 3u : 0u;
#pragma unroll
        for (int k = 0; k < 24; ++k)
          if ((uint32_t)k < nb) raw[r][k >> 2] |= (uint32_t)__ldg(q + k) << (8 * (k & 3));
      }
    }
  }
}

constexpr int kRecWords8 = 193;  // 772-byte record = 1 + 3*64 words

// ---- planar output (drop-in for Dct, libs/encoder.cpp:323-339) -------------------
// One lane = one 8x8 block, three channels in turn, rows stored straight from
// registers as 2 x 128-bit (32 lanes x 32 B = 1 KB contiguous per row).
constexpr int kPlanarWarps = 4;
__global__ void __launch_bounds__(kPlanarWarps * 32)
dct8x8_planar_kernel(DctParams p, uint32_t nbx, uint32_t nby) {
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t per_frame = nbx * nby;
  const uint32_t chunks_per_frame = (per_frame + 31u) / 32u;
  const uint64_t unit = (uint64_t)blockIdx.x * kPlanarWarps + wib;
  if (unit >= (uint64_t)chunks_per_frame * p.n_frames) return;
  const uint32_t f = (uint32_t)(unit / chunks_per_frame);
  const uint32_t n = (uint32_t)(unit % chunks_per_frame) * 32u + lane;
  const bool active = n < per_frame;
  const uint32_t tbx = active ? n % nbx : 0u, tby = active ? n / nbx : 0u;
  const uint32_t px = tbx * 8u, py = tby * 8u;
  uint32_t raw[8][6];
  load_block_rows(p, p.bgr + (uint64_t)f * p.h * p.w * 3u, active, px, py, raw);
  if (!active) return;
  float v[8][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c == 0) dct_block<0>(raw, v);
    else if (c == 1) dct_block<1>(raw, v);
    else dct_block<2>(raw, v);
    float* pl = p.planes + (((uint64_t)f * 3u + c) * p.ph + py) * (uint64_t)p.pw + px;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float4* o = reinterpret_cast<float4*>(pl + (uint64_t)u * p.pw);
      o[0] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
      o[1] = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
    }
  }
}

// ---- stream output (Dct + SerializeEncodedFrame) with fused Y extraction ------------
// A CTA = 3 warps = one unit of 32 consecutive 8x8 blocks in serializer order;
// warp c transforms channel c (B, G, R) of all 32 blocks, so the 772-byte
// records are assembled by three warps into ONE shared staging buffer
// (24.7 KB per 96 threads instead of per 32: three times the resident warps).
// After a CTA barrier one elected lane hands the contiguous 32-record span to
// the TMA engine (cp.async.bulk shared -> global); no scattered stores exist.
// With kWithY the same pass also emits the level-0 luma rows of the blocks
// (zero pad + BGR2YUV + extractChannel(0), libs/encoder.cpp:459-469): Y =
// (1868 B + 9617 G + 4899 R + 8192) >> 14 as two 16x8-bit dot products (dp2a);
// warp c takes block rows c, c+3, c+6.  Then K1 only has to downsample.
struct YOut {
  uint8_t* l0;          // slot array + level-0 offset (null: no Y output)
  uint64_t slot_bytes;
  uint32_t first_slot;  // frame f of this launch -> slot first_slot + f
  uint32_t pitch;
};

__device__ __forceinline__ uint32_t luma_q14(uint32_t px /* B | G<<8 | R<<16 | x<<24 */) {
  const uint32_t t = __dp2a_lo(1868u | (9617u << 16), px, 8192u);  // B, G
  return __dp2a_hi(4899u, px, t) >> 14;                            // R (top byte x 0)
}

__device__ __forceinline__ uint2 luma_row8(const uint32_t (&w)[6]) {
  const uint32_t y0 = luma_q14(w[0]);
  const uint32_t y1 = luma_q14(__funnelshift_r(w[0], w[1], 24));
  const uint32_t y2 = luma_q14(__funnelshift_r(w[1], w[2], 16));
  const uint32_t y3 = luma_q14(w[2] >> 8);
  const uint32_t y4 = luma_q14(w[3]);
  const uint32_t y5 = luma_q14(__funnelshift_r(w[3], w[4], 24));
  const uint32_t y6 = luma_q14(__funnelshift_r(w[4], w[5], 16));
  const uint32_t y7 = luma_q14(w[5] >> 8);
  return make_uint2(y0 | (y1 << 8) | (y2 << 16) | (y3 << 24),
                    y4 | (y5 << 8) | (y6 << 16) | (y7 << 24));
}

template <bool kWithY>
__global__ void __launch_bounds__(96)
dct8x8_stream_kernel(const DctParams p, const uint32_t nbx, const uint32_t nby_stream,
                     const uint32_t nby_total, const YOut yo) {
  __shared__ __align__(128) uint32_t stage[32 * kRecWords8];
  const uint32_t lane = threadIdx.x & 31u, c = threadIdx.x >> 5;  // c: channel of this warp
  const uint32_t per_frame = nbx * nby_total;      // blocks visited (Y needs the padded rows)
  const uint32_t per_stream = nbx * nby_stream;    // blocks that have a record
  const uint32_t chunks_per_frame = (per_frame + 31u) / 32u;
  const uint32_t f = blockIdx.x / chunks_per_frame;
  const uint32_t n0 = (blockIdx.x % chunks_per_frame) * 32u;
  const uint32_t n = n0 + lane;
  const bool active = n < per_frame;
  const uint32_t tbx = active ? n % nbx : 0u, tby = active ? n / nbx : 0u;
  const uint32_t px = tbx * 8u, py = tby * 8u;

  // 8 rows x 24 bytes; the host guarantees w % 8 == 0 and an 8-byte aligned base,
  // so only whole rows can fall outside the frame (zero padding at the bottom)
  uint32_t raw[8][7];
  {
    const uint8_t* q0 = p.bgr + (((uint64_t)f * p.h + py) * p.w + px) * 3u;
    const uint64_t row_bytes = (uint64_t)p.w * 3u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      uint2 a = make_uint2(0, 0), b = a, d = a;
      if (active && py + r < p.h) {
        const uint2* q = reinterpret_cast<const uint2*>(q0 + r * row_bytes);
        a = __ldg(q); b = __ldg(q + 1); d = __ldg(q + 2);
      }
      raw[r][0] = a.x; raw[r][1] = a.y; raw[r][2] = b.x;
      raw[r][3] = b.y; raw[r][4] = d.x; raw[r][5] = d.y; raw[r][6] = 0;
    }
  }

  if (kWithY && active) {
    uint8_t* yrow = yo.l0 + (uint64_t)(yo.first_slot + f) * yo.slot_bytes + (uint64_t)py * yo.pitch + px;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if ((uint32_t)(r % 3) == c) {
        const uint32_t (&w)[7] = raw[r];
        const uint32_t row6[6] = {w[0], w[1], w[2], w[3], w[4], w[5]};
        *reinterpret_cast<uint2*>(yrow + (uint64_t)r * yo.pitch) = luma_row8(row6);
      }
  }
  if (n0 >= per_stream) return;  // CTA-uniform: padded block rows carry no record

  // Rotate every row by c bytes so this warp's channel sits at byte 3j: one code
  // path (dct_block<0>) serves all three warps and the kernel stays I-cache sized.
  const uint32_t sh = c * 8u;
  uint32_t rot[8][6];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int k = 0; k < 6; ++k) rot[r][k] = __funnelshift_r(raw[r][k], raw[r][k + 1], sh);
  float v[8][8];
  dct_block<0>(rot, v);

  // word (lane*193 + const) -> bank (lane + const) % 32: conflict free
  uint32_t* rec = stage + lane * kRecWords8;
  if (c == 0) {
    uint32_t bt = 0;
    if (p.block_types && n < per_stream)
      bt = __ldg(p.block_types + (uint64_t)f * p.mv_field_w * p.mv_field_h +
                 (py / p.mv_block_h) * p.mv_field_w + px / p.mv_block_w);
    rec[0] = bt;
  }
  uint32_t* recc = rec + 1 + c * 64;
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) recc[u * 8 + j] = __float_as_uint(v[u][j]);

  const uint32_t n_act = min(32u, per_stream - n0);
  const uint32_t bytes = n_act * kRecWords8 * 4u;
  uint8_t* dst = p.stream + (uint64_t)f * p.frame_stream_bytes + (uint64_t)n0 * (kRecWords8 * 4u);
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
  if (bulk_ok) {
    // generic-proxy smem writes -> visible to the async proxy, then one lane issues
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(stage);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(dst), "r"(s_addr), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the read
    }
  } else {
    __syncthreads();
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    for (uint32_t k = threadIdx.x; k < bytes / 4u; k += 96u) d32[k] = stage[k];
  }
}

// ---- fused stream kernels for 16x16 and 4x4 transform blocks -------------------------
// Same contract as dct8x8_stream_kernel (Dct + SerializeEncodedFrame, optional level-0 luma, one
// contiguous TMA bulk store per unit of consecutive records) for the other square transform
// blocks the encoder accepts by default geometry (libs/encoder.cpp:119-139: any block that
// divides the 16x16 motion block).  Needs w == pw like the 8x8 kernel, so that the serializer's
// unpadded row stride and swapped loop bounds (libs/encoder.cpp:257-262) coincide with the plane.

// 8-point transform of the even half of a 16-point one: every factor times 1/sqrt(2); `dc_off`
// is what the DC sum carries when the inputs are magic floats (0 otherwise).
__device__ __forceinline__ void dct8_half(const float (&s)[8], float (&o)[8], const float dc_off) {
  const float s0 = s[0] + s[7], s1 = s[1] + s[6], s2 = s[2] + s[5], s3 = s[3] + s[4];
  const float d0 = s[0] - s[7], d1 = s[1] - s[6], d2 = s[2] - s[5], d3 = s[3] - s[4];
  const float e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
  o[0] = SVC_H4 * ((e0 + e1) - dc_off);
  o[4] = SVC_H4 * (e0 - e1);
  o[2] = fmaf(SVC_HB2, e2, SVC_HB6 * e3);
  o[6] = fmaf(SVC_HB6, e2, -SVC_HB2 * e3);
  o[1] = fmaf(SVC_HA, d0, fmaf(SVC_HB, d1, fmaf(SVC_HC, d2, SVC_HD * d3)));
  o[3] = fmaf(SVC_HB, d0, fmaf(-SVC_HD, d1, fmaf(-SVC_HA, d2, -SVC_HC * d3)));
  o[5] = fmaf(SVC_HC, d0, fmaf(-SVC_HA, d1, fmaf(SVC_HD, d2, SVC_HB * d3)));
  o[7] = fmaf(SVC_HD, d0, fmaf(-SVC_HC, d1, fmaf(SVC_HB, d2, -SVC_HA * d3)));
}

// sqrt(2/16) cos(pi q / 32), q = 1, 3, .., 15
__device__ __forceinline__ constexpr float dct16_odd_factor(int i, int m) {
  constexpr float K[8] = {0.35185093438159565f, 0.33832950029358816f, 0.31180625324666783f,
                          0.2733004667504394f,  0.2242918965856591f,  0.1666639146194367f,
                          0.10263113188058934f, 0.034654292299772925f};
  int q = ((2 * i + 1) * (2 * m + 1)) % 64;  // cos(pi q / 32), q odd
  if (q > 32) q = 64 - q;
  const bool neg = q > 16;
  if (neg) q = 32 - q;
  return neg ? -K[(q - 1) / 2] : K[(q - 1) / 2];
}

// 16-point orthonormal DCT-II: X[2m] = DCT8(x_i + x_{15-i})[m] / sqrt 2, X[2m+1] = 8x8 odd matrix
// on x_i - x_{15-i} (exact differences, so magic-float inputs need no correction there).
__device__ __forceinline__ void dct16(const float (&x)[16], float (&X)[16], const float dc_off) {
  float s[8], d[8], e[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = x[i] + x[15 - i]; d[i] = x[i] - x[15 - i]; }
  dct8_half(s, e, dc_off);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    X[2 * m] = e[m];
    float acc = dct16_odd_factor(0, m) * d[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) acc = fmaf(dct16_odd_factor(i, m), d[i], acc);
    X[2 * m + 1] = acc;
  }
}

constexpr int kRecWords16 = 1 + 3 * 256;  // 3076-byte record
constexpr int kUnit16 = 8;                // records per CTA (24.6 KB of staging)
constexpr int kTmp16Pitch = 20;           // scratch tile [frequency k][row r]: 16 floats + 4 of padding, so the
                                          // column pass reads its 16 values as 4 conflict-free 128-bit loads
// A warp serves blocks w and w + 4 of the unit (16 lanes each).  Their scratch tiles and their records
// sit 16 banks apart, so that every shared-memory access of the two passes is conflict free:
// 4 * 324 = 16 (mod 32) for the tiles; the records of blocks 4..7 start 12 words after those of
// blocks 0..3 end (4 * 769 + 12 = 16 (mod 32)), and each half leaves as its own bulk store.
constexpr int kTmp16Blk = 16 * kTmp16Pitch + 4;
constexpr int kTmp16Buf = kUnit16 * kTmp16Blk;
constexpr int kStageHalf16 = 4 * kRecWords16 + 12;
constexpr int kIter16 = 2;                // consecutive units per CTA: the next unit's pixels are in flight during
                                          // the transforms (6 CTAs per SM: 35.4 KB of shared memory, 80 registers)

// shared -> global hand-over of a unit's contiguous record span (TMA bulk store when 16-byte
// aligned, cooperative word copy otherwise); all threads of the CTA call it
__device__ __forceinline__ void store_span(const uint32_t* stage, uint8_t* dst, const uint32_t bytes) {
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
  if (bulk_ok) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(stage);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(dst), "r"(s_addr), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    __syncthreads();
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    for (uint32_t k = threadIdx.x; k < bytes / 4u; k += blockDim.x) d32[k] = stage[k];
  }
}

// CTA = 128 threads; a unit = 8 consecutive 16x16 blocks in serializer order, kIter16 consecutive
// units per CTA.  Row pass: thread (block, row) holds the row's 48 interleaved bytes
// (3 x 128 bit, loaded while the previous unit was transformed), emits its 16 luma bytes (kWithY) and
// the 16-point transform of each channel into a padded, transposed scratch tile; column pass: thread
// (block, column) transforms the three channels' columns and writes the coefficients at their place
// in the record.  Then one bulk store of the 8-record span; its shared-memory read is awaited just
// before the next unit's first record word is written.
struct Unit16 {
  uint32_t f, n0, n, px, py, bt;
  bool active;
};
struct Div16 {  // ceil(2^64 / d): __umul64hi(n, m) == n / d for every 32-bit n
  uint64_t m_chunks, m_nbx;
};

template <bool kWithY>
__global__ void __launch_bounds__(128)
dct16x16_stream_kernel(const DctParams p, const uint32_t nbx, const uint32_t nby_stream,
                       const uint32_t nby_total, const YOut yo, const uint32_t total_units,
                       const uint32_t units_per_cta, const Div16 dv) {
  __shared__ __align__(128) uint32_t stage[2 * kStageHalf16];
  __shared__ __align__(16) float tmp[kTmp16Buf];
  const uint32_t t = threadIdx.x, b = (t >> 5) + ((t >> 2) & 4u), r = t & 15u;
  const uint32_t per_frame = nbx * nby_total, per_stream = nbx * nby_stream;
  const uint32_t chunks_per_frame = (per_frame + kUnit16 - 1u) / kUnit16;
  auto locate = [&](const uint32_t u) {
    Unit16 q;
    q.f = chunks_per_frame > 1u ? (uint32_t)__umul64hi(u, dv.m_chunks) : u;
    q.n0 = (u - q.f * chunks_per_frame) * kUnit16;
    q.n = q.n0 + b;
    q.active = q.n < per_frame;
    const uint32_t tby = !q.active ? 0u : (nbx > 1u ? (uint32_t)__umul64hi(q.n, dv.m_nbx) : q.n);
    const uint32_t tbx = q.active ? q.n - tby * nbx : 0u;
    q.px = tbx * 16u;
    q.py = tby * 16u;
    q.bt = 0u;
    if (r == 0 && p.block_types && q.n < per_stream)
      q.bt = __ldg(p.block_types + (uint64_t)q.f * p.mv_field_w * p.mv_field_h +
                   (q.py / p.mv_block_h) * p.mv_field_w + q.px / p.mv_block_w);
    return q;
  };
  auto fetch = [&](const Unit16& q, uint4& v0, uint4& v1, uint4& v2) {
    v0 = make_uint4(0, 0, 0, 0); v1 = v0; v2 = v0;
    if (q.active && q.py + r < p.h) {  // whole rows below the frame are the zero padding
      const uint4* g = reinterpret_cast<const uint4*>(p.bgr + (((uint64_t)q.f * p.h + q.py + r) * p.w + q.px) * 3u);
      v0 = __ldg(g); v1 = __ldg(g + 1); v2 = __ldg(g + 2);
    }
  };
  uint32_t u = blockIdx.x * units_per_cta;
  const uint32_t u_end = min(u + units_per_cta, total_units);
  Unit16 nxt = locate(u);
  uint4 v0, v1, v2;
  fetch(nxt, v0, v1, v2);
  bool pending = false;  // a bulk store may still be reading `stage`
  uint32_t* rec = stage + (b & 3u) * kRecWords16 + (b >> 2) * kStageHalf16;
  float* buf = tmp + b * kTmp16Blk;
#pragma unroll 1
  for (; u < u_end; ++u) {
    const Unit16 q = nxt;
    uint32_t raw[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
    if (u + 1 < u_end) {
      nxt = locate(u + 1);
      fetch(nxt, v0, v1, v2);
    }
    if (kWithY && q.active) {
      const uint32_t lo[6] = {raw[0], raw[1], raw[2], raw[3], raw[4], raw[5]};
      const uint32_t hi[6] = {raw[6], raw[7], raw[8], raw[9], raw[10], raw[11]};
      const uint2 y0 = luma_row8(lo), y1 = luma_row8(hi);
      uint8_t* yrow = yo.l0 + (uint64_t)(yo.first_slot + q.f) * yo.slot_bytes + (uint64_t)(q.py + r) * yo.pitch + q.px;
      *reinterpret_cast<uint4*>(yrow) = make_uint4(y0.x, y0.y, y1.x, y1.y);
    }
    if (q.n0 >= per_stream) continue;  // CTA-uniform: padded block rows carry no record

    // channel c: row pass (thread = row r of block b) into the scratch tile, barrier, column pass
    // (thread = column r of block b) into the record, barrier (the barrier of the store after the
    // last channel).  The row is shifted down one byte per channel so that one code path (channel
    // at byte 3j) serves B, G and R.
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      {
        float x[16], X[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = byte_to_magic(raw[(3 * j) >> 2], (3 * j) & 3);
        dct16(x, X, 524288.0f);  // 8 sums of two magic floats: 8 * 2^16
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[k * kTmp16Pitch + r] = X[k];
#pragma unroll
        for (int k = 0; k < 11; ++k) raw[k] = __funnelshift_r(raw[k], raw[k + 1], 8);
        raw[11] >>= 8;
      }
      if (c == 0 && pending && t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
      if (c == 0 && r == 0) rec[0] = q.bt;
      {
        const float4* tcol = reinterpret_cast<const float4*>(buf + r * kTmp16Pitch);
        const float4 a0 = tcol[0], a1 = tcol[1], a2 = tcol[2], a3 = tcol[3];
        const float x[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                             a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
        float X[16];
        dct16(x, X, 0.f);
        uint32_t* o = rec + 1 + c * 256 + r;
#pragma unroll
        for (int k = 0; k < 16; ++k) o[k * 16] = __float_as_uint(X[k]);
      }
      if (c < 2) __syncthreads();
    }
    const uint32_t n_act = min((uint32_t)kUnit16, per_stream - q.n0);
    const uint32_t bytes0 = min(n_act, 4u) * kRecWords16 * 4u, bytes1 = (n_act - min(n_act, 4u)) * kRecWords16 * 4u;
    uint8_t* dst = p.stream + (uint64_t)q.f * p.frame_stream_bytes + (uint64_t)q.n0 * (kRecWords16 * 4u);
    if (((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && (((bytes0 | bytes1) & 15u) == 0)) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (t == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(stage)), "r"(bytes0) : "memory");
        if (bytes1)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(dst + bytes0), "r"((uint32_t)__cvta_generic_to_shared(stage + kStageHalf16)), "r"(bytes1)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending = true;
    } else {
      __syncthreads();
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
      for (uint32_t k = t; k < bytes0 / 4u; k += 128u) d32[k] = stage[k];
      for (uint32_t k = t; k < bytes1 / 4u; k += 128u) d32[bytes0 / 4u + k] = stage[kStageHalf16 + k];
    }
  }
  if (pending && t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the read
}

// 4-point orthonormal DCT-II
__device__ __forceinline__ void dct4(float& x0, float& x1, float& x2, float& x3, const float dc_off) {
  constexpr float a = 0.6532814824381883f, bq = 0.27059805007309856f;  // sqrt(1/2) cos(pi/8), cos(3 pi/8)
  const float s0 = x0 + x3, s1 = x1 + x2, d0 = x0 - x3, d1 = x1 - x2;
  x0 = 0.5f * ((s0 + s1) - dc_off);
  x2 = 0.5f * (s0 - s1);
  x1 = fmaf(a, d0, bq * d1);
  x3 = fmaf(bq, d0, -a * d1);
}

constexpr int kRecWords4 = 1 + 3 * 16;  // 196-byte record
constexpr int kUnit4 = 128;             // records per CTA (24.5 KB of staging)

// One lane = one 4x4 block (4 rows x 12 interleaved bytes, three channels in registers); a CTA of
// 128 lanes assembles 128 consecutive records (lane * 49 + const: conflict free) and hands the span
// to the TMA engine.
template <bool kWithY>
__global__ void __launch_bounds__(kUnit4)
dct4x4_stream_kernel(const DctParams p, const uint32_t nbx, const uint32_t nby_stream,
                     const uint32_t nby_total, const YOut yo) {
  __shared__ __align__(128) uint32_t stage[kUnit4 * kRecWords4];
  const uint32_t per_frame = nbx * nby_total, per_stream = nbx * nby_stream;
  const uint32_t chunks_per_frame = (per_frame + kUnit4 - 1u) / kUnit4;
  const uint32_t f = blockIdx.x / chunks_per_frame;
  const uint32_t n0 = (blockIdx.x % chunks_per_frame) * kUnit4;
  const uint32_t n = n0 + threadIdx.x;
  const bool active = n < per_frame;
  const uint32_t tbx = active ? n % nbx : 0u, tby = active ? n / nbx : 0u;
  const uint32_t px = tbx * 4u, py = tby * 4u;
  uint32_t raw[4][3];
  {
    const uint8_t* q0 = p.bgr + (((uint64_t)f * p.h + py) * p.w + px) * 3u;
    const uint64_t row_bytes = (uint64_t)p.w * 3u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      raw[r][0] = raw[r][1] = raw[r][2] = 0u;
      if (active && py + r < p.h) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(q0 + r * row_bytes);
        raw[r][0] = __ldg(q); raw[r][1] = __ldg(q + 1); raw[r][2] = __ldg(q + 2);
      }
    }
  }
  if (kWithY && active) {
    uint8_t* yrow = yo.l0 + (uint64_t)(yo.first_slot + f) * yo.slot_bytes + (uint64_t)py * yo.pitch + px;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t (&w)[3] = raw[r];
      const uint32_t y0 = luma_q14(w[0]), y1 = luma_q14(__funnelshift_r(w[0], w[1], 24));
      const uint32_t y2 = luma_q14(__funnelshift_r(w[1], w[2], 16)), y3 = luma_q14(w[2] >> 8);
      *reinterpret_cast<uint32_t*>(yrow + (uint64_t)r * yo.pitch) = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
    }
  }
  if (n0 >= per_stream) return;
  uint32_t* rec = stage + threadIdx.x * kRecWords4;
  {
    uint32_t bt = 0;
    if (p.block_types && n < per_stream)
      bt = __ldg(p.block_types + (uint64_t)f * p.mv_field_w * p.mv_field_h +
                 (py / p.mv_block_h) * p.mv_field_w + px / p.mv_block_w);
    rec[0] = bt;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[r][j] = byte_to_magic(raw[r][(3 * j + c) >> 2], (3 * j + c) & 3);
      dct4(v[r][0], v[r][1], v[r][2], v[r][3], 131072.0f);  // 4 * 2^15
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) dct4(v[0][j], v[1][j], v[2][j], v[3][j], 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) rec[1 + c * 16 + u * 4 + j] = __float_as_uint(v[u][j]);
  }
  const uint32_t n_act = min((uint32_t)kUnit4, per_stream - n0);
  store_span(stage, p.stream + (uint64_t)f * p.frame_stream_bytes + (uint64_t)n0 * (kRecWords4 * 4u),
             n_act * kRecWords4 * 4u);
}

// ---- generic separable path (any transform block up to 32 x 32) ---------------
// out = Ch . X . Cw^T per block, C[k][n] = s_k cos(pi (2n+1) k / (2N)) (cv::dct, libs/encoder.cpp:335).
// One CTA transforms a tile of whole blocks of one channel in shared memory: the tile's pixels are
// converted to float once, the row pass and the column pass each read a padded basis matrix that
// travelled as a kernel parameter (no constant-bank upload, no host synchronisation), and the
// coefficients leave as coalesced row segments.  Sums run over n ascending with fmaf.
struct DctBasis {
  float w[32 * 32];  // Cw, tbw x tbw, row k = frequency
  float h[32 * 32];  // Ch, tbh x tbh
};

constexpr int kGenTilePx = 1024;  // pixels per tile (upper bound; whole blocks only)

__global__ void __launch_bounds__(256)
dct_generic_planar_kernel(const uint8_t* __restrict__ bgr, uint32_t w, uint32_t h, uint32_t pw, uint32_t ph,
                          uint32_t tbw, uint32_t tbh, uint32_t tw, uint32_t th, uint32_t tiles_x,
                          uint32_t tiles_y, float* __restrict__ planes,
                          const __grid_constant__ DctBasis basis) {
  __shared__ float sIn[kGenTilePx + 64];   // th rows, pitch tw + 1
  __shared__ float sTmp[kGenTilePx + 64];
  __shared__ float sCw[32 * 33], sCh[32 * 33];
  __shared__ uint8_t sKx[64], sKy[32];     // frequency index of a tile column / row inside its block
  // blockDim = (64, 4): a thread keeps its column(s), rows advance by 4 -- no division per pixel
  const uint32_t tx = threadIdx.x, ty = threadIdx.y, tid = ty * 64u + tx;
  const uint32_t tile = blockIdx.x, c = blockIdx.y, f = blockIdx.z;
  const uint32_t tile_y = tile / tiles_x;
  const uint32_t tx0 = (tile - tile_y * tiles_x) * tw, ty0 = tile_y * th;
  const uint32_t pitch = tw + 1;
  for (uint32_t i = tid; i < tbw * tbw; i += 256) sCw[(i / tbw) * 33 + i % tbw] = basis.w[i];
  for (uint32_t i = tid; i < tbh * tbh; i += 256) sCh[(i / tbh) * 33 + i % tbh] = basis.h[i];
  if (tid < tw) sKx[tid] = (uint8_t)(tid % tbw);
  if (tid < th) sKy[tid] = (uint8_t)(tid % tbh);
  const uint8_t* src = bgr + (uint64_t)f * w * h * 3u + c;
  for (uint32_t y = ty; y < th; y += 4) {
    const uint32_t gy = ty0 + y;
    const uint8_t* row = src + (uint64_t)gy * w * 3u;
    for (uint32_t x = tx; x < tw; x += 64) {
      const uint32_t gx = tx0 + x;  // zero padding right / below (libs/encoder.cpp:459-461)
      sIn[y * pitch + x] = (gx < w && gy < h) ? (float)__ldg(row + gx * 3u) : 0.f;
    }
  }
  __syncthreads();
  for (uint32_t x = tx; x < tw; x += 64) {  // row pass
    const uint32_t k = sKx[x], x0 = x - k;
    const float* cw = sCw + k * 33;
    for (uint32_t y = ty; y < th; y += 4) {
      const float* in = sIn + y * pitch + x0;
      float acc = 0.f;
      for (uint32_t j = 0; j < tbw; ++j) acc = fmaf(cw[j], in[j], acc);
      sTmp[y * pitch + x] = acc;
    }
  }
  __syncthreads();
  float* dst = planes + ((uint64_t)f * 3u + c) * pw * ph;
  for (uint32_t y = ty; y < th; y += 4) {  // column pass
    const uint32_t k = sKy[y], y0 = y - k, gy = ty0 + y;
    const float* ch = sCh + k * 33;
    for (uint32_t x = tx; x < tw; x += 64) {
      const float* in = sTmp + y0 * pitch + x;
      float acc = 0.f;
      for (uint32_t j = 0; j < tbh; ++j) acc = fmaf(ch[j], in[j * pitch], acc);
      const uint32_t gx = tx0 + x;
      if (gx < pw && gy < ph) dst[(uint64_t)gy * pw + gx] = acc;
    }
  }
}

// Square power-of-two blocks (2, 4, 16, 32): the same two passes with the loops unrolled.  A thread
// takes one block row (then one block column): TB values into registers, TB x TB FMAs whose basis
// factors are constant-bank operands (the basis is a kernel parameter, the indices are compile
// time), so the FMA pipe is the only busy unit.  Tile = 64 x 64 pixels of one channel.
template <int TB>
__global__ void __launch_bounds__(256)
dct_square_planar_kernel(const uint8_t* __restrict__ bgr, uint32_t w, uint32_t h, uint32_t pw, uint32_t ph,
                         uint32_t tiles_x, float* __restrict__ planes, const __grid_constant__ DctBasis basis) {
  constexpr int T = 64, PITCH = T + 1, NBT = T / TB;  // blocks per tile row / column
  __shared__ float sA[T * PITCH], sB[T * PITCH];
  const uint32_t tid = threadIdx.x;
  const uint32_t tile = blockIdx.x, c = blockIdx.y, f = blockIdx.z;
  const uint32_t tile_y = tile / tiles_x;
  const uint32_t tx0 = (tile - tile_y * tiles_x) * T, ty0 = tile_y * T;
  const uint8_t* src = bgr + (uint64_t)f * w * h * 3u + c;
  for (uint32_t i = tid; i < T * T; i += 256) {
    const uint32_t y = i >> 6, x = i & 63u, gx = tx0 + x, gy = ty0 + y;
    sA[y * PITCH + x] = (gx < w && gy < h) ? (float)__ldg(src + ((uint64_t)gy * w + gx) * 3u) : 0.f;
  }
  __syncthreads();
  for (uint32_t it = tid; it < T * NBT; it += 256) {  // row pass: (row y, block bx)
    const uint32_t y = it / NBT, bx = it - y * NBT;
    float v[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) v[j] = sA[y * PITCH + bx * TB + j];
#pragma unroll
    for (int k = 0; k < TB; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < TB; ++j) acc = fmaf(basis.w[k * TB + j], v[j], acc);
      sB[y * PITCH + bx * TB + k] = acc;
    }
  }
  __syncthreads();
  float* dst = planes + ((uint64_t)f * 3u + c) * pw * ph;
  for (uint32_t it = tid; it < T * NBT; it += 256) {  // column pass: (block by, column x)
    const uint32_t by = it >> 6, x = it & 63u;
    float v[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) v[j] = sB[(by * TB + j) * PITCH + x];
    const uint32_t gx = tx0 + x;
#pragma unroll
    for (int k = 0; k < TB; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < TB; ++j) acc = fmaf(basis.h[k * TB + j], v[j], acc);
      const uint32_t gy = ty0 + by * TB + k;
      if (gx < pw && gy < ph) dst[(uint64_t)gy * pw + gx] = acc;
    }
  }
}

// SerializeEncodedFrame, libs/encoder.cpp:243-266: one thread per stream word of one frame
// (blockIdx.y = frame).  Record = block type word + 3 channels of tbw rows x tbh floats -- the
// reference's swapped loop bounds and unpadded row stride are kept.  Every division is a
// multiply-high by a host-computed reciprocal (exact: dividend x divisor < 2^32, checked on the host).
struct GatherDiv {
  uint32_t rec_words;               // words per record
  uint32_t nbx, m_nbx;              // records per block row
  uint32_t area, m_area;            // tbw * tbh
  uint32_t tbh, m_tbh;
};
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint32_t d, uint32_t m) {
  return d == 1u ? n : __umulhi(n, m);
}

__global__ void __launch_bounds__(256)
serialize_gather_kernel(const float* __restrict__ planes, uint64_t plane_elems,
                        const uint32_t* __restrict__ block_types,
                        uint32_t w, uint32_t tbw, uint32_t tbh,
                        uint32_t mbw, uint32_t mbh, uint32_t mvw, uint32_t mvh,
                        uint32_t* __restrict__ stream, uint32_t frame_words, GatherDiv dv) {
  const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= frame_words) return;
  const uint32_t f = blockIdx.y;
  const uint32_t rec = wi / dv.rec_words, k = wi - rec * dv.rec_words;  // (wi * rec_words may exceed 2^32)
  const uint32_t rec_y = div_magic(rec, dv.nbx, dv.m_nbx);
  const uint32_t tb_x = (rec - rec_y * dv.nbx) * tbw, tb_y = rec_y * tbh;
  uint32_t out;
  if (k == 0) {
    out = block_types ? block_types[(uint64_t)f * mvw * mvh + (tb_y / mbh) * mvw + tb_x / mbw] : 0u;
  } else {
    const uint32_t e = k - 1u, c = div_magic(e, dv.area, dv.m_area), rem = e - c * dv.area;
    const uint32_t row = div_magic(rem, dv.tbh, dv.m_tbh), j = rem - row * dv.tbh;  // tbw rows of tbh floats (sic)
    const uint64_t idx = (uint64_t)(tb_y + row) * w + tb_x + j;                     // unpadded stride (sic)
    out = idx < plane_elems ? __float_as_uint(planes[((uint64_t)f * 3u + c) * plane_elems + idx]) : 0u;
  }
  stream[(uint64_t)f * frame_words + wi] = out;
}

static void host_basis(uint32_t n, float* out) {
  for (uint32_t k = 0; k < n; ++k) {
    const double s = k == 0 ? sqrt(1.0 / n) : sqrt(2.0 / n);
    for (uint32_t i = 0; i < n; ++i)
      out[k * n + i] = (float)(s * cos(M_PI * (2.0 * i + 1.0) * k / (2.0 * n)));
  }
}

static cudaError_t planar_into(const DctParams& p, const uint8_t* bgr, uint32_t nf,
                               float* planes, float* tmp, cudaStream_t st, int* nl) {
  if (p.tbw == 8 && p.tbh == 8) {
    DctParams q = p;
    q.bgr = bgr; q.n_frames = nf; q.planes = planes;
    const uint32_t nbx = p.pw / 8, nby = p.ph / 8;
    const uint64_t units = (uint64_t)((nbx * nby + 31) / 32) * nf;
    dct8x8_planar_kernel<<<(uint32_t)((units + kPlanarWarps - 1) / kPlanarWarps), kPlanarWarps * 32, 0, st>>>(q, nbx, nby);
    if (nl) *nl += 1;
    return cudaGetLastError();
  }
  if (p.tbw > 32 || p.tbh > 32 || p.tbw == 0 || p.tbh == 0) return cudaErrorInvalidValue;
  DctBasis basis;
  host_basis(p.tbw, basis.w);
  host_basis(p.tbh, basis.h);
  if (p.tbw == p.tbh && (p.tbw == 2 || p.tbw == 4 || p.tbw == 16 || p.tbw == 32) && nf <= 65535u) {
    const uint32_t tiles_x = (p.pw + 63u) / 64u, tiles_y = (p.ph + 63u) / 64u;
    const dim3 grid(tiles_x * tiles_y, 3, nf);
    switch (p.tbw) {
      case 2: dct_square_planar_kernel<2><<<grid, 256, 0, st>>>(bgr, p.w, p.h, p.pw, p.ph, tiles_x, planes, basis); break;
      case 4: dct_square_planar_kernel<4><<<grid, 256, 0, st>>>(bgr, p.w, p.h, p.pw, p.ph, tiles_x, planes, basis); break;
      case 16: dct_square_planar_kernel<16><<<grid, 256, 0, st>>>(bgr, p.w, p.h, p.pw, p.ph, tiles_x, planes, basis); break;
      default: dct_square_planar_kernel<32><<<grid, 256, 0, st>>>(bgr, p.w, p.h, p.pw, p.ph, tiles_x, planes, basis); break;
    }
    if (nl) *nl += 1;
    return cudaGetLastError();
  }
  // tile = whole blocks, at most 1024 pixels: 16 rows (or one block row) x 64 columns where that fits
  const uint32_t th = p.tbh * std::max(1u, 16u / p.tbh);
  const uint32_t tw = p.tbw * std::max(1u, ((uint32_t)kGenTilePx / th) / p.tbw);
  if (th * (tw + 1) > (uint32_t)kGenTilePx + 64u) return cudaErrorInvalidValue;
  const uint32_t tiles_x = (p.pw + tw - 1) / tw, tiles_y = (p.ph + th - 1) / th;
  if (tw > 64u || th > 32u || nf > 65535u || (uint64_t)tiles_x * tiles_y > 0x7fffffffull) return cudaErrorInvalidValue;
  (void)tmp;
  dct_generic_planar_kernel<<<dim3(tiles_x * tiles_y, 3, nf), dim3(64, 4), 0, st>>>(
      bgr, p.w, p.h, p.pw, p.ph, p.tbw, p.tbh, tw, th, tiles_x, tiles_y, planes, basis);
  if (nl) *nl += 1;
  return cudaGetLastError();
}

cudaError_t prepare_dct_kernels() {
  // the staging buffers want the large shared-memory carve-out (9 CTAs x 24.7 KB)
  const void* fns[] = {(const void*)dct8x8_stream_kernel<true>,   (const void*)dct8x8_stream_kernel<false>,
                       (const void*)dct16x16_stream_kernel<true>, (const void*)dct16x16_stream_kernel<false>,
                       (const void*)dct4x4_stream_kernel<true>,   (const void*)dct4x4_stream_kernel<false>};
  for (const void* fn : fns) {
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// the fused stream kernels: square 8x8 / 16x16 / 4x4 blocks, no horizontal padding (so the
// serializer's unpadded row stride equals the plane stride), block rows aligned for vector loads
static bool dct_fast_stream_ok(const DctParams& p) {
  if (p.tbw != p.tbh || p.w != p.pw) return false;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p.bgr);
  switch (p.tbw) {
    case 8: return (p.w % 8u) == 0 && (a & 7u) == 0;
    case 16: return (p.w % 16u) == 0 && (a & 15u) == 0;
    case 4: return (p.w % 4u) == 0 && (a & 3u) == 0;
    default: return false;
  }
}

bool dct_needs_scratch(const DctParams& p) {
  const bool fast8 = (p.tbw == 8 && p.tbh == 8);
  return (p.planes && !fast8) || (p.stream && !dct_fast_stream_ok(p));
}

bool dct_can_fuse_y(const DctParams& p) {
  return p.stream && dct_fast_stream_ok(p) && (p.ph % p.tbh) == 0;
}

cudaError_t launch_dct(const DctParams& p, cudaStream_t st, int* nl) {
  if (p.n_frames == 0) return cudaSuccess;
  const bool fast8 = (p.tbw == 8 && p.tbh == 8);
  const uint64_t plane_elems = (uint64_t)p.pw * p.ph;
  const uint64_t in_frame = (uint64_t)p.w * p.h * 3u;
  cudaError_t e;
  if (p.planes) {
    if (fast8) {
      e = planar_into(p, p.bgr, p.n_frames, p.planes, nullptr, st, nl);
      if (e != cudaSuccess) return e;
    } else {
      if (!p.scratch_planes || p.scratch_frames == 0) return cudaErrorInvalidValue;
      for (uint32_t f0 = 0; f0 < p.n_frames; f0 += p.scratch_frames) {
        const uint32_t nf = min(p.scratch_frames, p.n_frames - f0);
        e = planar_into(p, p.bgr + f0 * in_frame, nf, p.planes + f0 * 3u * plane_elems,
                        p.scratch_planes, st, nl);
        if (e != cudaSuccess) return e;
      }
    }
  }
  if (p.stream) {
    if (dct_fast_stream_ok(p)) {
      const uint32_t tb = p.tbw, unit = tb == 8 ? 32u : (tb == 16 ? (uint32_t)kUnit16 : (uint32_t)kUnit4);
      const uint32_t nbx = p.w / tb, nby = (p.h + tb - 1) / tb;
      const bool with_y = p.y_l0 != nullptr;
      const uint32_t nby_total = with_y ? p.ph / tb : nby;
      const uint64_t units = (uint64_t)((nbx * nby_total + unit - 1) / unit) * p.n_frames;
      if (units > 0x7fffffffull) return cudaErrorInvalidValue;
      YOut yo{p.y_l0, p.y_slot_bytes, p.y_first_slot, p.y_pitch};
      if (tb == 16) {
        const uint32_t upc = (uint32_t)kIter16;  // units per CTA
        const uint32_t grid = (uint32_t)((units + upc - 1) / upc);
        auto magic64 = [](uint32_t d) { return d > 1u ? ~0ull / d + 1ull : 0ull; };  // ceil(2^64 / d)
        const Div16 dv{magic64((nbx * nby_total + kUnit16 - 1) / kUnit16), magic64(nbx)};
        if (with_y) dct16x16_stream_kernel<true><<<grid, 128, 0, st>>>(p, nbx, nby, nby_total, yo, (uint32_t)units, upc, dv);
        else dct16x16_stream_kernel<false><<<grid, 128, 0, st>>>(p, nbx, nby, nby_total, yo, (uint32_t)units, upc, dv);
      } else if (tb == 4) {
        if (with_y) dct4x4_stream_kernel<true><<<(uint32_t)units, kUnit4, 0, st>>>(p, nbx, nby, nby_total, yo);
        else dct4x4_stream_kernel<false><<<(uint32_t)units, kUnit4, 0, st>>>(p, nbx, nby, nby_total, yo);
      } else if (with_y) dct8x8_stream_kernel<true><<<(uint32_t)units, 96, 0, st>>>(p, nbx, nby, nby_total, yo);
      else dct8x8_stream_kernel<false><<<(uint32_t)units, 96, 0, st>>>(p, nbx, nby, nby_total, yo);
      if (nl) *nl += 1;
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    } else {
      // planar coefficients into scratch (second half of the scratch is the
      // row-pass temporary of the generic transform), then the exact gather
      if (!p.scratch_planes || p.scratch_frames == 0) return cudaErrorInvalidValue;
      float* sp = p.scratch_planes;
      float* tmp = p.scratch_planes + (uint64_t)p.scratch_frames * 3u * plane_elems;
      const uint64_t frame_words = p.frame_stream_bytes / 4u;
      for (uint32_t f0 = 0; f0 < p.n_frames; f0 += p.scratch_frames) {
        const uint32_t nf = min(p.scratch_frames, p.n_frames - f0);
        e = planar_into(p, p.bgr + f0 * in_frame, nf, sp, tmp, st, nl);
        if (e != cudaSuccess) return e;
        GatherDiv dv{};
        dv.rec_words = 1u + 3u * p.tbw * p.tbh;
        dv.nbx = (p.w + p.tbw - 1) / p.tbw;
        dv.area = p.tbw * p.tbh;
        dv.tbh = p.tbh;
        auto magic = [](uint32_t d) { return d > 1u ? 0xffffffffu / d + 1u : 0u; };  // ceil(2^32 / d)
        dv.m_nbx = magic(dv.nbx);
        dv.m_area = magic(dv.area);
        dv.m_tbh = magic(dv.tbh);
        // multiply-high by ceil(2^32/d) is the exact quotient while dividend * d < 2^32: records per
        // frame x records per row, and word-in-record x block area, are far below that
        if (frame_words >= (1ull << 32) || nf > 65535u) return cudaErrorInvalidValue;
        serialize_gather_kernel<<<dim3((uint32_t)((frame_words + 255) / 256), nf), 256, 0, st>>>(
            sp, plane_elems,
            p.block_types ? p.block_types + (uint64_t)f0 * p.mv_field_w * p.mv_field_h : nullptr,
            p.w, p.tbw, p.tbh, p.mv_block_w, p.mv_block_h, p.mv_field_w, p.mv_field_h,
            reinterpret_cast<uint32_t*>(p.stream + (uint64_t)f0 * p.frame_stream_bytes),
            (uint32_t)frame_words, dv);
        if (nl) *nl += 1;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
      }
    }
  }
  return cudaSuccess;
}

}  // namespace svc
