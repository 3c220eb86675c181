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
// 768 B out per block.
#include "common.cuh"

namespace svc {

#define SVC_C4 0.35355339059327376220f
#define SVC_A  0.49039264020161522456f
#define SVC_B2 0.46193976625564337806f
#define SVC_B  0.41573480615127261854f
#define SVC_C  0.27778511650980111237f
#define SVC_B6 0.19134171618254488586f
#define SVC_D  0.09754516100806413392f

// x = C^T X for the orthonormal 8-point DCT-II matrix C (even/odd split, 34 flops)
__device__ __forceinline__ void idct8(float& x0, float& x1, float& x2, float& x3,
                                      float& x4, float& x5, float& x6, float& x7) {
  const float p = SVC_C4 * (x0 + x4), q = SVC_C4 * (x0 - x4);
  const float r = fmaf(SVC_B2, x2, SVC_B6 * x6), s = fmaf(SVC_B6, x2, -SVC_B2 * x6);
  const float e0 = p + r, e3 = p - r, e1 = q + s, e2 = q - s;
  const float o0 = fmaf(SVC_A, x1, fmaf(SVC_B, x3, fmaf(SVC_C, x5, SVC_D * x7)));
  const float o1 = fmaf(SVC_B, x1, fmaf(-SVC_D, x3, fmaf(-SVC_A, x5, -SVC_C * x7)));
  const float o2 = fmaf(SVC_C, x1, fmaf(-SVC_A, x3, fmaf(SVC_D, x5, SVC_B * x7)));
  const float o3 = fmaf(SVC_D, x1, fmaf(-SVC_C, x3, fmaf(SVC_B, x5, -SVC_A * x7)));
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
    const float q = (float)(gazed ? 1u : (type == 0u ? p.bg_q : p.fg_q));
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float cf = __uint_as_float(rec[1 + c * 64 + r * 8 + j]);
        v[r][j] = __fmul_rn(roundf(__fdiv_rn(cf, q)), q);  // libs/decoder.cpp:140-142
      }
#pragma unroll
    for (int r = 0; r < 8; ++r) idct8(v[r][0], v[r][1], v[r][2], v[r][3], v[r][4], v[r][5], v[r][6], v[r][7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
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

cudaError_t launch_decode(const DecodeParams& p, cudaStream_t st) {
  if (p.n_frames == 0) return cudaSuccess;
  const uint32_t nbx = p.pw / 8u, nby = p.ph / 8u;
  const uint64_t ctas = (uint64_t)((nbx + 31u) / 32u) * nby * p.n_frames;
  if (ctas > 0x7fffffffull) return cudaErrorInvalidValue;
  idct8x8_decode_kernel<<<(uint32_t)ctas, 96, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace svc
