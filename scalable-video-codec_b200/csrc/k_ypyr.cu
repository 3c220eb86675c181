// k_ypyr.cu -- K1: zero pad + BGR->Y extraction and Gaussian pyramid.
//
// Replaces, on the device, the OpenCV calls of the reference encoder loop:
//   cv::copyMakeBorder (libs/encoder.cpp:459-461)   -> folded in: out-of-frame
//                                                      pixels read as zero
//   cv::cvtColor BGR2YUV + extractChannel(0) (:468-469)
//                                                   -> Y = (1868 B + 9617 G +
//                                                      4899 R + 8192) >> 14
//   cv::buildPyramid (:470) = iterated cv::pyrDown  -> [1 4 6 4 1]^2,
//                                                      (sum + 128) >> 8,
//                                                      BORDER_REFLECT_101
// All arithmetic is integer and bit-exact with OpenCV's 8-bit paths.
#include "common.cuh"

namespace svc {

// ---------------------------------------------------------------------------
// BGR -> Y.  One thread produces 4 luma bytes (one 32-bit store) from 12 input
// bytes.  The Q14 weights do not fit a byte, so each is split hi*256 + lo and
// the dot product runs as two packed-byte dp4a per pixel.
//   1868 = 7*256 + 76   9617 = 37*256 + 145   4899 = 19*256 + 35
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t y_from_packed(uint32_t px /* B | G<<8 | R<<16 */) {
  const uint32_t lo = __dp4a(px, 0x0023914Cu, 8192u);  // 76,145,35
  const uint32_t hi = __dp4a(px, 0x00132507u, 0u);     // 7,37,19
  return (hi * 256u + lo) >> 14;
}

template <bool kAligned>
__global__ void __launch_bounds__(128)
bgr_to_y_kernel(const uint8_t* __restrict__ bgr, uint32_t w, uint32_t h,
                uint8_t* __restrict__ pyr, uint64_t slot_bytes,
                uint32_t first_slot, uint32_t pitch, uint32_t pw) {
  const uint32_t x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  const uint32_t y = blockIdx.y;
  const uint32_t f = blockIdx.z;
  if (x4 >= pw) return;
  uint32_t out = 0;
  if (y < h && x4 < w) {
    const uint8_t* row = bgr + ((uint64_t)f * h + y) * (uint64_t)w * 3u;
    if (kAligned) {  // w % 4 == 0: whole group inside the frame, words aligned
      const uint32_t* p = reinterpret_cast<const uint32_t*>(row + (uint64_t)x4 * 3u);
      const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
      const uint32_t p0 = w0;                                // b0 g0 r0 (b1)
      const uint32_t p1 = __funnelshift_r(w0, w1, 24);       // b1 g1 r1 (b2)
      const uint32_t p2 = __funnelshift_r(w1, w2, 16);       // b2 g2 r2 (b3)
      const uint32_t p3 = w2 >> 8;                           // b3 g3 r3 0
      out = y_from_packed(p0) | (y_from_packed(p1) << 8) |
            (y_from_packed(p2) << 16) | (y_from_packed(p3) << 24);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x = x4 + j;
        if (x < w) {
          const uint8_t* q = row + (uint64_t)x * 3u;
          const uint32_t px = q[0] | (q[1] << 8) | (q[2] << 16);
          out |= y_from_packed(px) << (8 * j);
        }
      }
    }
  }
  uint8_t* dst = pyr + (uint64_t)(first_slot + f) * slot_bytes + (uint64_t)y * pitch + x4;
  *reinterpret_cast<uint32_t*>(dst) = out;
}

cudaError_t launch_bgr_to_y(const uint8_t* d_bgr, uint32_t w, uint32_t h,
                            uint8_t* d_pyr, const PyrLayout& lay,
                            uint32_t first_slot, uint32_t n_frames,
                            cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const uint32_t pw = lay.w[0], ph = lay.h[0];
  dim3 block(128);
  dim3 grid((pw / 4 + 127) / 128, ph, n_frames);
  // level-0 pitch is a multiple of 128 and pw a multiple of 2^(levels-1); the
  // 4-byte store of a partial last group stays inside the pitch.
  if (w % 4 == 0 && (reinterpret_cast<uintptr_t>(d_bgr) & 3) == 0)
    bgr_to_y_kernel<true><<<grid, block, 0, st>>>(d_bgr, w, h, d_pyr + lay.off[0],
                                                  lay.slot_bytes, first_slot,
                                                  lay.pitch[0], (pw + 3) & ~3u);
  else
    bgr_to_y_kernel<false><<<grid, block, 0, st>>>(d_bgr, w, h, d_pyr + lay.off[0],
                                                   lay.slot_bytes, first_slot,
                                                   lay.pitch[0], (pw + 3) & ~3u);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// pyrDown.  One thread produces 4 output pixels of one row: it needs source
// columns 2*x0-2 .. 2*x0+8 of 5 source rows.  Interior threads fetch them as
// four aligned words per row; threads at the left/right border take the
// byte path with cv::borderInterpolate(BORDER_REFLECT_101) semantics.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(128)
pyr_down_kernel(uint8_t* __restrict__ pyr, uint64_t slot_bytes,
                uint32_t first_slot, uint64_t src_off, uint32_t sw, uint32_t sh,
                uint32_t spitch, uint64_t dst_off, uint32_t dw, uint32_t dh,
                uint32_t dpitch) {
  const uint32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  const uint32_t oy = blockIdx.y;
  if (x0 >= dw) return;
  uint8_t* slot = pyr + (uint64_t)(first_slot + blockIdx.z) * slot_bytes;
  const uint8_t* src = slot + src_off;
  const bool interior = (x0 >= 2u) && (2u * x0 + 12u <= sw);
  int acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int wv = (k == 0 || k == 4) ? 1 : ((k == 2) ? 6 : 4);
    const int sy = reflect101((int)oy * 2 + k - 2, (int)sh);
    const uint8_t* row = src + (uint64_t)sy * spitch;
    int p[11];
    if (interior) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(row + 2u * x0 - 4u);
      const uint32_t a = q[0], b = q[1], c = q[2], d = q[3];
      p[0] = (a >> 16) & 0xff;  p[1] = a >> 24;
      p[2] = b & 0xff;  p[3] = (b >> 8) & 0xff;  p[4] = (b >> 16) & 0xff;  p[5] = b >> 24;
      p[6] = c & 0xff;  p[7] = (c >> 8) & 0xff;  p[8] = (c >> 16) & 0xff;  p[9] = c >> 24;
      p[10] = d & 0xff;
    } else {
#pragma unroll
      for (int j = 0; j < 11; ++j)
        p[j] = row[reflect101((int)(2u * x0) - 2 + j, (int)sw)];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      acc[i] += wv * (p[2 * i] + 4 * p[2 * i + 1] + 6 * p[2 * i + 2] +
                      4 * p[2 * i + 3] + p[2 * i + 4]);
  }
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) out |= (uint32_t)((acc[i] + 128) >> 8) << (8 * i);
  *reinterpret_cast<uint32_t*>(slot + dst_off + (uint64_t)oy * dpitch + x0) = out;
}

cudaError_t launch_pyr_down(uint8_t* d_pyr, const PyrLayout& lay,
                            uint32_t src_level, uint32_t first_slot,
                            uint32_t n_frames, cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const uint32_t l = src_level;
  const uint32_t dw = lay.w[l + 1], dh = lay.h[l + 1];
  dim3 block(128);
  dim3 grid(((dw + 3) / 4 + 127) / 128, dh, n_frames);
  pyr_down_kernel<<<grid, block, 0, st>>>(d_pyr, lay.slot_bytes, first_slot,
                                          lay.off[l], lay.w[l], lay.h[l],
                                          lay.pitch[l], lay.off[l + 1], dw, dh,
                                          lay.pitch[l + 1]);
  return cudaGetLastError();
}

}  // namespace svc
