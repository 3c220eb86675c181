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
#include <cuda.h>

#include <algorithm>

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

// 16 pixels per thread: three 128-bit loads, one 128-bit store (w % 16 == 0, 16-byte aligned
// frames).  Y via two 16x8-bit dot products per pixel (dp2a), as in the fused K3 path.
__device__ __forceinline__ uint32_t luma4(uint32_t w0, uint32_t w1, uint32_t w2) {
  auto y = [](uint32_t px) {
    return __dp2a_hi(4899u, px, __dp2a_lo(1868u | (9617u << 16), px, 8192u)) >> 14;
  };
  return y(w0) | (y(__funnelshift_r(w0, w1, 24)) << 8) | (y(__funnelshift_r(w1, w2, 16)) << 16) |
         (y(w2 >> 8) << 24);
}

__global__ void __launch_bounds__(128)
bgr_to_y16_kernel(const uint8_t* __restrict__ bgr, uint32_t w, uint32_t h, uint8_t* __restrict__ pyr,
                  uint64_t slot_bytes, uint32_t first_slot, uint32_t pitch, uint32_t pw) {
  const uint32_t x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16u;
  const uint32_t y = blockIdx.y, f = blockIdx.z;
  if (x16 >= pw) return;
  uint4 out = make_uint4(0, 0, 0, 0);
  if (y < h && x16 < w) {  // w % 16 == 0: the group is entirely inside the frame
    const uint4* p = reinterpret_cast<const uint4*>(bgr + (((uint64_t)f * h + y) * w + x16) * 3u);
    const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    out.x = luma4(a.x, a.y, a.z);
    out.y = luma4(a.w, b.x, b.y);
    out.z = luma4(b.z, b.w, c.x);
    out.w = luma4(c.y, c.z, c.w);
  }
  *reinterpret_cast<uint4*>(pyr + (uint64_t)(first_slot + f) * slot_bytes + (uint64_t)y * pitch + x16) = out;
}

cudaError_t launch_bgr_to_y(const uint8_t* d_bgr, uint32_t w, uint32_t h,
                            uint8_t* d_pyr, const PyrLayout& lay,
                            uint32_t first_slot, uint32_t n_frames,
                            cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const uint32_t pw = lay.w[0], ph = lay.h[0];
  dim3 block(128);
  if (w % 16 == 0 && (reinterpret_cast<uintptr_t>(d_bgr) & 15) == 0) {
    // level-0 pitch is a multiple of 128: a 16-byte store of the last group stays inside it
    dim3 grid((((pw + 15) / 16) + 127) / 128, ph, n_frames);
    bgr_to_y16_kernel<<<grid, block, 0, st>>>(d_bgr, w, h, d_pyr + lay.off[0], lay.slot_bytes,
                                               first_slot, lay.pitch[0], (pw + 15) & ~15u);
    return cudaGetLastError();
  }
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
// pyrDown (one level per launch), shared-memory tiled.
//
// A CTA of 128 threads produces a 128 x 32 tile of the destination level.  The
// (2*128+32) x (2*32+3) source region is staged in shared memory once with
// cv::borderInterpolate(BORDER_REFLECT_101) applied at load time (one TMA bulk copy
// per tile row -- SASS UBLKCP -- plus byte patches on the borders), so the arithmetic
// below never sees a border.  Each thread then owns 4 output columns x 8 output
// rows: it walks the 19 source rows of its strip once, forms the four
// horizontal [1 4 6 4 1] sums of a row with packed-byte dot products (dp4a)
// and scatters them into the (at most three) live vertical accumulators.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  // one fold covers every coordinate a tile can ask for unless the level is tiny
  if (p < 0) p = -p; else if (p >= len) p = 2 * len - 2 - p;
  if (p < 0 || p >= len) {
    const int period = 2 * len - 2;
    p = p % period;
    if (p < 0) p += period;
    if (p >= len) p = period - p;
  }
  return p;
}

// One fold + clamp: exact for every coordinate a valid output reads (overshoot <= 2)
// once the level is at least 4 wide/high; smaller levels take pyr_down_small_kernel.
__device__ __forceinline__ int reflect101_near(int p, int len) {
  p = p < 0 ? -p : (p >= len ? 2 * len - 2 - p : p);
  return min(max(p, 0), len - 1);
}

// kRpt vertically adjacent groups of 4 output pixels.  `tcol` points 6 bytes left of the first
// source column of the group's first output (and is 4-byte aligned): with p[j] = tcol[6 + j], output i
// of a row is sum_k w[k] p[2i+k] horizontally; source rows r = 0 .. 2 kRpt + 2 at tcol + r * kPitch.
// The horizontal taps are two packed-byte dot products over the two aligned words they straddle (no
// shifts, no byte extracts); vertical accumulators hold two 16-bit columns per register: a
// [1 4 6 4 1]^2 sum is at most 255 * 256 = 65280 (+128 rounding) < 2^16, so the halves never carry.
template <int kRpt, int kPitch, bool kAlign8>  // kAlign8: tcol (and kPitch) are multiples of 8 -> one LDS.64 for p[2..9]
__device__ __forceinline__ void pyr_down_rows(const uint8_t* __restrict__ tcol, uint32_t (&out)[kRpt]) {
  constexpr int kRows = 2 * kRpt + 3;
  uint32_t acc[kRpt][2];
#pragma unroll
  for (int y = 0; y < kRpt; ++y) acc[y][0] = acc[y][1] = 0;
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(tcol + r * kPitch + 4);  // p[-2..1]
    uint32_t b0, b1;  // p[2..5], p[6..9]
    if constexpr (kAlign8) {
      const uint2 b = *reinterpret_cast<const uint2*>(tcol + r * kPitch + 8);
      b0 = b.x;
      b1 = b.y;
    } else {
      b0 = *reinterpret_cast<const uint32_t*>(tcol + r * kPitch + 8);
      b1 = *reinterpret_cast<const uint32_t*>(tcol + r * kPitch + 12);
    }
    const uint32_t c = *reinterpret_cast<const uint32_t*>(tcol + r * kPitch + 16);  // p[10..13]
    const uint32_t h0 = __dp4a(b0, 0x00010406u, __dp4a(a1, 0x04010000u, 0u));  // p[0..4]
    const uint32_t h1 = __dp4a(b1, 0x00000001u, __dp4a(b0, 0x04060401u, 0u));  // p[2..6]
    const uint32_t h2 = __dp4a(b1, 0x00010406u, __dp4a(b0, 0x04010000u, 0u));  // p[4..8]
    const uint32_t h3 = __dp4a(c, 0x00000001u, __dp4a(b1, 0x04060401u, 0u));   // p[6..10]
    const uint32_t hp0 = h1 * 65536u + h0, hp1 = h3 * 65536u + h2;
#pragma unroll
    for (int y = 0; y < kRpt; ++y) {
      const int k = r - 2 * y;  // vertical tap index for output row y
      if (k >= 0 && k <= 4) {
        const uint32_t wv = (k == 0 || k == 4) ? 1u : ((k == 2) ? 6u : 4u);
        acc[y][0] += wv * hp0;
        acc[y][1] += wv * hp1;
      }
    }
  }
#pragma unroll
  for (int y = 0; y < kRpt; ++y) {
    const uint32_t t0 = (acc[y][0] + 0x00800080u) >> 8, t1 = (acc[y][1] + 0x00800080u) >> 8;
    out[y] = __byte_perm(t0, t1, 0x6420);
  }
}

constexpr int kPdTileW = 112;                             // destination tile width (28 lanes x 4)
constexpr int kPdSrcW = 2 * kPdTileW + 32;                 // 256: cols 2*x0-16 .. 2*x0+239 = the TMA box width

// kRpt = destination rows per thread (tile height = 4 * kRpt).  8 for the big
// level-0 -> 1 launch; 2 for the small upper levels, where the launch is bound
// by the latency of one CTA rather than by throughput.
//
// Staging: ONE TMA tensor load per CTA (cp.async.bulk.tensor.3d, box 256 x (8*kRpt+3) bytes
// of the source level, out-of-image bytes zero-filled) completing on an mbarrier.  CTAs on
// the image border then rebuild BORDER_REFLECT_101 inside shared memory: whole rows first
// (row -1 <- row 1, ...), then the few border columns of every row.
template <int kRpt>
__global__ void __launch_bounds__(128)
pyr_down_kernel(const __grid_constant__ CUtensorMap src_map, uint8_t* __restrict__ pyr,
                uint64_t slot_bytes, uint32_t first_slot, uint32_t sw, uint32_t sh,
                uint64_t dst_off, uint32_t dw, uint32_t dh, uint32_t dpitch) {
  constexpr int kTileH = 4 * kRpt;
  constexpr int kSrcH = 2 * kTileH + 3;
  __shared__ __align__(128) uint8_t tile[kSrcH * kPdSrcW];
  __shared__ __align__(8) uint64_t bar;
  grid_dependency_wait();
  grid_dependency_release();
  uint8_t* slot = pyr + (uint64_t)(first_slot + blockIdx.z) * slot_bytes;
  const int x0t = blockIdx.x * kPdTileW, y0t = blockIdx.y * kTileH;
  const int sx0 = 2 * x0t - 16, sy0 = 2 * y0t - 2;  // source coords of tile[0][0] (16-byte aligned)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                 "r"((uint32_t)(kSrcH * kPdSrcW)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"(&src_map), "r"(sx0), "r"(sy0),
        "r"((int)(first_slot + blockIdx.z)), "r"(bar_addr) : "memory");
  }
  __syncthreads();  // barrier initialised before anybody polls it
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
    }
  }
  // ---- border tiles: BORDER_REFLECT_101 rebuilt in shared memory ----------------------
  {
    const int xe = min(x0t + kPdTileW, (int)dw), ye = min(y0t + kTileH, (int)dh);
    const int n_rows = 2 * (ye - y0t) + 3;                 // source rows sy0 .. sy0+n_rows-1 are read
    const int need_lo = 2 * x0t - 2, need_hi = 2 * xe;     // source columns the outputs read
    const bool row_border = sy0 < 0 || sy0 + n_rows > (int)sh;
    const bool col_border = need_lo < 0 || need_hi >= (int)sw;
    if (row_border) {  // CTA-uniform
      for (int i = threadIdx.x; i < n_rows * (kPdSrcW / 16); i += 128) {
        const int row = i / (kPdSrcW / 16), ch = i - row * (kPdSrcW / 16);
        const int sy = sy0 + row;
        if (sy < 0 || sy >= (int)sh) {
          const int from = reflect101_near(sy, (int)sh) - sy0;  // an in-image row of this tile
          *reinterpret_cast<uint4*>(tile + row * kPdSrcW + ch * 16) =
              *reinterpret_cast<const uint4*>(tile + from * kPdSrcW + ch * 16);
        }
      }
      __syncthreads();
    }
    if (col_border) {  // CTA-uniform
      for (int row = threadIdx.x; row < n_rows; row += 128) {
        uint8_t* trow = tile + row * kPdSrcW - sx0;  // trow[k] = source column k
        for (int k = need_lo; k < 0; ++k) trow[k] = trow[reflect101_near(k, (int)sw)];
        for (int k = max((int)sw, need_lo); k <= need_hi; ++k) trow[k] = trow[reflect101_near(k, (int)sw)];
      }
      __syncthreads();
    }
  }

  // ---- 4 columns x kRpt rows per thread ------------------------------------------
  const int ox = x0t + 4 * tx, oy = y0t + kRpt * ty;
  if (tx >= kPdTileW / 4 || ox >= (int)dw || oy >= (int)dh) return;
  // source columns 2*ox-2 .. 2*ox+8 live at tile columns 8*tx+14 .. 8*tx+24
  uint32_t px[kRpt];
  pyr_down_rows<kRpt, kPdSrcW, true>(tile + (2 * kRpt * ty) * kPdSrcW + 8 * tx + 8, px);
  uint8_t* drow = slot + dst_off + (uint64_t)oy * dpitch + ox;
#pragma unroll
  for (int y = 0; y < kRpt; ++y)
    if (oy + y < (int)dh) *reinterpret_cast<uint32_t*>(drow + (uint64_t)y * dpitch) = px[y];
}

// ---------------------------------------------------------------------------
// Two pyramid levels per launch: level l -> l+1 -> l+2 (the small upper levels of a batch, where a
// launch per level is bound by launch gaps and by the latency of one CTA).  A CTA owns a 24 x 8 tile
// of level l+2.  It needs the 51 x 19 region of level l+1 around it (cols 2x0-2 .. 2x0+48), which in
// turn needs 105 x 41 pixels of level l (cols 4x0-6 .. 4x0+98): ONE TMA tensor load (box 144 x 41 from
// the 16-byte aligned column 4x0-16).  Phase 1 computes the level l+1 region into shared memory (halo
// included: neighbouring CTAs recompute it) and stores its 48 x 16 core to global memory; phase 2
// computes the level l+2 tile from it.  Each level is computed from the ROUNDED previous level
// (cv::buildPyramid, libs/encoder.cpp:470), and BORDER_REFLECT_101 is applied at the edge of each
// level: level l in the TMA tile as in pyr_down_kernel, level l+1 by reflecting the out-of-image part
// of the region inside shared memory before phase 2.
// ---------------------------------------------------------------------------
constexpr int kF2TileW = 24, kF2TileH = 16;                   // level l+2 tile
constexpr int kF1W = 2 * kF2TileW + 3, kF1H = 2 * kF2TileH + 3;  // 51 x 35: level l+1 region
constexpr int kF1Groups = (kF1W + 3) / 4;                     // 13 groups of 4 columns
constexpr int kF1Rpt = 4, kF1Strips = (kF1H + kF1Rpt - 1) / kF1Rpt;  // 9 strips of 4 rows
constexpr int kF0BoxW = 144, kF0BoxH = 2 * kF1H + 3;          // 144 x 73: TMA box of level l
constexpr int kF0Rows = 2 * kF1Rpt * kF1Strips + 3;           // 75: rows the phase-1 threads may touch
constexpr int kF1Pitch = 80, kF1Off = 14;                      // level l+1 region: column j at byte j + 14 (core at byte 16)
constexpr int kF2Rpt = 2;                                      // phase 2: level l+2 rows per thread
constexpr int kFusedThreads = 128;
static_assert((kF2TileW / 4) * (kF2TileH / kF2Rpt) <= kFusedThreads, "phase 2 does not fit the CTA");
static_assert(kF1Groups * kF1Strips <= kFusedThreads, "phase 1 does not fit the CTA");
static_assert(8 * (kF1Groups - 1) + 4 + 20 <= kF0BoxW, "phase 1 reads past the box");

__global__ void __launch_bounds__(kFusedThreads)
pyr_down2_kernel(const __grid_constant__ CUtensorMap src_map, uint8_t* __restrict__ pyr, uint64_t slot_bytes,
                 uint32_t first_slot, uint32_t sw, uint32_t sh, uint64_t off1, uint32_t w1, uint32_t h1,
                 uint32_t pitch1, uint64_t off2, uint32_t w2, uint32_t h2, uint32_t pitch2) {
  __shared__ __align__(128) uint8_t tile[kF0Rows * kF0BoxW];
  __shared__ __align__(16) uint8_t t1[(kF1Rpt * kF1Strips + 1) * kF1Pitch];
  __shared__ __align__(8) uint64_t bar;
  grid_dependency_wait();
  grid_dependency_release();
  uint8_t* slot = pyr + (uint64_t)(first_slot + blockIdx.z) * slot_bytes;
  const int x0 = blockIdx.x * kF2TileW, y0 = blockIdx.y * kF2TileH;  // level l+2 tile origin
  const int u0 = 2 * x0 - 2, v0 = 2 * y0 - 2;                        // level l+1 coords of region (0, 0)
  const int sx0 = 4 * x0 - 16, sy0 = 4 * y0 - 6;                     // level l coords of tile[0][0]
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                 "r"((uint32_t)(kF0BoxH * kF0BoxW)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"(&src_map), "r"(sx0), "r"(sy0),
        "r"((int)(first_slot + blockIdx.z)), "r"(bar_addr) : "memory");
  }
  __syncthreads();
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
    }
  }
  // in-image part of the level l+1 region, and of the level l+2 tile
  const int ulo = max(u0, 0), uhi = min(u0 + kF1W, (int)w1) - 1;
  const int vlo = max(v0, 0), vhi = min(v0 + kF1H, (int)h1) - 1;
  // ---- level l: BORDER_REFLECT_101 rebuilt in shared memory (border CTAs only) --------------
  {
    const int r_lo = 2 * vlo - 2, r_hi = 2 * vhi + 2;  // level l rows / columns the valid region reads
    const int c_lo = 2 * ulo - 2, c_hi = 2 * uhi + 2;
    if (r_lo < 0 || r_hi >= (int)sh) {  // CTA-uniform
      for (int i = threadIdx.x; i < (r_hi - r_lo + 1) * (kF0BoxW / 16); i += kFusedThreads) {
        const int rr = i / (kF0BoxW / 16), ch = i - rr * (kF0BoxW / 16);
        const int sy = r_lo + rr;
        if (sy < 0 || sy >= (int)sh) {
          const int from = reflect101_near(sy, (int)sh) - sy0;
          *reinterpret_cast<uint4*>(tile + (sy - sy0) * kF0BoxW + ch * 16) =
              *reinterpret_cast<const uint4*>(tile + from * kF0BoxW + ch * 16);
        }
      }
      __syncthreads();
    }
    if (c_lo < 0 || c_hi >= (int)sw) {  // CTA-uniform
      for (int rr = threadIdx.x; rr <= r_hi - r_lo; rr += kFusedThreads) {
        uint8_t* trow = tile + (r_lo + rr - sy0) * kF0BoxW - sx0;  // trow[k] = level l column k
        for (int k = c_lo; k < 0; ++k) trow[k] = trow[reflect101_near(k, (int)sw)];
        for (int k = max((int)sw, c_lo); k <= c_hi; ++k) trow[k] = trow[reflect101_near(k, (int)sw)];
      }
      __syncthreads();
    }
  }
  // ---- phase 1: level l+1 region, 4 columns x 4 rows per thread --------------------------------
  if (threadIdx.x < kF1Groups * kF1Strips) {
    const int g = threadIdx.x % kF1Groups, sidx = threadIdx.x / kF1Groups;
    // region column j = 4g + i reads tile columns 2j + 10 .. 2j + 14 = (8g + 4) + 6 + 2i ..
    uint32_t px[kF1Rpt];
    pyr_down_rows<kF1Rpt, kF0BoxW, false>(tile + (2 * kF1Rpt * sidx) * kF0BoxW + 8 * g + 4, px);
#pragma unroll
    for (int y = 0; y < kF1Rpt; ++y) {
      const int i1 = kF1Rpt * sidx + y;  // region row
      // shared copy (the group sits at byte 4g + 14, 2-byte aligned)
      uint16_t* q = reinterpret_cast<uint16_t*>(t1 + i1 * kF1Pitch + 4 * g + kF1Off);
      q[0] = (uint16_t)px[y];
      q[1] = (uint16_t)(px[y] >> 16);
    }
  }
  __syncthreads();
  // core of the region (u = 2x0 .. 2x0+47, v = 2y0 .. 2y0+31) -> global memory: three 16-byte stores per
  // row (the core starts at byte 16 of a region row and at a multiple of 48 in the level)
  for (int e = threadIdx.x; e < 2 * kF2TileH * 3; e += kFusedThreads) {
    const int row = e / 3, ch = e - row * 3;
    const int v = 2 * y0 + row, u = 2 * x0 + 16 * ch;
    if (v < (int)h1 && u < (int)w1)
      *reinterpret_cast<uint4*>(slot + off1 + (uint64_t)v * pitch1 + u) =
          *reinterpret_cast<const uint4*>(t1 + (row + 2) * kF1Pitch + 16 + 16 * ch);
  }
  // ---- level l+1: BORDER_REFLECT_101 of the region's out-of-image part ---------------------------
  {
    const int xe = min(x0 + kF2TileW, (int)w2), ye = min(y0 + kF2TileH, (int)h2);
    const int rv_lo = 2 * y0 - 2, rv_hi = 2 * (ye - 1) + 2;  // level l+1 rows / columns the tile reads
    const int cu_lo = 2 * x0 - 2, cu_hi = 2 * (xe - 1) + 2;
    if (rv_lo < 0 || rv_hi >= (int)h1) {  // CTA-uniform
      for (int i = threadIdx.x; i < (rv_hi - rv_lo + 1) * (kF1Pitch / 4); i += kFusedThreads) {
        const int rr = i / (kF1Pitch / 4), ch = i - rr * (kF1Pitch / 4);
        const int v = rv_lo + rr;
        if (v < 0 || v >= (int)h1) {
          const int from = reflect101_near(v, (int)h1) - v0;
          *reinterpret_cast<uint32_t*>(t1 + (v - v0) * kF1Pitch + ch * 4) =
              *reinterpret_cast<const uint32_t*>(t1 + from * kF1Pitch + ch * 4);
        }
      }
      __syncthreads();
    }
    if (cu_lo < 0 || cu_hi >= (int)w1) {  // CTA-uniform
      for (int rr = threadIdx.x; rr <= rv_hi - rv_lo; rr += kFusedThreads) {
        uint8_t* trow = t1 + (rv_lo + rr - v0) * kF1Pitch + kF1Off - u0;  // trow[k] = level l+1 column k
        for (int k = cu_lo; k < 0; ++k) trow[k] = trow[reflect101_near(k, (int)w1)];
        for (int k = max((int)w1, cu_lo); k <= cu_hi; ++k) trow[k] = trow[reflect101_near(k, (int)w1)];
      }
      __syncthreads();
    }
  }
  // ---- phase 2: level l+2 tile, 4 columns x 2 rows per thread -----------------------------------
  if (threadIdx.x < (kF2TileW / 4) * (kF2TileH / kF2Rpt)) {
    const int g = threadIdx.x % (kF2TileW / 4), sidx = threadIdx.x / (kF2TileW / 4);
    const int ox = x0 + 4 * g, oy = y0 + kF2Rpt * sidx;
    if (ox < (int)w2 && oy < (int)h2) {
      // output q = 4g + i reads region columns 2q .. 2q + 4 = bytes 2q + 14 .. = (8g + 8) + 6 + 2i ..
      uint32_t px[kF2Rpt];
      pyr_down_rows<kF2Rpt, kF1Pitch, true>(t1 + (2 * kF2Rpt * sidx) * kF1Pitch + 8 * g + 8, px);
#pragma unroll
      for (int y = 0; y < kF2Rpt; ++y)
        if (oy + y < (int)h2) *reinterpret_cast<uint32_t*>(slot + off2 + (uint64_t)(oy + y) * pitch2 + ox) = px[y];
    }
  }
}

// Levels narrower or lower than 4 pixels (tiny frames): one thread per output
// pixel with the full cv::borderInterpolate reflection.
__global__ void __launch_bounds__(128)
pyr_down_small_kernel(uint8_t* __restrict__ pyr, uint64_t slot_bytes, uint32_t first_slot,
                      uint64_t src_off, uint32_t sw, uint32_t sh, uint32_t spitch,
                      uint64_t dst_off, uint32_t dw, uint32_t dh, uint32_t dpitch) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dw * dh) return;
  const int ox = (int)(i % dw), oy = (int)(i / dw);
  uint8_t* slot = pyr + (uint64_t)(first_slot + blockIdx.z) * slot_bytes;
  const uint8_t* src = slot + src_off;
  int acc = 0;
  for (int k = 0; k < 5; ++k) {
    const int wv = (k == 0 || k == 4) ? 1 : ((k == 2) ? 6 : 4);
    const uint8_t* row = src + (uint64_t)reflect101(2 * oy + k - 2, (int)sh) * spitch;
    int hsum = 0;
    for (int j = 0; j < 5; ++j) {
      const int wh = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
      hsum += wh * row[reflect101(2 * ox + j - 2, (int)sw)];
    }
    acc += wv * hsum;
  }
  slot[dst_off + (uint64_t)oy * dpitch + ox] = (uint8_t)((acc + 128) >> 8);
}

// Hand-over of one pyramid slot (the previous frame, libs/encoder.cpp:661-663) as a kernel: a
// cudaMemcpyAsync D2D may be queued on the copy engine that is busy with a 400 MB device-to-host
// transfer of the host path, which would stall the whole motion stream behind it.
__global__ void __launch_bounds__(256) copy_slot_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
  grid_dependency_wait();
  grid_dependency_release();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

cudaError_t launch_copy_slot(uint8_t* dst, const uint8_t* src, size_t bytes, cudaStream_t st) {
  // slots are 256-byte multiples at 256-byte aligned offsets of a cudaMalloc'd array
  const size_t n = bytes / 16;
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)kNumSms * 8);
  copy_slot_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n);
  return cudaGetLastError();
}

typedef CUresult (*PyrEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);

// level `l` of the slot array as a 3-D tensor (x, y, slot) with a box_w x box_h x 1 box
static cudaError_t encode_level_map(CUtensorMap* map, uint8_t* d_pyr, const PyrLayout& lay, uint32_t l,
                                    uint32_t n_slots, uint32_t box_w, uint32_t box_h) {
  static PyrEncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return cudaErrorNotSupported;
    encode = reinterpret_cast<PyrEncodeTiledFn>(fp);
  }
  const cuuint64_t dims[3] = {lay.w[l], lay.h[l], (cuuint64_t)n_slots};
  const cuuint64_t strides[2] = {lay.pitch[l], lay.slot_bytes};
  const cuuint32_t box[3] = {box_w, box_h, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  if (encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_pyr + lay.off[l], dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  return cudaSuccess;
}

cudaError_t launch_pyr_down(uint8_t* d_pyr, const PyrLayout& lay,
                            uint32_t src_level, uint32_t first_slot,
                            uint32_t n_frames, cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const uint32_t l = src_level;
  const uint32_t dw = lay.w[l + 1], dh = lay.h[l + 1];
  dim3 block(128);
  if (lay.w[l] < 4 || lay.h[l] < 4) {
    pyr_down_small_kernel<<<dim3((dw * dh + 127) / 128, 1, n_frames), block, 0, st>>>(
        d_pyr, lay.slot_bytes, first_slot, lay.off[l], lay.w[l], lay.h[l], lay.pitch[l], lay.off[l + 1],
        dw, dh, lay.pitch[l + 1]);
    return cudaGetLastError();
  }
  // 8 destination rows per thread once the launch holds >= 8 Mi output pixels (throughput-bound),
  // 2 for small launches (bound by the latency of one CTA)
  const bool big = (uint64_t)dw * dh * n_frames >= (8ull << 20);
  const int rpt = big ? 8 : 2;
  const uint32_t tile_h = 4u * (uint32_t)rpt;
  CUtensorMap map;  // box = the tile's source region
  cudaError_t e = encode_level_map(&map, d_pyr, lay, l, first_slot + n_frames, kPdSrcW, 2 * tile_h + 3);
  if (e != cudaSuccess) return e;
  dim3 grid((dw + kPdTileW - 1) / kPdTileW, (dh + tile_h - 1) / tile_h, n_frames);
#define SVC_PYR_LAUNCH(RPT)                                                                             \
  launch_dependent(pyr_down_kernel<RPT>, grid, block, 0, st, map, d_pyr, lay.slot_bytes, first_slot, lay.w[l], \
                   lay.h[l], lay.off[l + 1], dw, dh, lay.pitch[l + 1])
  return rpt == 8 ? SVC_PYR_LAUNCH(8) : SVC_PYR_LAUNCH(2);
#undef SVC_PYR_LAUNCH
}

// levels src_level+1 and src_level+2 in one launch (pyr_down2_kernel)
static cudaError_t launch_pyr_down2(uint8_t* d_pyr, const PyrLayout& lay, uint32_t src_level, uint32_t first_slot,
                                    uint32_t n_frames, cudaStream_t st) {
  const uint32_t l = src_level;
  CUtensorMap map;
  cudaError_t e = encode_level_map(&map, d_pyr, lay, l, first_slot + n_frames, kF0BoxW, kF0BoxH);
  if (e != cudaSuccess) return e;
  dim3 grid((lay.w[l + 2] + kF2TileW - 1) / kF2TileW, (lay.h[l + 2] + kF2TileH - 1) / kF2TileH, n_frames);
  return launch_dependent(pyr_down2_kernel, grid, dim3(kFusedThreads), 0, st, map, d_pyr, lay.slot_bytes, first_slot,
                          lay.w[l], lay.h[l], lay.off[l + 1], lay.w[l + 1], lay.h[l + 1], lay.pitch[l + 1],
                          lay.off[l + 2], lay.w[l + 2], lay.h[l + 2], lay.pitch[l + 2]);
}

// All levels 1 .. L-1 of slots first_slot .. first_slot+n_frames-1 from their level 0: level 0 -> 1
// on the throughput kernel, the smaller levels two at a time.
cudaError_t launch_pyr_levels(uint8_t* d_pyr, const PyrLayout& lay, uint32_t first_slot, uint32_t n_frames,
                              cudaStream_t st, int* n_launches) {
  if (n_frames == 0) return cudaSuccess;
  uint32_t l = 0;
  while (l + 1 < lay.levels) {
    cudaError_t e;
    const bool pair = l >= 1 && l + 2 < lay.levels && lay.w[l + 1] >= 4 && lay.h[l + 1] >= 4 &&
                      lay.w[l] >= 4 && lay.h[l] >= 4 && n_frames <= 65535;
    if (pair) {
      e = launch_pyr_down2(d_pyr, lay, l, first_slot, n_frames, st);
      l += 2;
    } else {
      e = launch_pyr_down(d_pyr, lay, l, first_slot, n_frames, st);
      l += 1;
    }
    if (e != cudaSuccess) return e;
    if (n_launches) *n_launches += 1;
  }
  return cudaSuccess;
}

}  // namespace svc
