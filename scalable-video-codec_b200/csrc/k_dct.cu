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
// Generic path (any transform block up to 32x32): separable matrix form with
// the basis in constant memory, plus a gather kernel that reproduces the
// reference serializer index arithmetic exactly (including its use of the
// unpadded width as row stride, libs/encoder.cpp:257-262).
#include <math.h>

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

__device__ __forceinline__ void dct8(float& x0, float& x1, float& x2, float& x3,
                                     float& x4, float& x5, float& x6, float& x7) {
  const float s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
  const float d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
  const float e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
  x0 = SVC_C4 * (e0 + e1);
  x4 = SVC_C4 * (e0 - e1);
  x2 = fmaf(SVC_B2, e2, SVC_B6 * e3);
  x6 = fmaf(SVC_B6, e2, -SVC_B2 * e3);
  x1 = fmaf(SVC_A, d0, fmaf(SVC_B, d1, fmaf(SVC_C, d2, SVC_D * d3)));
  x3 = fmaf(SVC_B, d0, fmaf(-SVC_D, d1, fmaf(-SVC_A, d2, -SVC_C * d3)));
  x5 = fmaf(SVC_C, d0, fmaf(-SVC_A, d1, fmaf(SVC_D, d2, SVC_B * d3)));
  x7 = fmaf(SVC_D, d0, fmaf(-SVC_C, d1, fmaf(SVC_B, d2, -SVC_A * d3)));
}

// byte `b` of `w` -> exact float, via the 2^23 mantissa trick (PRMT + FADD)
__device__ __forceinline__ float byte_to_float(uint32_t w, int b) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t)b)) - 8388608.0f;
}

constexpr int kDctWarps = 4;          // warps per CTA
constexpr int kRecWords8 = 193;       // 772-byte record = 1 + 3*64 words
constexpr int kDctSmemStream = kDctWarps * 32 * kRecWords8 * 4;

enum { kModeStream = 0, kModePlanar = 1 };

// One lane = one 8x8 block (all three channels).  A warp covers 32 consecutive
// blocks in serializer order (row-major over the nbx x nby block grid).
template <int kMode>
__global__ void __launch_bounds__(kDctWarps * 32)
dct8x8_kernel(DctParams p, uint32_t nbx, uint32_t nby) {
  extern __shared__ __align__(128) uint32_t smem[];
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t per_frame = nbx * nby;
  const uint32_t chunks_per_frame = (per_frame + 31u) / 32u;
  const uint64_t unit = (uint64_t)blockIdx.x * kDctWarps + wib;
  if (unit >= (uint64_t)chunks_per_frame * p.n_frames) return;
  const uint32_t f = (uint32_t)(unit / chunks_per_frame);
  const uint32_t n0 = (uint32_t)(unit % chunks_per_frame) * 32u;
  const uint32_t n = n0 + lane;
  const bool active = n < per_frame;
  const uint32_t tbx = active ? n % nbx : 0u, tby = active ? n / nbx : 0u;
  const uint32_t px = tbx * 8u, py = tby * 8u;

  // ---- load 8 rows x 24 bytes (zero outside the unpadded frame) --------------
  uint32_t raw[8][6];
  const uint8_t* fr = p.bgr + (uint64_t)f * p.h * p.w * 3u;
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

  uint32_t* rec = smem + (wib * 32u + lane) * kRecWords8;  // stream mode staging
  if (kMode == kModeStream) {
    uint32_t bt = 0;
    if (p.block_types && active)
      bt = __ldg(p.block_types + (uint64_t)f * p.mv_field_w * p.mv_field_h +
                 (py / p.mv_block_h) * p.mv_field_w + px / p.mv_block_w);
    rec[0] = bt;
  }

#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[r][j] = byte_to_float(raw[r][(3 * j + c) >> 2], (3 * j + c) & 3);
      dct8(v[r][0], v[r][1], v[r][2], v[r][3], v[r][4], v[r][5], v[r][6], v[r][7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);

    if (kMode == kModeStream) {
      // word (lane*193 + const) -> bank (lane + const) % 32: conflict free
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) rec[1 + c * 64 + u * 8 + j] = __float_as_uint(v[u][j]);
    } else if (active) {
      float* pl = p.planes + (((uint64_t)f * 3u + c) * p.ph + py) * (uint64_t)p.pw + px;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float4* o = reinterpret_cast<float4*>(pl + (uint64_t)u * p.pw);
        o[0] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
        o[1] = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
      }
    }
  }

  if (kMode == kModeStream) {
    const uint32_t n_act = min(32u, per_frame - n0);
    const uint32_t bytes = n_act * kRecWords8 * 4u;
    uint8_t* dst = p.stream + (uint64_t)f * p.frame_stream_bytes + (uint64_t)n0 * (kRecWords8 * 4u);
    const uint32_t* src = smem + wib * 32u * kRecWords8;
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
    if (bulk_ok) {
      // make the generic-proxy smem writes visible to the async proxy, then one
      // elected lane hands the whole span to the TMA engine.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        const uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(src);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst), "r"(s_addr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    } else {
      __syncwarp();
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
      for (uint32_t k = lane; k < bytes / 4u; k += 32u) d32[k] = src[k];
    }
  }
}

// ---- generic separable path -------------------------------------------------
__constant__ float c_basis_w[32 * 32];
__constant__ float c_basis_h[32 * 32];

__global__ void __launch_bounds__(256)
dct_rows_kernel(const uint8_t* __restrict__ bgr, uint32_t w, uint32_t h,
                uint32_t pw, uint32_t ph, uint32_t tbw, float* __restrict__ tmp,
                uint64_t total) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t x = (uint32_t)(i % pw);
  const uint32_t y = (uint32_t)((i / pw) % ph);
  const uint32_t c = (uint32_t)((i / ((uint64_t)pw * ph)) % 3u);
  const uint64_t f = i / ((uint64_t)pw * ph * 3u);
  const uint32_t x0 = x / tbw * tbw, k = x - x0;
  float s = 0.f;
  if (y < h) {
    const uint8_t* row = bgr + ((f * h + y) * (uint64_t)w) * 3u + c;
    for (uint32_t j = 0; j < tbw; ++j)
      if (x0 + j < w) s = fmaf(c_basis_w[k * tbw + j], (float)__ldg(row + (uint64_t)(x0 + j) * 3u), s);
  }
  tmp[i] = s;
}

__global__ void __launch_bounds__(256)
dct_cols_kernel(const float* __restrict__ tmp, uint32_t pw, uint32_t ph,
                uint32_t tbh, float* __restrict__ out, uint64_t total) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t x = (uint32_t)(i % pw);
  const uint32_t y = (uint32_t)((i / pw) % ph);
  const uint64_t plane = i / ((uint64_t)pw * ph);
  const uint32_t y0 = y / tbh * tbh, k = y - y0;
  const float* col = tmp + plane * (uint64_t)pw * ph + (uint64_t)y0 * pw + x;
  float s = 0.f;
  for (uint32_t j = 0; j < tbh; ++j) s = fmaf(c_basis_h[k * tbh + j], col[(uint64_t)j * pw], s);
  out[i] = s;
}

// SerializeEncodedFrame, libs/encoder.cpp:243-266, one thread per stream word.
__global__ void __launch_bounds__(256)
serialize_gather_kernel(const float* __restrict__ planes, uint64_t plane_elems,
                        const uint32_t* __restrict__ block_types,
                        uint32_t w, uint32_t h, uint32_t tbw, uint32_t tbh,
                        uint32_t mbw, uint32_t mbh, uint32_t mvw, uint32_t mvh,
                        uint32_t* __restrict__ stream, uint64_t frame_words,
                        uint64_t total) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint64_t f = i / frame_words;
  const uint64_t wi = i % frame_words;
  const uint32_t rec_words = 1u + 3u * tbw * tbh;
  const uint32_t nbx = (w + tbw - 1) / tbw;
  const uint32_t rec = (uint32_t)(wi / rec_words), k = (uint32_t)(wi % rec_words);
  const uint32_t tb_x = (rec % nbx) * tbw, tb_y = (rec / nbx) * tbh;
  uint32_t out;
  if (k == 0) {
    out = block_types ? block_types[f * mvw * mvh + (tb_y / mbh) * mvw + tb_x / mbw] : 0u;
  } else {
    const uint32_t e = k - 1u, c = e / (tbw * tbh), rem = e % (tbw * tbh);
    const uint32_t row = rem / tbh, j = rem % tbh;  // tbw rows of tbh floats (sic)
    const uint64_t idx = (uint64_t)(tb_y + row) * w + tb_x + j;  // unpadded stride (sic)
    out = idx < plane_elems
              ? __float_as_uint(planes[(f * 3u + c) * plane_elems + idx])
              : 0u;
  }
  stream[i] = out;
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
    dct8x8_kernel<kModePlanar><<<(uint32_t)((units + kDctWarps - 1) / kDctWarps), kDctWarps * 32, 0, st>>>(q, nbx, nby);
    if (nl) *nl += 1;
    return cudaGetLastError();
  }
  float hw[32 * 32], hh[32 * 32];
  host_basis(p.tbw, hw);
  host_basis(p.tbh, hh);
  cudaError_t e = cudaMemcpyToSymbolAsync(c_basis_w, hw, sizeof(float) * p.tbw * p.tbw, 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_basis_h, hh, sizeof(float) * p.tbh * p.tbh, 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  // hw/hh live on this stack frame: finish the copies before returning
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  const uint64_t total = (uint64_t)nf * 3u * p.pw * p.ph;
  const uint32_t blocks = (uint32_t)((total + 255) / 256);
  dct_rows_kernel<<<blocks, 256, 0, st>>>(bgr, p.w, p.h, p.pw, p.ph, p.tbw, tmp, total);
  dct_cols_kernel<<<blocks, 256, 0, st>>>(tmp, p.pw, p.ph, p.tbh, planes, total);
  if (nl) *nl += 2;
  return cudaGetLastError();
}

cudaError_t prepare_dct_kernels() {
  return cudaFuncSetAttribute(dct8x8_kernel<kModeStream>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kDctSmemStream);
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
    if (fast8 && p.w == p.pw) {
      const uint32_t nbx = p.w / 8, nby = (p.h + 7) / 8;
      const uint64_t units = (uint64_t)((nbx * nby + 31) / 32) * p.n_frames;
      dct8x8_kernel<kModeStream><<<(uint32_t)((units + kDctWarps - 1) / kDctWarps), kDctWarps * 32, kDctSmemStream, st>>>(p, nbx, nby);
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
        const uint64_t total = frame_words * nf;
        serialize_gather_kernel<<<(uint32_t)((total + 255) / 256), 256, 0, st>>>(
            sp, plane_elems,
            p.block_types ? p.block_types + (uint64_t)f0 * p.mv_field_w * p.mv_field_h : nullptr,
            p.w, p.h, p.tbw, p.tbh, p.mv_block_w, p.mv_block_h, p.mv_field_w, p.mv_field_h,
            reinterpret_cast<uint32_t*>(p.stream + (uint64_t)f0 * p.frame_stream_bytes),
            frame_words, total);
        if (nl) *nl += 1;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
      }
    }
  }
  return cudaSuccess;
}

}  // namespace svc
