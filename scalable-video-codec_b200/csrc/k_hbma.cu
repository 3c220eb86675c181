// k_hbma.cu -- K2: hierarchical block-matching motion estimation (MAD criterion).
//
// Device implementation of EstimateMotionHierarchical /
// EstimateMotionHierarchical16x16Sse2 / EstimateMotionExhaustiveSearch
// (reference libs/motion.cpp:268-465, 691-749).  Bit-exact with the reference,
// including its scan-order semantics:
//   * top level (libs/motion.cpp:312-337): "<=" -> the LAST minimum in raster
//     order wins; if every candidate updated the minimum (the SAD sequence is
//     non-increasing in scan order) the vector is reset to (0,0) while the
//     minimum is kept;
//   * refinement levels (:364-409): window centred on anchor + 2*mv, clamped
//     to the frame; strict "<" against the MAD carried from the coarser level,
//     so the FIRST candidate reaching a new minimum wins, and a coarser vector
//     survives when nothing is strictly better;
//   * MAD = (float)sad / (float)(bw*bh) (:38-40), IEEE division.
// Within one level integer SAD order equals float MAD order (sad < 2^23), so
// the argmin runs on integers and only the winner is converted.
#include <float.h>

#include "common.cuh"

namespace svc {

// ---------------------------------------------------------------------------
// Generic kernel: one warp per MV block, lanes stride over the candidates of
// the current level, every lane computes whole-block SADs straight from
// global memory (L1/L2 resident).  Any level count, block shape and range.
// This is the universal path; the tiled fast path below covers the encoder's
// default configuration.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_sad(const uint8_t* __restrict__ T,
                                              const uint8_t* __restrict__ A,
                                              uint32_t pitch, uint32_t tx,
                                              uint32_t ty, uint32_t ax,
                                              uint32_t ay, uint32_t bw,
                                              uint32_t bh) {
  uint32_t sad = 0;
  if ((bw & 3u) == 0) {
    // anchor rows are word aligned (ax % bw == 0); tracked rows are realigned
    // from two aligned words with a funnel shift.
    const uint32_t sh = (tx & 3u) * 8u;
    for (uint32_t k = 0; k < bh; ++k) {
      const uint32_t* a = reinterpret_cast<const uint32_t*>(A + (uint64_t)(ay + k) * pitch + ax);
      const uint32_t* t = reinterpret_cast<const uint32_t*>(T + (uint64_t)(ty + k) * pitch + (tx & ~3u));
      uint32_t lo = __ldg(t);
      for (uint32_t j = 0; j < bw / 4; ++j) {
        const uint32_t hi = __ldg(t + j + 1);
        sad = __vsadu4(__funnelshift_r(lo, hi, sh), __ldg(a + j)) + sad;
        lo = hi;
      }
    }
  } else {
    for (uint32_t k = 0; k < bh; ++k) {
      const uint8_t* a = A + (uint64_t)(ay + k) * pitch + ax;
      const uint8_t* t = T + (uint64_t)(ty + k) * pitch + tx;
      for (uint32_t j = 0; j < bw; ++j) sad = __sad((int)__ldg(t + j), (int)__ldg(a + j), sad);
    }
  }
  return sad;
}

__global__ void __launch_bounds__(256)
hbma_generic_kernel(HbmaParams p) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t per_frame = p.mvw * p.mvh;
  if (warp >= (uint64_t)per_frame * p.n_frames) return;
  const uint32_t f = (uint32_t)(warp / per_frame);
  const uint32_t i = (uint32_t)(warp % per_frame);
  const uint32_t bx = i % p.mvw, by = i / p.mvw;
  const uint8_t* Tslot = p.pyr + (uint64_t)f * p.lay.slot_bytes;
  const uint8_t* Aslot = Tslot + p.lay.slot_bytes;
  const int r = (int)p.r;

  int mx = 0, my = 0;
  float cur = FLT_MAX;
  for (int l = (int)p.lay.levels - 1; l >= 0; --l) {
    const bool top = (l == (int)p.lay.levels - 1);
    const uint32_t bw = p.bw >> l, bh = p.bh >> l;
    const uint32_t fw = p.lay.w[l], fh = p.lay.h[l], pitch = p.lay.pitch[l];
    const uint8_t* T = Tslot + p.lay.off[l];
    const uint8_t* A = Aslot + p.lay.off[l];
    if (!top) { mx *= 2; my *= 2; }
    const int ax = (int)(bx * bw), ay = (int)(by * bh);
    const int cx = ax + mx, cy = ay + my;
    const int x0 = max(0, cx - r), x1 = min((int)(fw - bw + 1), cx + r + 1);
    const int y0 = max(0, cy - r), y1 = min((int)(fh - bh + 1), cy + r + 1);
    const uint32_t ncx = (uint32_t)(x1 - x0);
    const uint32_t n = ncx * (uint32_t)(y1 - y0);

    uint32_t best_s = 0xffffffffu, best_i = 0;
    bool viol = false;       // some s[i] > s[i-1] (top level only)
    uint32_t carry = 0;      // s of the candidate preceding this round's lane 0
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t ci = base + lane;
      const bool valid = ci < n;
      uint32_t s = 0xffffffffu;
      if (valid) {
        const uint32_t tx = (uint32_t)x0 + ci % ncx, ty = (uint32_t)y0 + ci / ncx;
        s = block_sad(T, A, pitch, tx, ty, (uint32_t)ax, (uint32_t)ay, bw, bh);
      }
      if (top) {
        uint32_t prev = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane == 0) prev = carry;
        if (valid && ci > 0 && s > prev) viol = true;
        carry = __shfl_sync(0xffffffffu, s, 31);
        if (valid && s <= best_s) { best_s = s; best_i = ci; }  // later index wins
      } else {
        if (valid && s < best_s) { best_s = s; best_i = ci; }   // earlier index wins
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t os = __shfl_xor_sync(0xffffffffu, best_s, o);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      const bool take = top ? (os < best_s || (os == best_s && oi > best_i))
                            : (os < best_s || (os == best_s && oi < best_i));
      if (take) { best_s = os; best_i = oi; }
    }
    const float m = __fdiv_rn((float)best_s, (float)(bw * bh));
    const int nmx = x0 + (int)(best_i % ncx) - ax, nmy = y0 + (int)(best_i / ncx) - ay;
    if (top) {
      cur = m;
      const bool any_viol = __any_sync(0xffffffffu, viol);
      mx = any_viol ? nmx : 0;
      my = any_viol ? nmy : 0;
    } else if (m < cur) {
      cur = m;
      mx = nmx;
      my = nmy;
    }
  }
  if (lane == 0) {
    const uint64_t o = (uint64_t)f * per_frame + i;
    if (p.mv) p.mv[o] = make_float2((float)mx, (float)my);
    if (p.mad) p.mad[o] = cur;
  }
}

cudaError_t launch_hbma(const HbmaParams& p, cudaStream_t st, int* n_launches) {
  if (p.n_frames == 0) return cudaSuccess;
  const uint64_t warps = (uint64_t)p.mvw * p.mvh * p.n_frames;
  const uint32_t threads = 256;
  const uint64_t blocks = (warps * 32 + threads - 1) / threads;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  hbma_generic_kernel<<<(uint32_t)blocks, threads, 0, st>>>(p);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace svc
