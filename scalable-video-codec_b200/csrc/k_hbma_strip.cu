// k_hbma_strip.cu -- K2 at the encoder default: 16x16 motion blocks, 4 pyramid levels, r = 1 at
// every level (search_range 8; reference libs/motion.cpp:268-465, 691-749).  Bit-exact with the
// reference like every other search kernel (scan-order rules: see the header of k_hbma.cu).
//
// hbma_tile_kernel (k_hbma.cu) serves a block with 2r+1 = 3 lanes, one per candidate column: every
// lane keeps the whole anchor block in registers and streams whole tracked rows, so a 16x16 block
// costs 24 shared-memory wavefronts at level 0 and the shared-memory pipe (88-90 % busy) binds; 47 %
// of its instructions run on the three coarse levels, which hold 11 % of the SADs, because every warp
// instruction there serves only 10 blocks.
//
// The kernels here keep the bounded-reach tile (the search window of a tile of blocks at level l is
// the tile grown by d_l = 2^(L-l) - 1 pixels: ONE TMA tensor load per level and tile) and change who
// does what:
//   * a lane owns a vertical STRIP of a block and all nine candidates of it: 8-byte strips of the
//     16x16 blocks and 4-byte strips of the 8x8 blocks (2 lanes per block, 16 blocks per warp), whole
//     4x4 and 2x2 blocks (1 lane per block, 32 blocks per warp).  A tracked row is read once per
//     strip (4 LDS.32 for 10 bytes), realigned to the lane's byte phase with 3 funnel shifts, and the
//     words of the candidate columns dx = +1, +2 are produced on the FMA pipe (multiplications by
//     2^24 / 2^16 the compiler cannot fold back into SHF): 21 integer-ALU instructions per 18 SADs;
//     the nine partial SADs of the two strips meet in one butterfly step;
//   * the anchor strip of a lane comes straight from global memory (coalesced 128-byte rows, issued
//     before the TMA wait): no anchor tile in shared memory, 35 KB per CTA, 6 CTAs per SM;
//   * one mbarrier per level: a level starts as soon as its own window has landed;
//   * argmin and the top level's "every candidate updated the minimum => zero vector" rule
//     (libs/motion.cpp:312-337) run inside a lane; interior warps skip the frame-clamp tests.
// Two launches: hbma_strip_coarse_kernel (levels 3, 2: one lane per block, 64 threads per 8 x 8 tile)
// leaves the level-2 vector and MAD in p.mv / p.mad -- the values the reference carries between its
// RefineHierMotionEst calls (libs/motion.cpp:451-464) -- and hbma_strip_fine_kernel (levels 1, 0: 128
// threads per 8 x 8 tile) refines them in place.  A single kernel with warps 0-1 on the coarse levels
// was measured first (bit-exact, 118 us per 100 pairs of 1080p against 133 us for hbma_tile_kernel on
// the same box): its CTAs spent half their life in the serial chain window wait -> 2x2 level -> 4x4
// level with two of four warps parked at a barrier.  Split: 18 + 89 us (A/B on one box, tools/ab_hbma.py:
// 112 us against 133 us).  ncu of the fine kernel (profiles/r02_ncu_hbma_strip.txt): LSU data pipe 72 %
// (4 LDS.32 per strip row at 2.7 wavefronts each -- strips start on even words, so an instruction
// reaches 16 of the 32 banks, and the two block rows of a warp lie 16 rows = a multiple of 128 bytes
// apart), ALU pipe 61 %, issue slots 64 %; time moved by < 1 % when the dx shifts went from the ALU
// to the FMA pipe and by 7 % from 6 to 4 CTAs per SM: bound by shared-memory wavefronts.
#include <cuda.h>
#include <float.h>

#include "common.cuh"
#include "hbma_dev.cuh"

namespace svc {

constexpr int kSL = 4;   // pyramid levels
struct HbmaStripMaps {
  CUtensorMap t[kSL];  // tracked-window box per level
};

namespace {

constexpr int kSTB = 8;  // tile edge in motion blocks

__host__ __device__ constexpr int sg_align(int v, int a) { return (v + a - 1) / a * a; }

struct StripGeom {
  __host__ __device__ static constexpr int b(int l) { return 16 >> l; }
  __host__ __device__ static constexpr int d(int l) { return (1 << (kSL - l)) - 1; }
  // box origin floored to 16 bytes in x (TMA), hence up to 15 extra bytes
  __host__ __device__ static constexpr int tw(int l) { return sg_align(kSTB * b(l) + 2 * d(l) + 15, 16); }
  __host__ __device__ static constexpr int th(int l) { return kSTB * b(l) + 2 * d(l); }
  // + 16: the last aligned word a strip reads may lie past the bytes it needs
  __host__ __device__ static constexpr int off(int l) {
    int o = 0;
    for (int i = 0; i < l; ++i) o += sg_align(tw(i) * th(i) + 16, 128);
    return o;
  }
};

__device__ __forceinline__ void mbar_wait(uint32_t bar_addr) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
  }
}

// Multipliers the compiler cannot see through (levels == 4 is checked on the host): a shift written as a
// multiplication by a visible power of two is turned back into SHF / LEA, i.e. put on the integer-ALU
// pipe that VABSDIFF4 already keeps busy, while the FMA pipe (IMAD) idles.
struct PipeConsts {
  uint32_t c8, c16, c24;  // 2^8, 2^16, 2^24
};

// (lo >> s) | (hi << (32 - s)) as two IMADs, m = 2^(32 - s)
__device__ __forceinline__ uint32_t shr_fma(uint32_t lo, uint32_t hi_times_m, uint32_t m) {
  uint32_t d;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(m), "r"(hi_times_m));
  return d;
}

// Nine SADs (index dy * 3 + dx) of an NR-row, 4*NW-byte anchor strip `a` against the tracked rows
// starting at `trow` (row 0 = candidate row dy = 0, aligned word holding the first byte of candidate
// column dx = 0), `sh` = 8 * (byte offset of that first byte in its word).  Per row: NW + 2 aligned
// words, NW + 1 funnel shifts to the lane's byte phase, then the words of candidate columns dx = 1, 2
// by multiplication (FMA pipe) and 9 * NW VABSDIFF4.
template <int NW, int NR, int PT>
__device__ __forceinline__ void strip_sad9(const uint8_t* __restrict__ trow, const uint32_t sh,
                                           const uint32_t (&a)[NR][NW], const PipeConsts& pc,
                                           uint32_t (&acc)[9]) {
#pragma unroll
  for (int t = 0; t < NR + 2; ++t) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(trow + t * PT);
    uint32_t w[NW + 2];
#pragma unroll
    for (int k = 0; k < NW + 2; ++k) w[k] = q[k];
    uint32_t u[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; ++k) u[k] = __funnelshift_r(w[k], w[k + 1], sh);
    uint32_t x[3][NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      x[0][k] = u[k];
      x[1][k] = shr_fma(u[k], u[k + 1] * pc.c24, pc.c24);
      x[2][k] = shr_fma(u[k], u[k + 1] * pc.c16, pc.c16);
    }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ar = t - dy;
      if (ar >= 0 && ar < NR) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int k = 0; k < NW; ++k) acc[dy * 3 + dx] = sad4_acc(x[dx][k], a[ar][k], acc[dy * 3 + dx]);
      }
    }
  }
}

__device__ __forceinline__ void count_work(const HbmaParams& p, int cx, int cy, int fw, int fh, int B) {
  const int nx = min(fw - B + 1, cx + 2) - max(0, cx - 1), ny = min(fh - B + 1, cy + 2) - max(0, cy - 1);
  atomicAdd(p.counters, (unsigned long long)(nx * ny));
  atomicAdd(p.counters + 1, (unsigned long long)(nx * ny) * B * B);
}

// First minimum in scan order over the candidates inside the frame, strict "<" against the MAD
// carried from the coarser level (libs/motion.cpp:381-405).  `interior`: warp-uniform, every
// candidate of every lane lies inside the frame.
template <int B>
__device__ __forceinline__ void refine_select(const uint32_t (&acc)[9], const bool interior, const int cx,
                                              const int cy, const int fw, const int fh, const uint32_t c8,
                                              int& mx, int& my, float& cur) {
  uint32_t key = 0xffffffffu;
  if (interior) {
#pragma unroll
    for (int i = 0; i < 9; ++i) key = min(key, acc[i] * c8 + (uint32_t)i);
  } else {
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const bool yok = (uint32_t)(cy - 1 + dy) <= (uint32_t)(fh - B);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const bool ok = yok && (uint32_t)(cx - 1 + dx) <= (uint32_t)(fw - B);
        if (ok) key = min(key, acc[dy * 3 + dx] * 256u + (uint32_t)(dy * 3 + dx));
      }
    }
  }
  if (key != 0xffffffffu) {
    const float m = (float)(key >> 8) * (1.0f / (float)(B * B));
    if (m < cur) {
      const int idx = (int)(key & 0xffu);
      const int dy = (idx * 11) >> 5;  // idx / 3 for 0..8
      cur = m;
      mx += idx - dy * 3 - 1;
      my += dy - 1;
    }
  }
}

// One refinement level with LPB lanes per block (strip width B / LPB bytes = 4 * NW).
template <int LV, int LPB>
__device__ __forceinline__ void strip_refine(const uint8_t* smem, const HbmaParams& p,
                                             const uint32_t (&a)[16 >> LV][(16 >> LV) / LPB / 4],
                                             const int tile_bx0, const int tile_by0, const int lbx,
                                             const int lby, const int q, const bool owner,
                                             const PipeConsts& pc, int& mx, int& my, float& cur) {
  constexpr int B = 16 >> LV, D = StripGeom::d(LV), PT = StripGeom::tw(LV);
  constexpr int SW = B / LPB, NW = SW / 4;
  const uint8_t* sT = smem + StripGeom::off(LV);
  const int fw = (int)p.lay.w[LV], fh = (int)p.lay.h[LV];
  mx *= 2;
  my *= 2;
  const int cx = (tile_bx0 + lbx) * B + mx, cy = (tile_by0 + lby) * B + my;
  const int sx = cx - 1 + q * SW - ((tile_bx0 * B - D) & ~15);
  const int sy = cy - 1 - (tile_by0 * B - D);
  uint32_t acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0;
  strip_sad9<NW, B, PT>(sT + sy * PT + (sx & ~3), (uint32_t)(sx & 3) * 8u, a, pc, acc);
  if constexpr (LPB == 2) {
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
  }
  if (p.counters && owner) count_work(p, cx, cy, fw, fh, B);
  const bool inside = fw >= B + 2 && fh >= B + 2 && (uint32_t)(cx - 1) <= (uint32_t)(fw - B - 2) &&
                      (uint32_t)(cy - 1) <= (uint32_t)(fh - B - 2);
  refine_select<B>(acc, __all_sync(0xffffffffu, inside), cx, cy, fw, fh, pc.c8, mx, my, cur);
}

// L3 exhaustive search of one 2x2 block (one lane), "<=" scan of libs/motion.cpp:312-337
__device__ __forceinline__ void top_level_2x2(const uint8_t* smem, const HbmaParams& p, const uint32_t (&a3)[2],
                                              const int tile_bx0, const int tile_by0, const int cbx, const int cby,
                                              const bool cactive, const PipeConsts& pc, int& mx, int& my,
                                              float& cur) {
  using Gm = StripGeom;
  constexpr int PT3 = Gm::tw(3);
  const int fw3 = (int)p.lay.w[3], fh3 = (int)p.lay.h[3];
  const int cx = (tile_bx0 + cbx) * 2, cy = (tile_by0 + cby) * 2;
  const int sx = cx - 1 - ((tile_bx0 * 2 - Gm::d(3)) & ~15), sy = cy - 1 - (tile_by0 * 2 - Gm::d(3));
  uint32_t acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0;
  {
    const uint8_t* trow = smem + Gm::off(3) + sy * PT3 + (sx & ~3);
    const uint32_t sh = (uint32_t)(sx & 3) * 8u;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint32_t* qq = reinterpret_cast<const uint32_t*>(trow + t * PT3);
      const uint32_t v = __funnelshift_r(qq[0], qq[1], sh);  // bytes cx-1 .. cx+2 of the row
      const uint32_t c[3] = {v & 0xffffu, __byte_perm(v, 0u, 0x4421), v >> 16};
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int ar = t - dy;
        if (ar >= 0 && ar < 2) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) acc[dy * 3 + dx] = sad4_acc(c[dx], a3[ar], acc[dy * 3 + dx]);
        }
      }
    }
  }
  if (p.counters && cactive) count_work(p, cx, cy, fw3, fh3, 2);
  mx = my = 0;
  const bool inside3 = fw3 >= 4 && fh3 >= 4 && (uint32_t)(cx - 1) <= (uint32_t)(fw3 - 4) &&
                       (uint32_t)(cy - 1) <= (uint32_t)(fh3 - 4);
  if (__all_sync(0xffffffffu, inside3)) {
    // all nine candidates of every lane inside the frame: the last minimum as a packed key, and
    // "every candidate updated the minimum" <=> the SADs never increase along the scan order
    uint32_t key = 0xffffffffu;
    bool viol = false;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      key = min(key, acc[i] * pc.c8 + (uint32_t)(255 - i));
      if (i > 0) viol |= acc[i] > acc[i - 1];
    }
    cur = (float)(key >> 8) * 0.25f;
    if (viol) {
      const int bi = 255 - (int)(key & 0xffu);
      const int dy = (bi * 11) >> 5;
      mx = bi - dy * 3 - 1;
      my = dy - 1;
    }
  } else {
    uint32_t best = 0xffffffffu;
    int bi = 4, upd = 0, nv = 0;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const bool yok = (uint32_t)(cy - 1 + dy) <= (uint32_t)(fh3 - 2);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const bool ok = yok && (uint32_t)(cx - 1 + dx) <= (uint32_t)(fw3 - 2);
        if (ok) {
          ++nv;
          if (acc[dy * 3 + dx] <= best) {  // the last minimum wins
            best = acc[dy * 3 + dx];
            bi = dy * 3 + dx;
            ++upd;
          }
        }
      }
    }
    cur = (float)best * 0.25f;
    if (upd != nv) {  // otherwise every candidate updated the minimum: zero vector, minimum kept
      const int dy = (bi * 11) >> 5;
      mx = bi - dy * 3 - 1;
      my = dy - 1;
    }
  }
}

// thread 0: barriers [l_lo, l_hi] and the TMA loads of those levels (coarsest first).  The caller puts a
// __syncthreads() between this and the first mbar_wait of any other thread.
__device__ __forceinline__ void issue_windows(const HbmaStripMaps& maps, uint8_t* smem, const uint32_t bar0,
                                              const int l_lo, const int l_hi, const int tile_bx0,
                                              const int tile_by0, const int f) {
  using Gm = StripGeom;
#pragma unroll
  for (int l = 0; l < kSL; ++l)
    if (l >= l_lo && l <= l_hi) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * l));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
  for (int l = kSL - 1; l >= 0; --l) {
    if (l < l_lo || l > l_hi) continue;
    const int b = Gm::b(l);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + Gm::off(l));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * l),
                 "r"((uint32_t)(Gm::tw(l) * Gm::th(l))) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(&maps.t[l]), "r"((tile_bx0 * b - Gm::d(l)) & ~15), "r"(tile_by0 * b - Gm::d(l)),
        "r"(f), "r"(bar0 + 8u * l) : "memory");
  }
}

}  // namespace

// Levels 3 and 2 (2x2 and 4x4 blocks): one lane per block, 64 threads per 8 x 8 tile; the level-2
// vector and MAD go to p.mv / p.mad, exactly the values the reference carries between its
// RefineHierMotionEst calls (libs/motion.cpp:451-464).
__global__ void __launch_bounds__(64)
hbma_strip_coarse_kernel(const __grid_constant__ HbmaStripMaps maps, const HbmaParams p) {
  using Gm = StripGeom;
  // only the windows of levels 2 and 3 are staged: offsets relative to level 2
  extern __shared__ __align__(128) uint8_t smem_c[];
  uint8_t* smem = smem_c - Gm::off(2);
  __shared__ __align__(8) uint64_t bars[kSL];
  const int tile_bx0 = blockIdx.x * kSTB, tile_by0 = blockIdx.y * kSTB;
  const int f = blockIdx.z;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]);
  grid_dependency_wait();  // the pyramid levels of this batch (launch_pyr_levels) are complete
  grid_dependency_release();
  if (threadIdx.x == 0) issue_windows(maps, smem, bar0, 2, 3, tile_bx0, tile_by0, f);
  const PipeConsts pc = {p.lay.levels << 6, p.lay.levels << 14, p.lay.levels << 22};
  const int cb = threadIdx.x;
  const int cbx = cb & 7, cby = cb >> 3;
  const uint32_t bx = (uint32_t)(tile_bx0 + cbx), by = (uint32_t)(tile_by0 + cby);
  const bool cactive = bx < p.mvw && by < p.mvh;
  const uint8_t* A = p.pyr + (uint64_t)(f + 1) * p.lay.slot_bytes;
  uint32_t a2[4][1], a3[2];
  {
    const uint32_t xc = min(bx, p.mvw - 1u), yc = min(by, p.mvh - 1u);
    const uint32_t p2 = p.lay.pitch[2], p3 = p.lay.pitch[3];
    const uint8_t* r2 = A + p.lay.off[2] + (uint64_t)(yc * 4u) * p2 + xc * 4u;
    const uint8_t* r3 = A + p.lay.off[3] + (uint64_t)(yc * 2u) * p3 + xc * 2u;
#pragma unroll
    for (int k = 0; k < 2; ++k) a3[k] = (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(r3 + (uint32_t)k * p3));
#pragma unroll
    for (int k = 0; k < 4; ++k) a2[k][0] = __ldg(reinterpret_cast<const uint32_t*>(r2 + (uint32_t)k * p2));
  }
  __syncthreads();  // barrier initialisation visible to every waiter
  int mx, my;
  float cur;
  mbar_wait(bar0 + 8u * 3);
  top_level_2x2(smem, p, a3, tile_bx0, tile_by0, cbx, cby, cactive, pc, mx, my, cur);
  mbar_wait(bar0 + 8u * 2);
  strip_refine<2, 1>(smem, p, a2, tile_bx0, tile_by0, cbx, cby, 0, cactive, pc, mx, my, cur);
  if (cactive) {
    const uint64_t o = ((uint64_t)f * p.mvh + by) * p.mvw + bx;
    p.mv[o] = make_float2((float)mx, (float)my);
    p.mad[o] = cur;
  }
}

// Levels 1 and 0 (8x8 and 16x16 blocks): two lanes per block, a warp owns two block rows of the 8 x 8
// tile; starts from the level-2 result in p.mv / p.mad and overwrites it with the final one.
__global__ void __launch_bounds__(128, 6)
hbma_strip_fine_kernel(const __grid_constant__ HbmaStripMaps maps, const HbmaParams p) {
  using Gm = StripGeom;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[kSL];
  const int tile_bx0 = blockIdx.x * kSTB, tile_by0 = blockIdx.y * kSTB;
  const int f = blockIdx.z;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]);
  grid_dependency_wait();  // level-2 vectors and MADs of the coarse kernel are complete
  grid_dependency_release();
  if (threadIdx.x == 0) issue_windows(maps, smem, bar0, 0, 1, tile_bx0, tile_by0, f);
  const PipeConsts pc = {p.lay.levels << 6, p.lay.levels << 14, p.lay.levels << 22};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 1, q = lane & 1;
  const int lbx = g & 7, lby = 2 * w + (g >> 3);
  const uint32_t bx = (uint32_t)(tile_bx0 + lbx), by = (uint32_t)(tile_by0 + lby);
  const bool active = bx < p.mvw && by < p.mvh;
  // Anchor strips straight from global memory (the anchor frame is slot f + 1).  Lanes of blocks
  // outside the motion field read the last block row / column instead (never used): no predicates.
  const uint8_t* A = p.pyr + (uint64_t)(f + 1) * p.lay.slot_bytes;
  const uint32_t xb = min(bx, p.mvw - 1u), yb = min(by, p.mvh - 1u);
  const uint64_t o = ((uint64_t)f * p.mvh + yb) * p.mvw + xb;
  const float2 mv2 = p.mv[o];
  float cur = p.mad[o];
  uint32_t a0[16][2], a1[8][1];
  {
    const uint32_t p0 = p.lay.pitch[0], p1 = p.lay.pitch[1];
    const uint8_t* r0 = A + p.lay.off[0] + (uint64_t)(yb * 16u) * p0 + (xb * 16u + q * 8u);
    const uint8_t* r1 = A + p.lay.off[1] + (uint64_t)(yb * 8u) * p1 + (xb * 8u + q * 4u);
#pragma unroll
    for (int k = 0; k < 8; ++k) a1[k][0] = __ldg(reinterpret_cast<const uint32_t*>(r1 + (uint32_t)k * p1));
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(r0 + (uint32_t)k * p0));
      a0[k][0] = v.x;
      a0[k][1] = v.y;
    }
  }
  __syncthreads();  // barrier initialisation visible to every waiter
  if (!__any_sync(0xffffffffu, active)) return;  // block rows below the motion field
  int mx = active ? (int)mv2.x : 0, my = active ? (int)mv2.y : 0;
  mbar_wait(bar0 + 8u * 1);
  strip_refine<1, 2>(smem, p, a1, tile_bx0, tile_by0, lbx, lby, q, active && q == 0, pc, mx, my, cur);
  mbar_wait(bar0 + 8u * 0);
  strip_refine<0, 2>(smem, p, a0, tile_bx0, tile_by0, lbx, lby, q, active && q == 0, pc, mx, my, cur);
  if (active && q == 0) {
    const uint64_t oo = ((uint64_t)f * p.mvh + by) * p.mvw + bx;
    p.mv[oo] = make_float2((float)mx, (float)my);
    p.mad[oo] = cur;
  }
}

bool strip_supported(const HbmaParams& p) {
  return p.bw == 16 && p.bh == 16 && p.lay.levels == kSL && p.r == 1 && p.mv && p.mad && p.n_frames <= 65535 &&
         (p.mvh + kSTB - 1) / kSTB <= 65535;
}

cudaError_t launch_strip(const HbmaParams& p, cudaStream_t st, int* extra_launches) {
  using Gm = StripGeom;
  static_assert(Gm::tw(0) <= 256 && Gm::th(0) <= 256, "TMA box limit");
  HbmaStripMaps maps;
  for (int l = 0; l < kSL; ++l) {
    if (!encode_box(&maps.t[l], p.pyr + p.lay.off[l], p.lay.w[l], p.lay.h[l], p.lay.pitch[l], p.lay.slot_bytes,
                    p.n_frames + 1, Gm::tw(l), Gm::th(l)))
      return cudaErrorNotSupported;
  }
  constexpr int kFineSmem = Gm::off(2), kCoarseSmem = Gm::off(kSL) - Gm::off(2);
  dim3 grid((p.mvw + kSTB - 1) / kSTB, (p.mvh + kSTB - 1) / kSTB, p.n_frames);
  cudaError_t e = cudaFuncSetAttribute(hbma_strip_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFineSmem);
  if (e != cudaSuccess) return e;
  e = launch_dependent(hbma_strip_coarse_kernel, grid, dim3(64), kCoarseSmem, st, maps, p);
  if (e != cudaSuccess) return e;
  if (extra_launches) *extra_launches += 1;
  return launch_dependent(hbma_strip_fine_kernel, grid, dim3(128), kFineSmem, st, maps, p);
}

}  // namespace svc
