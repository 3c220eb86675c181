// k_hbma_rs.cu -- K2, mid search ranges (16x16 motion blocks, top-level range r = 3..8): ONE pyramid
// level per launch with the candidate windows in shared memory exactly as TMA delivers them
// ("rs" = register shift; the level-synchronous path of k_hbma_pool.cu launches these).
//
// Same arithmetic and scan-order rules as k_hbma.cu (reference libs/motion.cpp:268-465, 691-749):
// clamped windows (:297-310, 375-385), "<=" / last-minimum + all-updates => zero vector at the top
// level (:312-337), strict "<" against the MAD carried from the coarser level below it (:401-405),
// the carried vector doubled between levels (:459).
//
// Why another layout.  For windows of at most 17 x 17 candidates the pooled kernels of
// k_hbma_pool.cu spend more time around the SAD loop than in it (ncu, r = 8, 16x16 level: 59 % of
// the instructions are VABSDIFF4, the ALU pipe is 64 % busy, a CTA waits on its TMA round trip,
// builds three shifted copies of every window behind a block-wide barrier, and only three 128-thread
// CTAs fit an SM next to 71 KB of window copies).  Here nothing is rebuilt:
//   * a work item is 4 candidate columns of ONE byte phase (columns p, p+4, p+8, p+12 of the window)
//     times a chunk of <= 6 candidate rows.  The four columns read the same aligned words of a window
//     row, so one funnel shift per word serves all four: 7 SHF per row against 16 VABSDIFF4 per row
//     and candidate row -- the realignment costs ~9 % of the SAD work and needs no copies, no build
//     phase and no barrier; the 17th column of a full window forms its own 1-column items;
//   * a window is 48 x (B + 2r) bytes (1.6 KB for the 16x16 level instead of 9.9 KB with copies), so
//     8 blocks per 128-thread CTA need 15 KB and the register file, not shared memory, limits the
//     CTAs per SM (5-6 CTAs: while one waits for its windows the others keep the ALU pipe busy);
//   * items map to lanes statically (block = lane / 12 ..): no pooled item decode;
//   * the 8 blocks of a CTA are horizontal neighbours and share ONE anchor tile (one TMA request);
//     an item keeps 4 x 6 running SADs and reads each anchor row once, when the row loop reaches it.
// Measured and rejected on the way (45 pairs of 1080p, R=64/L=4, us per level 16x16 / 8x8 / 4x4):
//   straight-line items 589 / 201 / 119 (this file: 555 / 206 / 113); 9 rows per item, 12 blocks per
//   CTA: 128 registers, 4 CTAs per SM, -13 %; a two-stage TMA pipeline over consecutive block groups
//   inside a persistent CTA: -20 %; warps uniform in the chunk index (a 5-row instance for the short
//   chunk): the 8 windows a warp then reads lie 128-byte aligned in shared memory -> bank conflicts, -20 %;
//   16 / 32 blocks per CTA for the small levels: 263 / 186; windows staged by 16-byte global loads
//   instead of TMA: 613 / 248 / 140; windows of 9 x 9 (r = 3, 4) on these kernels: 30-35 % of the SAD
//   peak against 41-50 % on the bounded-reach tile kernel, which keeps them.  ncu of the final 16x16
//   level: ALU pipe 84-87 % busy, 64 % of all instructions are VABSDIFF4, 7 % SHF
//   (profiles/r02_ncu_hbma_rs_*.txt).
#include <float.h>

#include "common.cuh"
#include "hbma_dev.cuh"

namespace svc {

namespace {

// window pitch = TMA box width.  The box starts at the 16-byte boundary below the window (<= 15 bytes of
// slack); the last candidate column starts at most at byte 15 + 2r, its aligned word at 12 + 2r (2r a
// multiple of 4 for the classes used), and an item reads B/4 + 1 words from there: 32 + 2r bytes in all.
// (64 bytes would do for r <= 16, but chunks of 6 rows at a pitch of 16 words all start in the same bank:
// the lanes of a block -- one per chunk -- would conflict 6 ways; 80 spreads them)
__host__ __device__ constexpr int rs_pitch(int rmax) { return rmax > 8 ? 80 : (32 + 2 * rmax + 15) & ~15; }  // 48 / 80

struct RsLv {        // one motion block at this level
  int x0, y0;        // origin of the clamped candidate window
  int ncx, ncy;      // its size (0 x 0: no block)
  int csz;           // candidate rows per chunk
  int sxb, aoff;     // byte offset of x0 inside the TMA box; of the anchor block inside its box
};

// NC candidate columns of one byte phase (4 bytes apart) x NDY candidate rows of a BxB anchor block.
// `trow` = first window row of the chunk at the aligned word that holds the first column, `sh` = 8 *
// (byte offset of that column inside the word).  Streams B + NDY - 1 window rows once.
//
// Written as a loop over groups of NDY anchor rows: the NDY shifted window rows an anchor row needs
// rotate through NDY register slots, which is a static renaming when the loop advances NDY anchor rows
// per trip.  The body is NDY x (one window row in, one anchor row in, NC x NDY x NW SADs): a few KB of
// code that stays in the instruction cache (the straight-line version of a 16x16 item is ~30 KB and
// measured 5 % slower: `no_instruction` stalls).
template <int B, int NC, int NDY, int PA, int PT>  // PA: pitch of the anchor tile, PT: of the window
__device__ __forceinline__ void rs_item(const uint8_t* __restrict__ trow, const uint32_t sh,
                                        const uint8_t* __restrict__ ablk, uint32_t (&acc)[NC][NDY]) {
  constexpr int NW = B >= 4 ? B / 4 : 1;
  constexpr uint32_t MASK = B >= 4 ? 0xffffffffu : (B == 2 ? 0xffffu : 0xffu);
  constexpr int NS = NW + NC - 1;
  uint32_t S[NDY][NS];
  auto load_row = [&](const uint8_t* rowp, uint32_t (&dst)[NS]) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(rowp);
    uint32_t raw[NS + 1];
#pragma unroll
    for (int k = 0; k <= NS; ++k) raw[k] = q[k];
#pragma unroll
    for (int k = 0; k < NS; ++k) dst[k] = __funnelshift_r(raw[k], raw[k + 1], sh) & MASK;
  };
#pragma unroll
  for (int d = 0; d < NDY - 1; ++d) load_row(trow + d * PT, S[d]);
#pragma unroll 1
  for (int ar0 = 0; ar0 < B; ar0 += NDY) {
    const uint8_t* tp = trow + (ar0 + NDY - 1) * PT;
    const uint8_t* ap = ablk + ar0 * PA;
#pragma unroll
    for (int u = 0; u < NDY; ++u) {
      if (ar0 + u < B) {  // uniform
        load_row(tp + u * PT, S[(u + NDY - 1) % NDY]);  // window row ar0 + u + NDY - 1
        uint32_t a[NW];
        const uint8_t* q = ap + u * PA;
        if constexpr (B == 16) {
          const uint4 v = *reinterpret_cast<const uint4*>(q);
          a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        } else if constexpr (B == 8) {
          const uint2 v = *reinterpret_cast<const uint2*>(q);
          a[0] = v.x; a[1] = v.y;
        } else if constexpr (B == 4) {
          a[0] = *reinterpret_cast<const uint32_t*>(q);
        } else if constexpr (B == 2) {
          a[0] = *reinterpret_cast<const uint16_t*>(q);
        } else {
          a[0] = *q;
        }
#pragma unroll
        for (int d = 0; d < NDY; ++d)
#pragma unroll
          for (int i = 0; i < NC; ++i)
#pragma unroll
            for (int k = 0; k < NW; ++k) acc[i][d] = sad4_acc(S[(u + d) % NDY][i + k], a[k], acc[i][d]);
      }
    }
  }
}

// keys of one item into the block's running minimum.  Top level: also returns whether the SADs of this
// item alone already increase somewhere along the scan order (later row, or later column of the same row):
// a sequence that never increases cannot contain ANY later element above an earlier one, so such a pair
// proves that not every candidate updated the minimum (libs/motion.cpp:333-337) without looking at the
// neighbouring items.
template <int NC, int NDY, bool TOP>
__device__ __forceinline__ bool rs_commit(const uint32_t (&acc)[NC][NDY], const RsLv& v, const int col0,
                                          const int dy0, const int nc, uint32_t* best) {
  const int ndy = min(v.csz, v.ncy - dy0);
  uint32_t key = 0xffffffffu;
  bool viol = false;
  uint32_t prev = 0xffffffffu;
  uint32_t idx_row = (uint32_t)(dy0 * v.ncx + col0);  // scan order inside the clamped window
#pragma unroll
  for (int d = 0; d < NDY; ++d) {
    if (d < ndy) {
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        if (i < nc && col0 + 4 * i < v.ncx) {
          const uint32_t idx = idx_row + 4u * i;
          if (TOP) {
            key = min(key, acc[i][d] * 65536u + (0xffffu - idx));  // "<=": the last minimum wins
            viol |= acc[i][d] > prev;
            prev = acc[i][d];
          } else {
            key = min(key, acc[i][d] * 65536u + idx);              // "<": the first minimum wins
          }
        }
      }
    }
    idx_row += (uint32_t)v.ncx;
  }
  if (key != 0xffffffffu) atomicMin(best, key);
  return viol;
}

template <int B, int NDY, int NB, int RMAX>
struct RsGeom {
  static constexpr int PT = rs_pitch(RMAX);
  static constexpr int NCAND = (2 * RMAX + 1) * (2 * RMAX + 1);
  static constexpr int ROWS = B + 2 * RMAX + NDY;                 // + slack rows a short last chunk streams
  static constexpr int WIN = (PT * ROWS + 127) & ~127;
  static constexpr int BLK = WIN;                                 // one window per block
  // the NB blocks of a CTA are horizontal neighbours: ONE anchor tile (one TMA request instead of NB)
  static constexpr int PA = (NB * B) % 16 == 0 ? NB * B : ((NB * B + 15 + 15) & ~15);  // (+ 16-byte origin slack)
  static constexpr int OFF_ANC = NB * BLK;
  static constexpr int ANC = (PA * B + 127) & ~127;
  static constexpr int SADS = NCAND * 2 + 2;                      // u16 per candidate (top level only)
  static constexpr int OFF_SADS = OFF_ANC + ANC;
};

// B: block size at level `lvl`; TOP: exhaustive top level (EstimateMotionExhaustiveSearch semantics) or
// refinement of the vector / MAD found in p.mv / p.mad; NC columns per item, NDY rows per chunk, NCH
// chunks per window, NB blocks per CTA; LAST: a full window has a 17th column (r = 8), searched by
// 1-column items on an extra warp.
// KG groups of NC columns per byte phase (1: windows up to 17 wide, 2: up to 33), RMAX: largest range.
template <int B, bool TOP, int NC, int NDY, int NCH, int NB, bool LAST, int MINB, int KG, int RMAX>
__global__ void __launch_bounds__(NB * 4 * KG * NCH + (LAST ? (NB * NCH + 31) / 32 * 32 : 0), MINB)
hbma_rs_kernel(const __grid_constant__ EbmaMaps maps, const __grid_constant__ HbmaParams p, const int lvl) {
  using G = RsGeom<B, NDY, NB, RMAX>;
  constexpr int kRsPT = G::PT;
  constexpr int MAIN = NB * 4 * KG * NCH;
  static_assert(MAIN % 32 == 0, "main items must fill whole warps");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ RsLv sLv[NB];
  __shared__ uint32_t sBest[NB], sViol[NB];
  const int tid = threadIdx.x;
  // CTA -> NB horizontally adjacent motion blocks (frame f, block row by, blocks bx0 .. bx0 + NB - 1)
  const uint32_t tiles_per_row = (p.mvw + NB - 1) / NB, tiles_per_frame = tiles_per_row * p.mvh;
  const uint32_t f = blockIdx.x / tiles_per_frame, ti = blockIdx.x - f * tiles_per_frame;
  const int by = (int)(ti / tiles_per_row), bx0 = (int)(ti - (uint32_t)by * tiles_per_row) * NB;
  const int r = (int)p.r;
  const int box_h = B + 2 * r;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(NB));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ---- thread j < NB owns motion block j of this CTA: window geometry, TMA loads -------------
  int mx = 0, my = 0, x0 = 0, y0 = 0, ncx = 1, ax = 0, ay = 0;
  float cur = FLT_MAX;
  uint32_t gb = 0;
  bool own = false;
  if (tid < NB) {
    const int bx = bx0 + tid;
    own = (uint32_t)bx < p.mvw;
    gb = (f * p.mvh + (uint32_t)by) * p.mvw + (uint32_t)bx;
    RsLv v{};
    if (own) {
      if (!TOP) {
        const float2 m = p.mv[gb];  // integer valued (libs/motion.cpp:326-327, 403-404)
        cur = p.mad[gb];
        mx = 2 * (int)m.x;          // motion_field *= 2 between levels (libs/motion.cpp:459)
        my = 2 * (int)m.y;
      }
      const int fw = (int)p.lay.w[lvl], fh = (int)p.lay.h[lvl];
      ax = bx * B;
      ay = by * B;
      const int cx = ax + mx, cy = ay + my;
      x0 = max(0, cx - r);
      y0 = max(0, cy - r);
      const int x1 = min(fw - B + 1, cx + r + 1), y1 = min(fh - B + 1, cy + r + 1);
      ncx = x1 - x0;
      const int ncy = y1 - y0;
      v.x0 = x0; v.y0 = y0; v.ncx = ncx; v.ncy = ncy;
      v.csz = (ncy + NCH - 1) / NCH;
      v.sxb = x0 & 15;
      v.aoff = ((bx0 * B) & 15) + tid * B;  // inside the CTA's anchor tile
      sLv[tid] = v;
      sBest[tid] = 0xffffffffu;
      sViol[tid] = 0u;
      // thread 0 (its block always exists) also fetches the anchor tile of the whole CTA
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                   "r"((uint32_t)(kRsPT * box_h + (tid == 0 ? G::PA * B : 0))) : "memory");
      uint8_t* blk = smem + tid * G::BLK;
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"((uint32_t)__cvta_generic_to_shared(blk)), "l"(&maps.t), "r"(x0 & ~15), "r"(y0), "r"((int)f),
          "r"(bar_addr) : "memory");
      if (tid == 0)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"((uint32_t)__cvta_generic_to_shared(smem + G::OFF_ANC)), "l"(&maps.a), "r"((bx0 * B) & ~15), "r"(ay),
            "r"((int)f + 1), "r"(bar_addr) : "memory");
    } else {
      sLv[tid] = v;
      sBest[tid] = 0xffffffffu;
      sViol[tid] = 0u;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
    }
  }
  {  // windows landed; sLv / sBest published (arrive = release, wait = acquire)
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
    }
  }
  // ---- work items: static lane -> (block, byte phase, chunk) ----------------------------------
  int it_j = -1, it_col0 = 0, it_dy0 = 0, it_nc = 0;  // this lane's item
  uint32_t acc[NC][NDY];
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int d = 0; d < NDY; ++d) acc[i][d] = 0;
  if (tid < MAIN) {
    const int j = tid / (4 * KG * NCH), rem = tid - j * (4 * KG * NCH);
    const int ph = rem / (KG * NCH), rem2 = rem - ph * (KG * NCH);
    const int kg = rem2 / NCH, c = rem2 - kg * NCH;
    const int col0 = ph + 4 * NC * kg;  // columns col0, col0 + 4, .. of byte phase ph
    const RsLv v = sLv[j];
    const int dy0 = c * v.csz;
    if (col0 < v.ncx && dy0 < v.ncy) {
      it_j = j; it_col0 = col0; it_dy0 = dy0; it_nc = NC;
      const uint8_t* blk = smem + j * G::BLK;
      const int sx = v.sxb + col0;
      rs_item<B, NC, NDY, G::PA, kRsPT>(blk + dy0 * kRsPT + (sx & ~3), (uint32_t)(sx & 3) * 8u, smem + G::OFF_ANC + v.aoff, acc);
      if (rs_commit<NC, NDY, TOP>(acc, v, col0, dy0, NC, &sBest[j])) sViol[j] = 1u;
    }
  } else if (LAST) {
    const int l = tid - MAIN;
    const int j = l / NCH, c = l - j * NCH;
    if (j < NB) {
      const RsLv v = sLv[j];
      const int dy0 = c * v.csz;
      if (v.ncx > 4 * NC * KG && dy0 < v.ncy) {  // the window has a column 4*NC*KG (= 16 or 32)
        it_j = j; it_col0 = 4 * NC * KG; it_dy0 = dy0; it_nc = 1;
        const uint8_t* blk = smem + j * G::BLK;
        const int sx = v.sxb + 4 * NC * KG;
        uint32_t a1[1][NDY];
#pragma unroll
        for (int d = 0; d < NDY; ++d) a1[0][d] = 0;
        rs_item<B, 1, NDY, G::PA, kRsPT>(blk + dy0 * kRsPT + (sx & ~3), (uint32_t)(sx & 3) * 8u, smem + G::OFF_ANC + v.aoff, a1);
#pragma unroll
        for (int d = 0; d < NDY; ++d) acc[0][d] = a1[0][d];
        if (rs_commit<NC, NDY, TOP>(acc, v, 4 * NC * KG, dy0, 1, &sBest[j])) sViol[j] = 1u;
      }
    }
  }
  __syncthreads();
  if (TOP) {
    // "every candidate updated the minimum" <=> the SADs never increase along the scan order of the
    // clamped window (libs/motion.cpp:312-337).  Nearly every block has been settled by the in-item test
    // above; the rest (flat or saturated content) exchange their SADs through shared memory and every
    // lane compares its candidates with their scan-order predecessors.
    const bool need = it_j >= 0 && sViol[it_j] == 0u;
    if (__syncthreads_or(need)) {  // CTA-uniform
      uint16_t* sd = reinterpret_cast<uint16_t*>(smem + G::OFF_SADS) + (it_j >= 0 ? it_j : 0) * (G::SADS / 2);
      const RsLv v = sLv[it_j >= 0 ? it_j : 0];
      const int ndy = min(v.csz, v.ncy - it_dy0);
      if (need) {
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i < it_nc && it_col0 + 4 * i < v.ncx)
#pragma unroll
            for (int d = 0; d < NDY; ++d)
              if (d < ndy) sd[(it_dy0 + d) * v.ncx + it_col0 + 4 * i] = (uint16_t)acc[i][d];
      }
      __syncthreads();
      if (need) {
        bool viol = false;
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i < it_nc && it_col0 + 4 * i < v.ncx)
#pragma unroll
            for (int d = 0; d < NDY; ++d) {
              const int idx = (it_dy0 + d) * v.ncx + it_col0 + 4 * i;
              if (d < ndy && idx > 0 && acc[i][d] > sd[idx - 1]) viol = true;
            }
        if (viol) sViol[it_j] = 1u;
      }
      __syncthreads();
    }
  }
  if (own) {  // tid < NB
    const uint32_t best = sBest[tid];
    const float m = (float)(best >> 16) * (1.0f / (float)(B * B));
    if (TOP) {
      const int idx = (int)(0xffffu - (best & 0xffffu));
      const bool any_viol = sViol[tid] != 0u;
      p.mv[gb] = any_viol ? make_float2((float)(x0 + idx % ncx - ax), (float)(y0 + idx / ncx - ay))
                          : make_float2(0.f, 0.f);
      p.mad[gb] = m;
    } else {
      // strict "<" against the MAD carried from the coarser level (libs/motion.cpp:401-405)
      if (best != 0xffffffffu && m < cur) {
        const int idx = (int)(best & 0xffffu);
        cur = m;
        mx = x0 + idx % ncx - ax;
        my = y0 + idx / ncx - ay;
      }
      p.mv[gb] = make_float2((float)mx, (float)my);
      p.mad[gb] = cur;
    }
  }
}

template <int B, bool TOP, int NC, int NDY, int NCH, int NB, bool LAST, int MINB, int KG = 1, int RMAX = 8>
cudaError_t launch_rs(const HbmaParams& p, uint32_t lvl, cudaStream_t st) {
  using G = RsGeom<B, NDY, NB, RMAX>;
  constexpr int kRsPT = G::PT;
  constexpr int SMEM = G::OFF_SADS + (TOP ? NB * G::SADS : 0);
  constexpr int THREADS = NB * 4 * KG * NCH + (LAST ? (NB * NCH + 31) / 32 * 32 : 0);
  EbmaMaps maps;
  const uint32_t n_slots = p.n_frames + 1;
  const uint8_t* base = p.pyr + p.lay.off[lvl];
  if (!encode_box(&maps.t, base, p.lay.w[lvl], p.lay.h[lvl], p.lay.pitch[lvl], p.lay.slot_bytes, n_slots, kRsPT,
                  B + 2 * p.r) ||
      !encode_box(&maps.a, base, p.lay.w[lvl], p.lay.h[lvl], p.lay.pitch[lvl], p.lay.slot_bytes, n_slots, G::PA, B))
    return cudaErrorNotSupported;
  static_assert(G::PA <= 256, "TMA box limit");
  auto kern = hbma_rs_kernel<B, TOP, NC, NDY, NCH, NB, LAST, MINB, KG, RMAX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  if (e != cudaSuccess) return e;
  const uint64_t n_ctas = (uint64_t)((p.mvw + NB - 1) / NB) * p.mvh * p.n_frames;
  kern<<<(uint32_t)n_ctas, THREADS, SMEM, st>>>(maps, p, (int)lvl);
  return cudaGetLastError();
}

// range classes: r = 5..8 -> windows up to 17 x 17: 4 columns x 6 rows per item, 3 chunks, 8 blocks per
// CTA and the 17th column on a fourth warp; r = 3, 4 -> up to 9 x 9: 3 columns x 5 rows, 2 chunks, 16
// blocks.  (More blocks per CTA for the small levels measured SLOWER -- 16 / 32 blocks at 8x8 / 4x4:
// 263 / 186 us against 206 / 114 us for 45 pairs of 1080p: those levels are bound by the two TMA
// requests every block issues, not by a CTA's fixed latency.)
template <int B, bool TOP>
cudaError_t launch_rs_class(const HbmaParams& p, uint32_t lvl, cudaStream_t st) {
  constexpr int MINB = B == 16 ? 5 : 6;
  // r = 9..16 -> windows up to 33 x 33: two groups of 4 columns per byte phase, 6 chunks, 2 blocks per CTA
  if (p.r >= 9) return launch_rs<B, TOP, 4, 6, 6, 2, true, MINB, 2, 16>(p, lvl, st);
  if (p.r >= 5) return launch_rs<B, TOP, 4, 6, 3, 8, true, MINB>(p, lvl, st);
  return launch_rs<B, TOP, 3, 5, 2, 16, false, MINB>(p, lvl, st);
}

}  // namespace

bool rs_level_supported(const HbmaParams& p) {
  return p.bw == 16 && p.bh == 16 && p.r >= 3 && p.r <= 16 && p.mv && p.mad && p.lay.levels <= 5 &&
         (uint64_t)p.mvw * p.mvh * p.n_frames <= 0x7fffffffull;
}

cudaError_t launch_rs_level(const HbmaParams& p, uint32_t lvl, bool top, cudaStream_t st) {
  const int B = 16 >> lvl;
#define SVC_RS_CASE(BB)                                                          \
  case BB:                                                                       \
    return top ? launch_rs_class<BB, true>(p, lvl, st) : launch_rs_class<BB, false>(p, lvl, st);
  switch (B) {
    SVC_RS_CASE(16)
    SVC_RS_CASE(8)
    SVC_RS_CASE(4)
    SVC_RS_CASE(2)
    SVC_RS_CASE(1)
  }
#undef SVC_RS_CASE
  return cudaErrorInvalidValue;
}

}  // namespace svc
