// k_hbma_pool.cu -- K2, large search ranges (16x16 blocks, top-level range r >= 5): the
// SAD-bound regime of the range / level sweep (BASELINE config 3).  Four kernels:
//   hbma_pool_kernel         any pyramid depth, r = 5..64: per-block windows, pooled work items
//   hbma_refine_kernel       one refinement level of a pyramid (one launch per level, r <= 32)
//   hbma_ebma_tile_kernel    L = 1, r <= 32: one shared window per tile of adjacent blocks
//   hbma_ebma_stripe_kernel  L = 1, r = 33..112: one block per CTA, window in column stripes
// (L = 1 is EstimateMotionExhaustiveSearch, libs/motion.cpp:268-340; the window position is then
// not data dependent; the tile kernel also searches the top level of deeper pyramids).  All share the work-item code (pool_level / sad_column_pre).
//
// Same arithmetic and scan-order rules as k_hbma.cu (reference libs/motion.cpp:268-465,
// 691-749); this file only changes how the work is laid out on the SM, to keep the
// integer-ALU pipe (the binding unit: VABSDIFF4 issues on it at 64 lanes/clk/SM) doing SADs
// and nothing else:
//   * per level, each motion block's clamped search window ((B+2r)^2 bytes, position data
//     dependent on the coarser level's vector) and its anchor block arrive by two TMA tensor
//     loads; the CTA then writes three more copies of the window, shifted left by 1, 2 and
//     3 bytes.  A candidate column at byte offset sx reads copy (sx & 3) at word sx >> 2:
//     every tracked word is a plain aligned LDS with an immediate offset -- no funnel
//     shifts and no address arithmetic in the SAD loop (the previous kernel spent 1 VIADD +
//     4 SHF per 32 VABSDIFF4 on them, all on the same pipe).  The copies sit 32 bytes
//     (mod 128) apart so that the 4 phases x 8 words a warp reads hit 32 different banks;
//   * a CTA walks NB motion blocks down the pyramid in lock step and pools their work items
//     (candidate column x chunk of candidate rows) over all its lanes, so windows with an
//     awkward number of columns (17, 33, 65, 129) still fill whole warps;
//   * the rows-per-item count NDY is chosen per range class so that 2r+1 candidate rows
//     split into full chunks (17 = 1x17, 33 = 3x11, 65 = 5x13, 129 = 10x13 - 1);
//   * the top level's "every candidate updated the minimum => zero vector" rule
//     (libs/motion.cpp:333-337) is checked with warp shuffles between neighbouring columns;
//     only items on a warp or window-row boundary go through (small) shared-memory arrays.
#include <float.h>

#include <algorithm>

#include "common.cuh"
#include "hbma_dev.cuh"

namespace svc {

struct PoolMaps {
  CUtensorMap t[5];  // tracked window box per level: PT x (B + 2r)
  CUtensorMap a[5];  // anchor block box per level: 16 x B
};

template <int RC, int NB, int NDY>
struct PoolGeom {
  static constexpr int PT = (16 + 2 * RC + 15 + 15) & ~15;  // window pitch: block + range + 16-byte origin slack
  static constexpr int ROWS = 16 + 2 * RC + NDY;            // + slack rows streamed by a short last chunk
  static constexpr int CS = ((PT * ROWS + 16 + 127) & ~127) + 32;  // copy stride, = 32 (mod 128)
  static constexpr int BLK = 4 * CS + 256;                  // 4 copies + anchor block (pitch 16)
  static constexpr int NCH = (2 * RC + 1 + NDY - 1) / NDY;  // chunks of candidate rows
  static constexpr int MAXWR = (NB * (2 * RC + 1) * NCH + 31) / 32;  // warp rounds per level
  static constexpr int EDGE = 2 * (MAXWR + NB * NCH) * NDY * 2;      // bytes of u16 edge arrays
  static constexpr int SMEM = NB * BLK + EDGE;
  static constexpr int PA = 16;                                      // anchor pitch
  static constexpr int kNB = NB, kNDY = NDY;
  // copy `ph` of block j's window / its anchor block
  __device__ static const uint8_t* window(const uint8_t* smem, int j, int ph) { return smem + j * BLK + ph * CS; }
  __device__ static const uint8_t* anchor(const uint8_t* smem, int j) { return smem + j * BLK + 4 * CS; }
};

// Top level only (L = 1, i.e. plain EBMA): the window position does not depend on data, so the NBX
// horizontally adjacent blocks of a CTA share ONE window (one TMA load, one copy build): it is
// 16 NBX + 2r wide instead of NBX (16 + 2r).
template <int RC, int NBX, int NDY, int B = 16>
struct TileGeomE {
  static_assert((B * NBX) % 16 == 0, "the anchor tile is a TMA box: its rows must be multiples of 16 bytes");
  static constexpr int PT = (B * NBX + 2 * RC + 15 + 15) & ~15;
  static constexpr int ROWS = B + 2 * RC + NDY;
  static constexpr int CS = ((PT * ROWS + 16 + 127) & ~127) + 32;
  static constexpr int PA = B * NBX;
  static constexpr int OFF_A = 4 * CS;                       // anchor tile: B rows x PA
  static constexpr int NCH = (2 * RC + 1 + NDY - 1) / NDY;
  static constexpr int MAXWR = (NBX * (2 * RC + 1) * NCH + 31) / 32;
  static constexpr int EDGE = 2 * (MAXWR + NBX * NCH) * NDY * 2;
  static constexpr int OFF_EDGE = (OFF_A + B * PA + 127) & ~127;
  static constexpr int SMEM = OFF_EDGE + EDGE;
  static constexpr int kNB = NBX, kNDY = NDY;
  __device__ static const uint8_t* window(const uint8_t* smem, int, int ph) { return smem + ph * CS; }
  __device__ static const uint8_t* anchor(const uint8_t* smem, int j) { return smem + OFF_A + j * B; }
};

struct PoolLv {  // one motion block at the current level
  int x0, y0, ncx, ncy;  // clamped candidate window (origin, size)
  int nch, csz;          // chunks of candidate rows, rows per chunk (balanced)
  int sxb, aoff;         // byte offset of x0 inside the TMA box; of the anchor block in its box
  uint32_t magic;        // ceil(2^32 / ncx) (0 when ncx == 1)
  int n_items;           // ncx * nch (0: no block)
  // scan order of the reference (libs/motion.cpp:312-330): index = dy * scan_ncx + scan_dx0 + dx.
  // Equal to (ncx, 0) unless the window is searched in column stripes.
  int scan_ncx, scan_dx0;
};

// NDY vertically adjacent candidates of one candidate column: streams B+NDY-1 aligned rows
// of the pre-shifted copy once, anchor block in registers.
template <int B, int NDY, int PT, int PA>
__device__ __forceinline__ void sad_column_pre(const uint8_t* __restrict__ tcol,
                                               const uint8_t* __restrict__ ablk,
                                               uint32_t (&acc)[NDY]) {
  constexpr int NW = B >= 4 ? B / 4 : 1;
  constexpr uint32_t MASK = B >= 4 ? 0xffffffffu : (B == 2 ? 0xffffu : 0xffu);
  uint32_t a[B][NW];
#pragma unroll
  for (int k = 0; k < B; ++k) {
    const uint8_t* q = ablk + k * PA;
    if constexpr (B == 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(q);
      a[k][0] = v.x; a[k][1] = v.y; a[k][2] = v.z; a[k][3] = v.w;
    } else if constexpr (B == 8) {
      const uint2 v = *reinterpret_cast<const uint2*>(q);
      a[k][0] = v.x; a[k][1] = v.y;
    } else if constexpr (B == 4) {
      a[k][0] = *reinterpret_cast<const uint32_t*>(q);
    } else if constexpr (B == 2) {
      a[k][0] = *reinterpret_cast<const uint16_t*>(q);
    } else {
      a[k][0] = *q;
    }
  }
#pragma unroll
  for (int t = 0; t < B + NDY - 1; ++t) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(tcol + t * PT);
    uint32_t tw[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) tw[k] = B >= 4 ? q[k] : (q[k] & MASK);
#pragma unroll
    for (int dyi = 0; dyi < NDY; ++dyi) {
      const int ar = t - dyi;
      if (ar >= 0 && ar < B) {
#pragma unroll
        for (int k = 0; k < NW; ++k) acc[dyi] = sad4_acc(tw[k], a[ar][k], acc[dyi]);
      }
    }
  }
}

// item -> (block j, chunk c, column dx)
template <int NB>
__device__ __forceinline__ void pool_decode(const int item, const int (&beg)[NB + 1], const PoolLv* sLv,
                                            int& j, int& c, int& dx, int& dy0, int& ndy) {
  j = 0;
  int bj = 0;
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    if (item >= beg[k]) { j = k; bj = beg[k]; }  // beg[] is non-decreasing (no dynamic indexing: registers)
  }
  const PoolLv& v = sLv[j];
  const int local = item - bj;
  c = v.magic ? (int)__umulhi((uint32_t)local, v.magic) : local;
  dx = local - c * v.ncx;
  dy0 = c * v.csz;
  ndy = min(v.csz, v.ncy - dy0);
}

template <int B, class G, int THREADS>
__device__ __forceinline__ void pool_level(const uint8_t* smem, const PoolLv* sLv, const int (&beg)[G::kNB + 1],
                                           const bool top, uint32_t* sBest, uint32_t* sViol,
                                           uint16_t* sTailW, uint16_t* sHeadW, uint16_t* sTailC,
                                           uint16_t* sHeadC) {
  constexpr int NB = G::kNB, NDY = G::kNDY;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total = beg[NB];
  for (int base = warp * 32; base < total; base += THREADS) {  // warp-uniform trip count
    const int item = base + lane;
    const bool act = item < total;
    int j, c, dx, dy0, ndy;
    pool_decode<NB>(act ? item : total - 1, beg, sLv, j, c, dx, dy0, ndy);
    const int ncx = sLv[j].ncx;
    uint32_t acc[NDY];
#pragma unroll
    for (int i = 0; i < NDY; ++i) acc[i] = 0;
    if (act) {
      const int sx = sLv[j].sxb + dx;
      sad_column_pre<B, NDY, G::PT, G::PA>(G::window(smem, j, sx & 3) + dy0 * G::PT + (sx & ~3),
                                           G::anchor(smem, j) + sLv[j].aoff, acc);
    }
    // packed (sad << 16 | scan index) minimum.  top: "<=" -> the last minimum wins (index stored
    // complemented); refinement: "<" -> the first minimum wins.  The pack is an integer
    // multiply-add (FMA pipe), only the minimum runs on the ALU pipe.
    uint32_t key = 0xffffffffu;
    {
      const int sncx = sLv[j].scan_ncx;
      const uint32_t idx0 = (uint32_t)(dy0 * sncx + sLv[j].scan_dx0 + dx);  // scan order inside the clamped window
      const uint32_t k0 = top ? 0xffffu - idx0 : idx0;
      const uint32_t kstep = top ? (uint32_t)(-sncx) : (uint32_t)sncx;
      uint32_t k[NDY];
#pragma unroll
      for (int i = 0; i < NDY; ++i) k[i] = acc[i] * 65536u + (k0 + (uint32_t)i * kstep);
      if (__all_sync(0xffffffffu, !act || ndy == NDY)) {  // whole chunks: no per-row predicates
#pragma unroll
        for (int i = 0; i < NDY; ++i) key = min(key, k[i]);
      } else {
#pragma unroll
        for (int i = 0; i < NDY; ++i) key = min(key, i < ndy ? k[i] : 0xffffffffu);
      }
      if (!act) key = 0xffffffffu;
    }
    // top level: has the SAD sequence of block j increased anywhere in scan order?  Skipped once
    // a violation is known (the common case on textured content after the first few items).
    // (volatile: the flag is raised by other warps while this one loops)
    if (top && __any_sync(0xffffffffu, act && *reinterpret_cast<volatile uint32_t*>(&sViol[j]) == 0u)) {
      // the scan-order predecessor of candidate (dx, dy) is (dx-1, dy): the previous lane
      bool viol = false;
#pragma unroll
      for (int i = 0; i < NDY; ++i) {
        const uint32_t pv = __shfl_up_sync(0xffffffffu, acc[i], 1);
        viol |= (i < ndy && acc[i] > pv);
      }
      viol &= (lane > 0 && dx > 0);
      if (act) {
        const int wr = base >> 5;
        if (lane == 0 && dx > 0) {
#pragma unroll
          for (int i = 0; i < NDY; ++i) sHeadW[wr * NDY + i] = (uint16_t)acc[i];
        }
        if (lane == 31) {
#pragma unroll
          for (int i = 0; i < NDY; ++i) sTailW[wr * NDY + i] = (uint16_t)acc[i];
        }
        if (dx == 0) {
#pragma unroll
          for (int i = 0; i < NDY; ++i) sHeadC[(j * G::NCH + c) * NDY + i] = (uint16_t)acc[i];
        }
        if (dx == ncx - 1) {
#pragma unroll
          for (int i = 0; i < NDY; ++i) sTailC[(j * G::NCH + c) * NDY + i] = (uint16_t)acc[i];
        }
        if (viol) sViol[j] = 1u;
      }
    }
    if constexpr (NB == 1) {
      key = __reduce_min_sync(0xffffffffu, key);
      if (lane == 0) atomicMin(&sBest[0], key);
    } else {
      const int j0 = __shfl_sync(0xffffffffu, j, 0);
      if (__all_sync(0xffffffffu, !act || j == j0)) {
        key = __reduce_min_sync(0xffffffffu, key);
        if (lane == 0) atomicMin(&sBest[j0], key);
      } else if (act) {
        atomicMin(&sBest[j], key);
      }
    }
  }
}

template <int RC, int NB, int NDY, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
hbma_pool_kernel(const __grid_constant__ PoolMaps maps, const __grid_constant__ HbmaParams p) {
  using G = PoolGeom<RC, NB, NDY>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ PoolLv sLv[NB];
  __shared__ uint32_t sBest[NB], sViol[NB];
  uint16_t* sTailW = reinterpret_cast<uint16_t*>(smem + NB * G::BLK);
  uint16_t* sHeadW = sTailW + G::MAXWR * NDY;
  uint16_t* sTailC = sHeadW + G::MAXWR * NDY;
  uint16_t* sHeadC = sTailC + NB * G::NCH * NDY;

  const int tid = threadIdx.x;
  const uint32_t per_frame = p.mvw * p.mvh;
  const uint64_t n_blocks = (uint64_t)per_frame * p.n_frames;
  const int r = (int)p.r, L = (int)p.lay.levels;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);

  // thread j < NB owns the state of motion block j of this CTA
  int mx = 0, my = 0, bx = 0, by = 0;
  float cur = FLT_MAX;
  uint32_t f = 0, bi = 0;
  bool own = false;
  if (tid < NB) {
    const uint32_t gb = blockIdx.x * NB + tid;  // < 2^31 (checked on the host)
    own = gb < (uint32_t)n_blocks;
    if (own) {
      f = gb / per_frame;
      bi = gb - f * per_frame;
      bx = (int)(bi % p.mvw);
      by = (int)(bi / p.mvw);
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(NB));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t parity = 0;
  for (int l = L - 1; l >= 0; --l) {
    const bool top = (l == L - 1);
    const int B = 16 >> l;
    const int box_h = B + 2 * r;
    int x0 = 0, y0 = 0, ncx = 1, ax = 0, ay = 0;
    if (tid < NB) {
      PoolLv v{};
      if (own) {
        const int fw = (int)p.lay.w[l], fh = (int)p.lay.h[l];
        if (!top) { mx *= 2; my *= 2; }
        ax = bx * B;
        ay = by * B;
        const int cx = ax + mx, cy = ay + my;
        x0 = max(0, cx - r);
        y0 = max(0, cy - r);
        const int x1 = min(fw - B + 1, cx + r + 1), y1 = min(fh - B + 1, cy + r + 1);
        ncx = x1 - x0;
        const int ncy = y1 - y0;
        const int nch = (ncy + NDY - 1) / NDY;
        v.x0 = x0; v.y0 = y0; v.ncx = ncx; v.ncy = ncy;
        v.nch = nch;
        v.csz = (ncy + nch - 1) / nch;
        v.sxb = x0 & 15;
        v.aoff = ax & 15;
        v.magic = ncx > 1 ? 0xffffffffu / (uint32_t)ncx + 1u : 0u;
        v.n_items = ncx * nch;
        v.scan_ncx = ncx;
        v.scan_dx0 = 0;
        if (p.counters) {
          atomicAdd(p.counters, (unsigned long long)(ncx * ncy));
          atomicAdd(p.counters + 1, (unsigned long long)(ncx * ncy) * B * B);
        }
      }
      sLv[tid] = v;
      sBest[tid] = 0xffffffffu;
      sViol[tid] = 0u;
      if (own) {
        // order the generic-proxy accesses of the previous level before the async-proxy overwrite
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                     "r"((uint32_t)(G::PT * box_h + 16 * B)) : "memory");
        uint8_t* blk = smem + tid * G::BLK;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"((uint32_t)__cvta_generic_to_shared(blk)), "l"(&maps.t[l]), "r"(x0 & ~15), "r"(y0), "r"((int)f),
            "r"(bar_addr) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"((uint32_t)__cvta_generic_to_shared(blk + 4 * G::CS)), "l"(&maps.a[l]), "r"(ax & ~15), "r"(ay),
            "r"((int)f + 1), "r"(bar_addr) : "memory");
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
      }
    }
    {  // windows landed; sLv / sBest / sViol published (arrive = release, wait = acquire)
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
      }
      parity ^= 1u;
    }
    // copies shifted left by 1, 2, 3 bytes (copy s, word i = bytes 4i+s .. 4i+s+3 of the window);
    // all NB windows as one flat list of 16-byte vectors, 4 independent vectors per thread and trip
    {
      const int nvec = (G::PT / 16) * box_h, total = NB * nvec;
      const uint32_t vmagic = 0xffffffffu / (uint32_t)nvec + 1u;
      const int lane = tid & 31;
      for (int vb = tid - lane; vb < total; vb += 4 * THREADS) {  // warp-uniform trip count (shuffles inside)
        const int v0 = vb + lane;
        uint4 v[4];
        uint32_t nx[4];
        uint8_t* dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int vi = min(v0 + u * THREADS, total - 1);
          const int j = NB > 1 ? (int)__umulhi((uint32_t)vi, vmagic) : 0;
          uint8_t* q = smem + j * G::BLK + (vi - j * nvec) * 16;
          v[u] = *reinterpret_cast<const uint4*>(q);
          dst[u] = q;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          // first word of the next vector: the next lane holds it (lane 31: from shared memory; a
          // window's last vector takes a word of the slack rows, whose value is never used)
          nx[u] = __shfl_down_sync(0xffffffffu, v[u].x, 1);
          if (lane == 31) nx[u] = *reinterpret_cast<const uint32_t*>(dst[u] + 16);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (v0 + u * THREADS < total) {
#pragma unroll
            for (int sft = 1; sft < 4; ++sft) {
              uint4 o;
              o.x = __funnelshift_r(v[u].x, v[u].y, 8 * sft);
              o.y = __funnelshift_r(v[u].y, v[u].z, 8 * sft);
              o.z = __funnelshift_r(v[u].z, v[u].w, 8 * sft);
              o.w = __funnelshift_r(v[u].w, nx[u], 8 * sft);
              *reinterpret_cast<uint4*>(dst[u] + sft * G::CS) = o;
            }
          }
        }
      }
    }
    __syncthreads();
    int beg[NB + 1];
    beg[0] = 0;
#pragma unroll
    for (int j = 0; j < NB; ++j) beg[j + 1] = beg[j] + sLv[j].n_items;
    switch (B) {
      case 16: pool_level<16, G, THREADS>(smem, sLv, beg, top, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC); break;
      case 8:  pool_level<8, G, THREADS>(smem, sLv, beg, top, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC); break;
      case 4:  pool_level<4, G, THREADS>(smem, sLv, beg, top, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC); break;
      case 2:  pool_level<2, G, THREADS>(smem, sLv, beg, top, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC); break;
      default: pool_level<1, G, THREADS>(smem, sLv, beg, top, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC); break;
    }
    __syncthreads();
    if (top) {
      // scan-order neighbours that were not in adjacent lanes
      const int total = beg[NB];
      const int nwr = (total + 31) >> 5;
      for (int e = tid; e < (nwr - 1) * NDY; e += THREADS) {  // first lane of a warp round vs. last lane of the previous one
        const int wr = 1 + e / NDY, i = e - (wr - 1) * NDY;
        int j, c, dx, dy0, ndy;
        pool_decode<NB>(wr * 32, beg, sLv, j, c, dx, dy0, ndy);
        if (dx > 0 && i < ndy && sHeadW[wr * NDY + i] > sTailW[(wr - 1) * NDY + i]) sViol[j] = 1u;
      }
      for (int e = tid; e < NB * G::NCH * NDY; e += THREADS) {  // first column vs. last column of the previous row
        const int jc = e / NDY, i = e - jc * NDY;
        const int j = jc / G::NCH, c = jc - j * G::NCH;
        const PoolLv& v = sLv[j];
        if (v.n_items == 0 || c >= v.nch) continue;
        if (i >= min(v.csz, v.ncy - c * v.csz)) continue;
        uint32_t pv;
        if (i > 0) pv = sTailC[jc * NDY + i - 1];
        else if (c > 0) pv = sTailC[(jc - 1) * NDY + v.csz - 1];  // only the last chunk can be short
        else continue;
        if (sHeadC[jc * NDY + i] > pv) sViol[j] = 1u;
      }
      __syncthreads();
    }
    if (own) {  // tid < NB
      const uint32_t best = sBest[tid];
      const float m = (float)(best >> 16) * (1.0f / (float)(B * B));
      const int idx = top ? (int)(0xffffu - (best & 0xffffu)) : (int)(best & 0xffffu);
      const int nmx = x0 + idx % ncx - ax, nmy = y0 + idx / ncx - ay;
      if (top) {
        const bool any_viol = sViol[tid] != 0u;
        cur = m;
        mx = any_viol ? nmx : 0;
        my = any_viol ? nmy : 0;
      } else if (m < cur) {
        cur = m;
        mx = nmx;
        my = nmy;
      }
    }
  }
  if (own) {
    const uint64_t o = (uint64_t)f * per_frame + bi;
    if (p.mv) p.mv[o] = make_float2((float)mx, (float)my);
    if (p.mad) p.mad[o] = cur;
  }
}

// three copies of ONE window (at smem, `total` 16-byte vectors) shifted left by 1, 2, 3 bytes, copy s at
// smem + s * cs (see hbma_pool_kernel); 4 independent vectors per thread and trip
template <int THREADS>
__device__ __forceinline__ void tile_build_copies(uint8_t* smem, const int total, const int cs) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int vb = tid - lane; vb < total; vb += 4 * THREADS) {  // warp-uniform trip count (shuffles inside)
    const int v0 = vb + lane;
    uint4 v[4];
    uint32_t nx[4];
    uint8_t* q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      q[u] = smem + min(v0 + u * THREADS, total - 1) * 16;
      v[u] = *reinterpret_cast<const uint4*>(q[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      nx[u] = __shfl_down_sync(0xffffffffu, v[u].x, 1);
      if (lane == 31) nx[u] = *reinterpret_cast<const uint32_t*>(q[u] + 16);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (v0 + u * THREADS < total) {
#pragma unroll
        for (int sft = 1; sft < 4; ++sft) {
          uint4 o;
          o.x = __funnelshift_r(v[u].x, v[u].y, 8 * sft);
          o.y = __funnelshift_r(v[u].y, v[u].z, 8 * sft);
          o.z = __funnelshift_r(v[u].z, v[u].w, 8 * sft);
          o.w = __funnelshift_r(v[u].w, nx[u], 8 * sft);
          *reinterpret_cast<uint4*>(q[u] + sft * cs) = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// L = 1 (EstimateMotionExhaustiveSearch with 16x16 blocks, and the L = 1 column of the sweep):
// one shared window per tile of NBX horizontally adjacent blocks, see TileGeomE.  Same work
// items, same argmin, same zero-vector rule as pool_level's top level.
// ---------------------------------------------------------------------------------------
// B = block size at the searched level `lvl` (16 >> lvl): 16 / level 0 for L = 1; the top level of a
// deeper pyramid otherwise, whose vector and MAD it leaves in p.mv / p.mad for hbma_refine_kernel.
template <int RC, int NBX, int NDY, int THREADS, int MINB, int B = 16>
__global__ void __launch_bounds__(THREADS, MINB)
hbma_ebma_tile_kernel(const __grid_constant__ EbmaMaps maps, const __grid_constant__ HbmaParams p,
                      const int lvl) {
  using G = TileGeomE<RC, NBX, NDY, B>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ PoolLv sLv[NBX];
  __shared__ uint32_t sBest[NBX], sViol[NBX];
  uint16_t* sTailW = reinterpret_cast<uint16_t*>(smem + G::OFF_EDGE);
  uint16_t* sHeadW = sTailW + G::MAXWR * NDY;
  uint16_t* sTailC = sHeadW + G::MAXWR * NDY;
  uint16_t* sHeadC = sTailC + NBX * G::NCH * NDY;

  const int tid = threadIdx.x;
  const int r = (int)p.r;
  const uint32_t tiles_per_row = (p.mvw + NBX - 1) / NBX, tiles_per_frame = tiles_per_row * p.mvh;
  const uint32_t f = blockIdx.x / tiles_per_frame, ti = blockIdx.x - f * tiles_per_frame;
  const int by = (int)(ti / tiles_per_row), bx0 = (int)(ti - (uint32_t)by * tiles_per_row) * NBX;
  const int fw = (int)p.lay.w[lvl], fh = (int)p.lay.h[lvl];
  const int ay = by * B;
  const int y0 = max(0, ay - r), y1 = min(fh - B + 1, ay + r + 1);
  const int wx = max(0, bx0 * B - r) & ~15;  // 16-byte aligned origin of the shared window
  const int box_h = B + 2 * r;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);

  // thread j < NBX owns block bx0 + j
  int x0 = 0, ncx = 1, ax = 0;
  const bool own = tid < NBX && (uint32_t)(bx0 + tid) < p.mvw;
  if (tid < NBX) {
    PoolLv v{};
    if (own) {
      ax = (bx0 + tid) * B;
      x0 = max(0, ax - r);
      const int x1 = min(fw - B + 1, ax + r + 1);
      ncx = x1 - x0;
      const int ncy = y1 - y0;
      const int nch = (ncy + NDY - 1) / NDY;
      v.x0 = x0; v.y0 = y0; v.ncx = ncx; v.ncy = ncy;
      v.nch = nch;
      v.csz = (ncy + nch - 1) / nch;
      v.sxb = x0 - wx;
      v.aoff = 0;
      v.magic = ncx > 1 ? 0xffffffffu / (uint32_t)ncx + 1u : 0u;
      v.n_items = ncx * nch;
      v.scan_ncx = ncx;
      v.scan_dx0 = 0;
      if (p.counters) {
        atomicAdd(p.counters, (unsigned long long)(ncx * ncy));
        atomicAdd(p.counters + 1, (unsigned long long)(ncx * ncy) * (unsigned long long)(B * B));
      }
    }
    sLv[tid] = v;
    sBest[tid] = 0xffffffffu;
    sViol[tid] = 0u;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                 "r"((uint32_t)(G::PT * box_h + B * G::PA)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(&maps.t), "r"(wx), "r"(y0), "r"((int)f),
        "r"(bar_addr) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"((uint32_t)__cvta_generic_to_shared(smem + G::OFF_A)), "l"(&maps.a), "r"(bx0 * B), "r"(ay),
        "r"((int)f + 1), "r"(bar_addr) : "memory");
  }
  __syncthreads();  // barrier initialised before anyone polls it; sLv / sBest / sViol published
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
  tile_build_copies<THREADS>(smem, (G::PT / 16) * box_h, G::CS);
  __syncthreads();
  int beg[NBX + 1];
  beg[0] = 0;
#pragma unroll
  for (int j = 0; j < NBX; ++j) beg[j + 1] = beg[j] + sLv[j].n_items;
  pool_level<B, G, THREADS>(smem, sLv, beg, true, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC);
  __syncthreads();
  {  // scan-order neighbours that were not in adjacent lanes (as in hbma_pool_kernel)
    const int total = beg[NBX];
    const int nwr = (total + 31) >> 5;
    for (int e = tid; e < (nwr - 1) * NDY; e += THREADS) {
      const int wr = 1 + e / NDY, i = e - (wr - 1) * NDY;
      int j, c, dx, dy0, ndy;
      pool_decode<NBX>(wr * 32, beg, sLv, j, c, dx, dy0, ndy);
      if (dx > 0 && i < ndy && sHeadW[wr * NDY + i] > sTailW[(wr - 1) * NDY + i]) sViol[j] = 1u;
    }
    for (int e = tid; e < NBX * G::NCH * NDY; e += THREADS) {
      const int jc = e / NDY, i = e - jc * NDY;
      const int j = jc / G::NCH, c = jc - j * G::NCH;
      const PoolLv& v = sLv[j];
      if (v.n_items == 0 || c >= v.nch) continue;
      if (i >= min(v.csz, v.ncy - c * v.csz)) continue;
      uint32_t pv;
      if (i > 0) pv = sTailC[jc * NDY + i - 1];
      else if (c > 0) pv = sTailC[(jc - 1) * NDY + v.csz - 1];
      else continue;
      if (sHeadC[jc * NDY + i] > pv) sViol[j] = 1u;
    }
  }
  __syncthreads();
  if (own) {
    const uint32_t best = sBest[tid];
    const int idx = (int)(0xffffu - (best & 0xffffu));
    const bool any_viol = sViol[tid] != 0u;
    const uint64_t o = ((uint64_t)f * p.mvh + (uint32_t)by) * p.mvw + (uint32_t)(bx0 + tid);
    if (p.mv) p.mv[o] = any_viol ? make_float2((float)(x0 + idx % ncx - ax), (float)(y0 + idx / ncx - ay))
                                 : make_float2(0.f, 0.f);
    if (p.mad) p.mad[o] = (float)(best >> 16) * (1.0f / (float)(B * B));
  }
}

// ---------------------------------------------------------------------------------------
// L = 1, wide ranges (r = 33 .. 112): one block per CTA, its window searched in STRIPES of CW
// candidate columns.  A stripe's window (CW + 15 bytes wide) and its three shifted copies take
// 40 KB instead of 100 KB for r = 64, so 4-5 CTAs share an SM and cover each other's load, build
// and hand-over phases; best key and violation flag are carried across the stripes in shared
// memory, as are the SADs of each stripe's last candidate column (the scan-order predecessors of
// the next stripe's first column) and of the window's first column (successors of the last one).
// ---------------------------------------------------------------------------------------
template <int RC, int CW, int NDY>
struct StripeGeomE {
  static constexpr int PT = (16 + CW - 1 + 15 + 15) & ~15;
  static constexpr int ROWS = 16 + 2 * RC + NDY;
  static constexpr int CS = ((PT * ROWS + 16 + 127) & ~127) + 32;
  static constexpr int PA = 16;
  static constexpr int OFF_A = 4 * CS;
  static constexpr int NCH = (2 * RC + 1 + NDY - 1) / NDY;
  static constexpr int MAXWR = (CW * NCH + 31) / 32;
  static constexpr int EDGE = 2 * (MAXWR + NCH) * NDY * 2;
  static constexpr int OFF_EDGE = (OFF_A + 256 + 127) & ~127;
  static constexpr int OFF_CARRY = (OFF_EDGE + EDGE + 15) & ~15;   // u16[2 RC + 1]: last column of the previous stripe
  static constexpr int OFF_HEAD0 = OFF_CARRY + ((2 * (2 * RC + 1) + 15) & ~15);  // u16[2 RC + 1]: first column of the window
  static constexpr int SMEM = OFF_HEAD0 + ((2 * (2 * RC + 1) + 15) & ~15);
  static constexpr int kNB = 1, kNDY = NDY;
  __device__ static const uint8_t* window(const uint8_t* smem, int, int ph) { return smem + ph * CS; }
  __device__ static const uint8_t* anchor(const uint8_t* smem, int) { return smem + OFF_A; }
};

template <int RC, int CW, int NDY, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
hbma_ebma_stripe_kernel(const __grid_constant__ EbmaMaps maps, const __grid_constant__ HbmaParams p) {
  using G = StripeGeomE<RC, CW, NDY>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ PoolLv sLv[1];
  __shared__ uint32_t sBest[1], sViol[1];
  uint16_t* sTailW = reinterpret_cast<uint16_t*>(smem + G::OFF_EDGE);
  uint16_t* sHeadW = sTailW + G::MAXWR * NDY;
  uint16_t* sTailC = sHeadW + G::MAXWR * NDY;
  uint16_t* sHeadC = sTailC + G::NCH * NDY;
  uint16_t* sCarry = reinterpret_cast<uint16_t*>(smem + G::OFF_CARRY);
  uint16_t* sHead0 = reinterpret_cast<uint16_t*>(smem + G::OFF_HEAD0);

  const int tid = threadIdx.x;
  const int r = (int)p.r;
  const uint32_t per_frame = p.mvw * p.mvh;
  const uint32_t f = blockIdx.x / per_frame, bi = blockIdx.x - f * per_frame;
  const int bx = (int)(bi % p.mvw), by = (int)(bi / p.mvw);
  const int fw = (int)p.lay.w[0], fh = (int)p.lay.h[0];
  const int ax = bx * 16, ay = by * 16;
  const int x0 = max(0, ax - r), x1 = min(fw - 16 + 1, ax + r + 1);
  const int y0 = max(0, ay - r), y1 = min(fh - 16 + 1, ay + r + 1);
  const int ncx = x1 - x0, ncy = y1 - y0;
  const int nch = (ncy + NDY - 1) / NDY, csz = (ncy + nch - 1) / nch;
  const int box_h = 16 + 2 * r;
  const int n_stripes = (ncx + CW - 1) / CW;
  const int cw = (ncx + n_stripes - 1) / n_stripes;  // balanced stripes, cw <= CW
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sBest[0] = 0xffffffffu;
    sViol[0] = 0u;
    if (p.counters) {
      atomicAdd(p.counters, (unsigned long long)(ncx * ncy));
      atomicAdd(p.counters + 1, (unsigned long long)(ncx * ncy) * 256ull);
    }
  }
  for (int k = 0; k < n_stripes; ++k) {
    const int xs = x0 + k * cw, ncs = min(cw, ncx - k * cw);
    if (tid == 0) {
      PoolLv v{};
      v.x0 = xs; v.y0 = y0; v.ncx = ncs; v.ncy = ncy;
      v.nch = nch; v.csz = csz;
      v.sxb = xs & 15;
      v.aoff = 0;
      v.magic = ncs > 1 ? 0xffffffffu / (uint32_t)ncs + 1u : 0u;
      v.n_items = ncs * nch;
      v.scan_ncx = ncx;
      v.scan_dx0 = k * cw;
      sLv[0] = v;
      // the previous stripe's window was read through the generic proxy (all threads are past the
      // barrier that closed it)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                   "r"((uint32_t)(G::PT * box_h + (k == 0 ? 256 : 0))) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(&maps.t), "r"(xs & ~15), "r"(y0), "r"((int)f),
          "r"(bar_addr) : "memory");
      if (k == 0)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"((uint32_t)__cvta_generic_to_shared(smem + G::OFF_A)), "l"(&maps.a), "r"(ax), "r"(ay),
            "r"((int)f + 1), "r"(bar_addr) : "memory");
    }
    __syncthreads();  // barrier initialised / re-armed before anyone polls it; sLv published
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_addr), "r"((uint32_t)(k & 1)) : "memory");
      }
    }
    tile_build_copies<THREADS>(smem, (G::PT / 16) * box_h, G::CS);
    __syncthreads();
    int beg[2] = {0, ncs * nch};
    pool_level<16, G, THREADS>(smem, sLv, beg, true, sBest, sViol, sTailW, sHeadW, sTailC, sHeadC);
    __syncthreads();
    {  // scan-order neighbours that were not in adjacent lanes
      const int nwr = (beg[1] + 31) >> 5;
      for (int e = tid; e < (nwr - 1) * NDY; e += THREADS) {  // warp-round boundaries inside the stripe
        const int wr = 1 + e / NDY, i = e - (wr - 1) * NDY;
        int j, c, dx, dy0, ndy;
        pool_decode<1>(wr * 32, beg, sLv, j, c, dx, dy0, ndy);
        if (dx > 0 && i < ndy && sHeadW[wr * NDY + i] > sTailW[(wr - 1) * NDY + i]) sViol[0] = 1u;
      }
      for (int dy = tid; dy < ncy; dy += THREADS) {  // first column of the stripe, row dy
        const int c = dy / csz, i = dy - c * csz;
        const uint32_t head = sHeadC[c * NDY + i];
        if (k > 0) {
          if (head > sCarry[dy]) sViol[0] = 1u;  // predecessor: last column of the previous stripe, same row
        } else {
          sHead0[dy] = (uint16_t)head;
        }
        if (k == n_stripes - 1 && dy > 0) {
          // predecessor of the window's first column: its last column, one row up
          const int cp = (dy - 1) / csz, ip = (dy - 1) - cp * csz;
          if (sHead0[dy] > sTailC[cp * NDY + ip]) sViol[0] = 1u;
        }
        sCarry[dy] = sTailC[c * NDY + i];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const uint32_t best = sBest[0];
    const int idx = (int)(0xffffu - (best & 0xffffu));
    const bool any_viol = sViol[0] != 0u;
    const uint64_t o = (uint64_t)f * per_frame + bi;
    if (p.mv) p.mv[o] = any_viol ? make_float2((float)(x0 + idx % ncx - ax), (float)(y0 + idx / ncx - ay))
                                 : make_float2(0.f, 0.f);
    if (p.mad) p.mad[o] = (float)(best >> 16) * (1.0f / 256.0f);
  }
}

// ---------------------------------------------------------------------------------------
// Level-synchronous path for pyramids (2 <= L <= 5, r <= 32): one launch per level.  The top
// level runs on hbma_ebma_tile_kernel (shared windows); every refinement level on the kernel
// below, which is hbma_pool_kernel reduced to ONE level with a compile-time block size: the
// anchor block takes B*B/4 registers instead of 64, no level loop, no top-level code, so more
// CTAs share an SM.  The vector and MAD of the coarser level travel through p.mv / p.mad (the
// output arrays), exactly the values the reference carries between its RefineHierMotionEst calls
// (libs/motion.cpp:451-464).
// ---------------------------------------------------------------------------------------
template <int RC, int NB, int NDY, int B>
struct RefineGeom {
  static constexpr int PT = (B + 2 * RC + 15 + 15) & ~15;
  static constexpr int ROWS = B + 2 * RC + NDY;
  static constexpr int CS = ((PT * ROWS + 16 + 127) & ~127) + 32;
  static constexpr int BLK = 4 * CS + ((16 * B + 127) & ~127);  // 4 copies + anchor block (pitch 16)
  static constexpr int SMEM = NB * BLK;
  static constexpr int PA = 16, NCH = 1;
  static constexpr int kNB = NB, kNDY = NDY;
  __device__ static const uint8_t* window(const uint8_t* smem, int j, int ph) { return smem + j * BLK + ph * CS; }
  __device__ static const uint8_t* anchor(const uint8_t* smem, int j) { return smem + j * BLK + 4 * CS; }
};

template <int RC, int NB, int NDY, int THREADS, int MINB, int B>
__global__ void __launch_bounds__(THREADS, MINB)
hbma_refine_kernel(const __grid_constant__ EbmaMaps maps, const __grid_constant__ HbmaParams p, const int lvl) {
  using G = RefineGeom<RC, NB, NDY, B>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ PoolLv sLv[NB];
  __shared__ uint32_t sBest[NB];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t per_frame = p.mvw * p.mvh;
  const uint32_t n_blocks = per_frame * p.n_frames;  // < 2^31 (checked on the host)
  const int r = (int)p.r;
  const int box_h = B + 2 * r;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(NB));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int mx = 0, my = 0, x0 = 0, y0 = 0, ncx = 1, ax = 0, ay = 0;
  float cur = 0.f;
  uint32_t gb = 0;
  bool own = false;
  if (tid < NB) {
    gb = blockIdx.x * NB + tid;
    own = gb < n_blocks;
    PoolLv v{};
    if (own) {
      const uint32_t f = gb / per_frame, bi = gb - f * per_frame;
      const int bx = (int)(bi % p.mvw), by = (int)(bi / p.mvw);
      const float2 m = p.mv[gb];  // integer valued (libs/motion.cpp:326-327, 403-404)
      cur = p.mad[gb];
      mx = 2 * (int)m.x;          // motion_field *= 2 between levels (libs/motion.cpp:459)
      my = 2 * (int)m.y;
      const int fw = (int)p.lay.w[lvl], fh = (int)p.lay.h[lvl];
      ax = bx * B;
      ay = by * B;
      const int cx = ax + mx, cy = ay + my;
      x0 = max(0, cx - r);
      y0 = max(0, cy - r);
      const int x1 = min(fw - B + 1, cx + r + 1), y1 = min(fh - B + 1, cy + r + 1);
      ncx = x1 - x0;
      const int ncy = y1 - y0;
      const int nch = (ncy + NDY - 1) / NDY;
      v.x0 = x0; v.y0 = y0; v.ncx = ncx; v.ncy = ncy;
      v.nch = nch;
      v.csz = (ncy + nch - 1) / nch;
      v.sxb = x0 & 15;
      v.aoff = ax & 15;
      v.magic = ncx > 1 ? 0xffffffffu / (uint32_t)ncx + 1u : 0u;
      v.n_items = ncx * nch;
      v.scan_ncx = ncx;
      v.scan_dx0 = 0;
      if (p.counters) {
        atomicAdd(p.counters, (unsigned long long)(ncx * ncy));
        atomicAdd(p.counters + 1, (unsigned long long)(ncx * ncy) * (unsigned long long)(B * B));
      }
      sLv[tid] = v;
      sBest[tid] = 0xffffffffu;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                   "r"((uint32_t)(G::PT * box_h + 16 * B)) : "memory");
      uint8_t* blk = smem + tid * G::BLK;
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"((uint32_t)__cvta_generic_to_shared(blk)), "l"(&maps.t), "r"(x0 & ~15), "r"(y0), "r"((int)f),
          "r"(bar_addr) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"((uint32_t)__cvta_generic_to_shared(blk + 4 * G::CS)), "l"(&maps.a), "r"(ax & ~15), "r"(ay),
          "r"((int)f + 1), "r"(bar_addr) : "memory");
    } else {
      sLv[tid] = v;
      sBest[tid] = 0xffffffffu;
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
  {  // shifted copies of the NB windows, as one flat list of 16-byte vectors (see hbma_pool_kernel)
    const int nvec = (G::PT / 16) * box_h, total = NB * nvec;
    const uint32_t vmagic = 0xffffffffu / (uint32_t)nvec + 1u;
    for (int vb = tid - lane; vb < total; vb += 4 * THREADS) {
      const int v0 = vb + lane;
      uint4 v[4];
      uint32_t nx[4];
      uint8_t* dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int vi = min(v0 + u * THREADS, total - 1);
        const int j = NB > 1 ? (int)__umulhi((uint32_t)vi, vmagic) : 0;
        uint8_t* q = smem + j * G::BLK + (vi - j * nvec) * 16;
        v[u] = *reinterpret_cast<const uint4*>(q);
        dst[u] = q;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        nx[u] = __shfl_down_sync(0xffffffffu, v[u].x, 1);
        if (lane == 31) nx[u] = *reinterpret_cast<const uint32_t*>(dst[u] + 16);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v0 + u * THREADS < total) {
#pragma unroll
          for (int sft = 1; sft < 4; ++sft) {
            uint4 o;
            o.x = __funnelshift_r(v[u].x, v[u].y, 8 * sft);
            o.y = __funnelshift_r(v[u].y, v[u].z, 8 * sft);
            o.z = __funnelshift_r(v[u].z, v[u].w, 8 * sft);
            o.w = __funnelshift_r(v[u].w, nx[u], 8 * sft);
            *reinterpret_cast<uint4*>(dst[u] + sft * G::CS) = o;
          }
        }
      }
    }
  }
  __syncthreads();
  int beg[NB + 1];
  beg[0] = 0;
#pragma unroll
  for (int j = 0; j < NB; ++j) beg[j + 1] = beg[j] + sLv[j].n_items;
  pool_level<B, G, THREADS>(smem, sLv, beg, false, sBest, nullptr, nullptr, nullptr, nullptr, nullptr);
  __syncthreads();
  if (own) {  // strict "<" against the MAD carried from the coarser level (libs/motion.cpp:401-405)
    const uint32_t best = sBest[tid];
    const float m = (float)(best >> 16) * (1.0f / (float)(B * B));
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

template <int RC, int NBX, int NDY, int THREADS, int MINB, int B>
static cudaError_t launch_top_tile(const HbmaParams& p, uint32_t lvl, cudaStream_t st) {
  using G = TileGeomE<RC, NBX, NDY, B>;
  static_assert(G::SMEM <= 227 * 1024 && G::PT <= 256 && G::PA <= 256, "tile geometry does not fit");
  EbmaMaps maps;
  const uint32_t n_slots = p.n_frames + 1;
  const uint8_t* base = p.pyr + p.lay.off[lvl];
  if (!encode_box(&maps.t, base, p.lay.w[lvl], p.lay.h[lvl], p.lay.pitch[lvl], p.lay.slot_bytes, n_slots, G::PT,
                  B + 2 * p.r) ||
      !encode_box(&maps.a, base, p.lay.w[lvl], p.lay.h[lvl], p.lay.pitch[lvl], p.lay.slot_bytes, n_slots, G::PA, B))
    return cudaErrorNotSupported;
  auto kern = hbma_ebma_tile_kernel<RC, NBX, NDY, THREADS, MINB, B>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return e;
  const uint64_t tiles = (uint64_t)((p.mvw + NBX - 1) / NBX) * p.mvh * p.n_frames;
  kern<<<(uint32_t)tiles, THREADS, G::SMEM, st>>>(maps, p, (int)lvl);
  return cudaGetLastError();
}

template <int RC, int NB, int NDY, int THREADS, int MINB, int B>
static cudaError_t launch_refine(const HbmaParams& p, uint32_t lvl, cudaStream_t st) {
  using G = RefineGeom<RC, NB, NDY, B>;
  static_assert(G::SMEM <= 227 * 1024 && G::PT <= 256, "refine geometry does not fit");
  EbmaMaps maps;
  const uint32_t n_slots = p.n_frames + 1;
  const uint8_t* base = p.pyr + p.lay.off[lvl];
  if (!encode_box(&maps.t, base, p.lay.w[lvl], p.lay.h[lvl], p.lay.pitch[lvl], p.lay.slot_bytes, n_slots, G::PT,
                  B + 2 * p.r) ||
      !encode_box(&maps.a, base, p.lay.w[lvl], p.lay.h[lvl], p.lay.pitch[lvl], p.lay.slot_bytes, n_slots, 16, B))
    return cudaErrorNotSupported;
  auto kern = hbma_refine_kernel<RC, NB, NDY, THREADS, MINB, B>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return e;
  const uint64_t n_blocks = (uint64_t)p.mvw * p.mvh * p.n_frames;
  kern<<<(uint32_t)((n_blocks + NB - 1) / NB), THREADS, G::SMEM, st>>>(maps, p, (int)lvl);
  return cudaGetLastError();
}

// top level with block size BT = 16 >> (L-1), then the refinement levels down to 16x16
template <int RC, int NDY, int NB_REF, int THR_REF, int MINB_REF>
static cudaError_t launch_levels(const HbmaParams& p, cudaStream_t st, int* n_launches) {
  const uint32_t L = p.lay.levels;
  constexpr int PAW = RC <= 16 ? 80 : 48;  // anchor tile width of the top-level tiles (bytes)
  cudaError_t e;
  switch (L) {
    case 2: e = launch_top_tile<RC, PAW / 8, NDY, 256, 3, 8>(p, 1, st); break;
    case 3: e = launch_top_tile<RC, PAW / 4, NDY, 256, 3, 4>(p, 2, st); break;
    case 4: e = launch_top_tile<RC, PAW / 2, NDY, 256, 3, 2>(p, 3, st); break;
    default: e = launch_top_tile<RC, PAW, NDY, 256, 3, 1>(p, 4, st); break;
  }
  if (e != cudaSuccess) return e;
  if (L >= 5) {
    e = launch_refine<RC, NB_REF, NDY, THR_REF, 4, 2>(p, 3, st);
    if (e != cudaSuccess) return e;
  }
  if (L >= 4) {
    e = launch_refine<RC, NB_REF, NDY, THR_REF, 4, 4>(p, 2, st);
    if (e != cudaSuccess) return e;
  }
  if (L >= 3) {
    e = launch_refine<RC, NB_REF, NDY, THR_REF, 4, 8>(p, 1, st);
    if (e != cudaSuccess) return e;
  }
  e = launch_refine<RC, NB_REF, NDY, THR_REF, MINB_REF, 16>(p, 0, st);
  if (n_launches) *n_launches += (int)L - 1;  // the caller counts one
  return e;
}

template <int RC, int CW, int NDY, int THREADS, int MINB>
static cudaError_t launch_ebma_stripe(const HbmaParams& p, cudaStream_t st) {
  using G = StripeGeomE<RC, CW, NDY>;
  static_assert(G::SMEM <= 227 * 1024 && G::PT <= 256 && 16 + 2 * RC <= 256, "stripe geometry does not fit");
  EbmaMaps maps;
  const uint32_t n_slots = p.n_frames + 1;
  const uint8_t* base = p.pyr + p.lay.off[0];
  if (!encode_box(&maps.t, base, p.lay.w[0], p.lay.h[0], p.lay.pitch[0], p.lay.slot_bytes, n_slots, G::PT,
                  16 + 2 * p.r) ||
      !encode_box(&maps.a, base, p.lay.w[0], p.lay.h[0], p.lay.pitch[0], p.lay.slot_bytes, n_slots, 16, 16))
    return cudaErrorNotSupported;
  auto kern = hbma_ebma_stripe_kernel<RC, CW, NDY, THREADS, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return e;
  const uint64_t n_blocks = (uint64_t)p.mvw * p.mvh * p.n_frames;
  kern<<<(uint32_t)n_blocks, THREADS, G::SMEM, st>>>(maps, p);
  return cudaGetLastError();
}

template <int RC, int NBX, int NDY, int THREADS, int MINB>
static cudaError_t launch_ebma_tile(const HbmaParams& p, cudaStream_t st) {
  using G = TileGeomE<RC, NBX, NDY>;
  static_assert(G::SMEM <= 227 * 1024 && G::PT <= 256 && G::PA <= 256, "tile geometry does not fit");
  EbmaMaps maps;
  const uint32_t n_slots = p.n_frames + 1;
  const uint8_t* base = p.pyr + p.lay.off[0];
  if (!encode_box(&maps.t, base, p.lay.w[0], p.lay.h[0], p.lay.pitch[0], p.lay.slot_bytes, n_slots, G::PT,
                  16 + 2 * p.r) ||
      !encode_box(&maps.a, base, p.lay.w[0], p.lay.h[0], p.lay.pitch[0], p.lay.slot_bytes, n_slots, G::PA, 16))
    return cudaErrorNotSupported;
  auto kern = hbma_ebma_tile_kernel<RC, NBX, NDY, THREADS, MINB, 16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return e;
  const uint64_t tiles = (uint64_t)((p.mvw + NBX - 1) / NBX) * p.mvh * p.n_frames;
  kern<<<(uint32_t)tiles, THREADS, G::SMEM, st>>>(maps, p, 0);
  return cudaGetLastError();
}

template <int RC, int NB, int NDY, int THREADS, int MINB>
static cudaError_t launch_pool(const HbmaParams& p, cudaStream_t st) {
  using G = PoolGeom<RC, NB, NDY>;
  static_assert(G::SMEM <= 227 * 1024, "pool geometry does not fit");
  static_assert(G::PT <= 256, "TMA box limit");
  const uint32_t L = p.lay.levels, r = p.r;
  PoolMaps maps;
  const uint32_t n_slots = p.n_frames + 1;
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t B = 16u >> l;
    const uint8_t* base = p.pyr + p.lay.off[l];
    if (!encode_box(&maps.t[l], base, p.lay.w[l], p.lay.h[l], p.lay.pitch[l], p.lay.slot_bytes, n_slots,
                    G::PT, B + 2 * r) ||
        !encode_box(&maps.a[l], base, p.lay.w[l], p.lay.h[l], p.lay.pitch[l], p.lay.slot_bytes, n_slots, 16, B))
      return cudaErrorNotSupported;
  }
  auto kern = hbma_pool_kernel<RC, NB, NDY, THREADS, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return e;
  const uint64_t n_blocks = (uint64_t)p.mvw * p.mvh * p.n_frames;
  kern<<<(uint32_t)((n_blocks + NB - 1) / NB), THREADS, G::SMEM, st>>>(maps, p);
  return cudaGetLastError();
}

bool try_launch_pool(const HbmaParams& p, cudaStream_t st, cudaError_t* err, int* extra_launches) {
  const uint32_t L = p.lay.levels, r = p.r;
  // p.family == kHbmaPool (test hook): the single-launch pooled kernel wherever it is defined, so that
  // its parity tests also cover configurations the dispatcher gives to faster kernels
  const bool force = p.family == kHbmaPool;
  // r <= 4: hbma_tile_kernel, or (5 levels, r = 3, 4: the reach does not fit a tile) the hybrid below
  if (!force && p.bw == 16 && p.bh == 16 && L == 5 && (r == 3 || r == 4) && p.mv && p.mad &&
      p.n_frames <= 65535 && (uint64_t)p.mvw * p.mvh * p.n_frames <= 0x7fffffffull) {
    // tile kernel over levels 4..2, then one refinement launch per remaining level (k_hbma_rs.cu)
    *err = launch_tile_upper3(p, st);
    if (*err == cudaSuccess) *err = launch_rs_level(p, 1, false, st);
    if (*err == cudaSuccess) *err = launch_rs_level(p, 0, false, st);
    if (extra_launches) *extra_launches += 2;
    return true;
  }
  if (p.bw != 16 || p.bh != 16 || L > 5 || r < 5 || r > (L == 1 ? 112u : 64u)) return false;
  if ((uint64_t)p.mvw * p.mvh * p.n_frames > 0x7fffffffull) return false;
  if (r <= 16 && (L >= 2 || r <= 8) && !force && rs_level_supported(p)) {
    // windows of at most 33 x 33 candidates (r = 5..16): one launch per level, windows realigned in
    // registers (k_hbma_rs.cu); the top level -- or the only one, r <= 8: L = 1, plain EBMA -- with the
    // "<=" scan rule.  (r = 3, 4 stay on the bounded-reach tile kernel: 41-50 % of the SAD peak against
    // 30-35 % here; L = 1 with r = 9..16 on the shared-window tile kernel below: 69 %.)
    for (int l = (int)L - 1; l >= 0 && *err == cudaSuccess; --l) *err = launch_rs_level(p, (uint32_t)l, l == (int)L - 1, st);
    if (extra_launches) *extra_launches += (int)L - 1;  // the caller counts one
    return true;
  }
  if (L >= 2 && r <= 32 && p.mv && p.mad && !force) {
    // one launch per level; <range class, candidate rows per item, blocks per CTA and CTAs per SM of
    // the 16x16 refinement level>
    *err = launch_levels<32, 13, 2, 128, 3>(p, st, extra_launches);  // r = 17..32
    return true;
  }
  // <range class, ..., candidate rows per item, threads, CTAs per SM>: measured best of several shapes
  // per class on B200 (profiles/r01_sweep_hbma_v12.md)
  if (L == 1 && r > 32 && !force) {
    // <range class, stripe width in candidate columns, candidate rows per item, threads, CTAs per SM>
    if (r > 64) *err = launch_ebma_stripe<112, 65, 13, 128, 2>(p, st);
    else *err = launch_ebma_stripe<64, 65, 13, 128, 3>(p, st);
    return true;
  }
  if (L == 1 && r <= 32 && !force) {
    // <range class, blocks per tile, candidate rows per item, threads, CTAs per SM>
    if (r <= 8) *err = launch_ebma_tile<8, 5, 9, 96, 6>(p, st);
    else if (r <= 16) *err = launch_ebma_tile<16, 5, 11, 128, 4>(p, st);
    else *err = launch_ebma_tile<32, 3, 13, 128, 4>(p, st);
    return true;
  }
  if (r > 64) return false;  // L = 1, r = 65..112 exists only as the striped kernel
  // Deep pyramids with a small top-level range and only one output wanted (no level-synchronous path):
  // the warp-per-block window kernel of k_hbma.cu, which has no block-wide barrier, measures faster.
  if (L >= 3 && r <= 16 && !force) return false;
  // <range class, blocks per CTA, candidate rows per item, threads, CTAs per SM>
  if (r <= 8) *err = launch_pool<8, 7, 17, 128, 3>(p, st);
  else if (r <= 16) *err = launch_pool<16, 3, 11, 128, 4>(p, st);
  else if (r <= 32) *err = launch_pool<32, 2, 13, 128, 3>(p, st);
  else *err = launch_pool<64, 1, 13, 256, 2>(p, st);
  return true;
}

}  // namespace svc
