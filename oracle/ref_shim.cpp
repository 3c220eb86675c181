// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" forwarding shims so that python (ctypes) can call the UNMODIFIED
// reference motion code.  Compiled together with
// /root/reference/libs/motion.cpp (sources stay where they lie; nothing is
// copied into this repository) into oracle/_ref/libref_motion.so by
// oracle/Makefile.  Used by tests/ to pin the oracle and by
// `bench.py --impl reference` as the CPU baseline of kind "reference".
#include <random>
#include <vector>

#include "motion.hpp"  // -I$(REF)/libs

// EstimateGlobalMotionRansac seeds a function-static std::default_random_engine from
// std::random_device (libs/motion.cpp:186-187).  To pin that seed WITHOUT touching the reference
// source, this library defines std::random_device::_M_getval() itself and is linked with
// -Wl,-Bsymbolic, so the reference's rdev() call binds here.  The engine is seeded on the first
// call per loaded copy of the library: a test that wants another seed loads another copy.
static unsigned g_ref_seed = 1;
unsigned int std::random_device::_M_getval() { return g_ref_seed; }

extern "C" {

void ref_set_seed(unsigned seed) { g_ref_seed = seed; }

// libs/motion.hpp:99-103; inliers: room for n entries; mv_xy must hold n + 1 vectors (the
// reference samples index n too, libs/motion.cpp:208)
void ref_ransac(const float* mv_xy, uint n, uint subset_sz, float inlier_thresh, float success_prob,
                float inlier_ratio, float* rmse, float* gm_xy, uint* inliers, uint* n_inliers) {
  RansacParams p;
  p.subset_sz = subset_sz;
  p.inlier_thresh = inlier_thresh;
  p.success_prob = success_prob;
  p.inlier_ratio = inlier_ratio;
  std::vector<uint> in;
  Vec2f gm{gm_xy[0], gm_xy[1]};
  EstimateGlobalMotionRansac(reinterpret_cast<const Vec2f*>(mv_xy), n, p, rmse, &gm, &in);
  gm_xy[0] = gm.x;
  gm_xy[1] = gm.y;
  *n_inliers = (uint)in.size();
  for (size_t i = 0; i < in.size(); ++i) inliers[i] = in[i];
}

// libs/motion.hpp:40
void ref_global_motion_avg(const float* mv_xy, uint n, float* gm_xy) {
  const Vec2f g = EstimateGlobalMotionAvg(reinterpret_cast<const Vec2f*>(mv_xy), n);
  gm_xy[0] = g.x;
  gm_xy[1] = g.y;
}

// libs/motion.hpp:106-110
void ref_ebma(const uchar* tracked, const uchar* anchor, uint fw, uint fh,
              uint search_range, uint bw, uint bh, float* mv_xy,
              float* min_mad) {
  EstimateMotionExhaustiveSearch(tracked, anchor, fw, fh, search_range, bw, bh,
                                 reinterpret_cast<Vec2f*>(mv_xy), min_mad);
}

// libs/motion.hpp:134-138
void ref_hbma(const uchar* const* tracked_pyr, const uchar* const* anchor_pyr,
              uint levels, uint fw, uint fh, uint search_range, uint bw,
              uint bh, float* mv_xy, float* min_mad) {
  EstimateMotionHierarchical(tracked_pyr, anchor_pyr, levels, fw, fh,
                             search_range, bw, bh,
                             reinterpret_cast<Vec2f*>(mv_xy), min_mad);
}

// libs/motion.hpp:148-152
void ref_hbma_16x16_sse2(const uchar* const* tracked_pyr,
                         const uchar* const* anchor_pyr, uint fw, uint fh,
                         uint search_range, float* mv_xy, float* min_mad) {
#ifdef __SSE2__
  EstimateMotionHierarchical16x16Sse2(tracked_pyr, anchor_pyr, fw, fh,
                                      search_range,
                                      reinterpret_cast<Vec2f*>(mv_xy), min_mad);
#else
  EstimateMotionHierarchical(tracked_pyr, anchor_pyr, 4, fw, fh, search_range,
                             16, 16, reinterpret_cast<Vec2f*>(mv_xy), min_mad);
#endif
}

int ref_has_sse2() {
#ifdef __SSE2__
  return 1;
#else
  return 0;
#endif
}
}
