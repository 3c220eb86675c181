// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" forwarding shims so that python (ctypes) can call the UNMODIFIED
// reference motion code.  Compiled together with
// /root/reference/libs/motion.cpp (sources stay where they lie; nothing is
// copied into this repository) into oracle/_ref/libref_motion.so by
// oracle/Makefile.  Used by tests/ to pin the oracle and by
// `bench.py --impl reference` as the CPU baseline of kind "reference".
#include "motion.hpp"  // -I$(REF)/libs

extern "C" {

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
