// motion.cpp -- forwards the reference motion signatures to the C ABI.
#include "motion.hpp"

#include "../../include/svc_b200.h"

static_assert(sizeof(Vec2f) == 2 * sizeof(float), "Vec2f must be two packed floats");

static void check(int rc) {
  if (rc != SVC_OK) throw svc::Error(rc, svc_last_error());
}

void EstimateMotionExhaustiveSearch(const uchar* tracked_frame, const uchar* anchor_frame,
                                    uint frame_w, uint frame_h, uint search_range, uint block_w,
                                    uint block_h, Vec2f* motion_field, float* min_mad) {
  check(svc_estimate_motion_exhaustive(tracked_frame, anchor_frame, frame_w, frame_h, search_range,
                                       block_w, block_h, reinterpret_cast<float*>(motion_field),
                                       min_mad));
}

void EstimateMotionHierarchical(const uchar* const* tracked_pyramid,
                                const uchar* const* anchor_pyramid, uint level_count, uint frame_w,
                                uint frame_h, uint search_range, uint block_w, uint block_h,
                                Vec2f* motion_field, float* min_mad) {
  check(svc_estimate_motion_hierarchical(tracked_pyramid, anchor_pyramid, level_count, frame_w,
                                         frame_h, search_range, block_w, block_h,
                                         reinterpret_cast<float*>(motion_field), min_mad));
}

void EstimateMotionHierarchical16x16Sse2(const uchar* const* tracked_pyramid,
                                         const uchar* const* anchor_pyramid, uint frame_w,
                                         uint frame_h, uint search_range, Vec2f* mv_field,
                                         float* min_mad) {
  check(svc_estimate_motion_hierarchical_16x16(tracked_pyramid, anchor_pyramid, frame_w, frame_h,
                                               search_range, reinterpret_cast<float*>(mv_field),
                                               min_mad));
}
