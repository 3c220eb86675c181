// motion.hpp -- C++ host face of the B200 motion kernels.
//
// Same names, argument order, argument meaning and output layout as the
// reference's libs/motion.hpp (:106-110, :134-138, :148-152), so
// Encoder::operator() (libs/encoder.cpp:472-482) and every consumer of the
// motion field (RANSAC libs/encoder.cpp:495-497, draw libs/draw.cpp:55-89)
// compile and run unchanged against this header.  The bodies forward to the C
// ABI (include/svc_b200.h) -- there is no CPU implementation behind them.
//
// Error behaviour: the reference functions return void and assert() their
// preconditions (libs/motion.cpp:417-433, 701-712).  Here a violated
// precondition or a CUDA failure throws svc::Error (derived from
// std::runtime_error) carrying the C ABI status code and message.
#ifndef SVC_B200_HOST_MOTION_HPP
#define SVC_B200_HOST_MOTION_HPP

#include <stdexcept>
#include <string>

// When the reference's own math.hpp / types.hpp are on the include path and were
// included first, their Vec2f / uint / uchar are used; otherwise the same POD
// types are provided here (ABI: struct { float x, y; }, libs/math.hpp:177-181).
#ifndef SCALABLE_VIDEO_CODEC_TYPES_HPP
typedef unsigned int uint;
typedef unsigned char uchar;
#endif
#ifndef SCALABLE_VIDEO_CODEC_MATH_HPP
struct Vec2f {
  float x;
  float y;
  float& operator[](uint i) { return (&x)[i]; }
};
#endif

namespace svc {
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
}  // namespace svc

/* Calculates the motion field using EBMA (libs/motion.hpp:106-110). */
void EstimateMotionExhaustiveSearch(const uchar* tracked_frame, const uchar* anchor_frame,
                                    uint frame_w, uint frame_h, uint search_range, uint block_w,
                                    uint block_h, Vec2f* motion_field, float* min_mad);

/* Calculates the motion field using the reference's HBMA variation
   (libs/motion.hpp:112-138): top-level range search_range / 2^(level_count-1),
   the same range again at every finer level. */
void EstimateMotionHierarchical(const uchar* const* tracked_pyramid,
                                const uchar* const* anchor_pyramid, uint level_count, uint frame_w,
                                uint frame_h, uint search_range, uint block_w, uint block_h,
                                Vec2f* motion_field, float* min_mad);

/* 4 levels, 16x16 blocks (libs/motion.hpp:148-152).  The name is kept so the
   call site at libs/encoder.cpp:472-476 needs no edit; nothing here is SSE2. */
void EstimateMotionHierarchical16x16Sse2(const uchar* const* tracked_pyramid,
                                         const uchar* const* anchor_pyramid, uint frame_w,
                                         uint frame_h, uint search_range, Vec2f* mv_field,
                                         float* min_mad);

#endif  // SVC_B200_HOST_MOTION_HPP
