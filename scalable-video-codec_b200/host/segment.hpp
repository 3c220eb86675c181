// segment.hpp -- the block-type stages that consume the motion field:
// RANSAC global motion (libs/motion.cpp:157-266), foreground mask + morphological
// close/open (libs/encoder.cpp:507-522), k-means over the foreground vectors
// (libs/encoder.cpp:296-321, 557-578) and connected components per cluster
// (libs/encoder.cpp:597-624).
//
// The reference runs these on the CPU with OpenCV between the motion search and the DCT;
// this image has no C++ OpenCV, so morphologyEx / kmeans / connectedComponents are
// restated here from their documented algorithms and pinned against python cv2 in
// tests/test_segment.py (bit-exact labels for the same generator state).  They stay on
// the host: a 1080p motion field is 8 160 vectors, every stage is microseconds, and
// svc::Encoder spreads the frames of a GPU batch over worker threads.
#ifndef SVC_B200_HOST_SEGMENT_HPP
#define SVC_B200_HOST_SEGMENT_HPP

#include <cstdint>
#include <vector>

#include "motion.hpp"

#ifndef SCALABLE_VIDEO_CODEC_MOTION_HPP
struct RansacParams {  // libs/motion.hpp:60-80
  uint subset_sz;
  float inlier_thresh;
  float success_prob;
  float inlier_ratio;
};
#endif

#ifndef SCALABLE_VIDEO_CODEC_ENCODER_HPP
struct KMeansParams {  // libs/encoder.hpp:16-21
  uint cluster_count;
  uint attempt_count;
  uint max_iter_count;
  float epsilon;
};
#endif

// Reference signatures (libs/motion.hpp:40, 99-103).  Like the reference, the RANSAC
// entry draws from one process-wide std::minstd_rand0 seeded from std::random_device on
// first use; svc::SeedGlobalMotionRng() pins that seed.
Vec2f EstimateGlobalMotionAvg(const Vec2f* motion_field, uint sz);
void EstimateGlobalMotionRansac(const Vec2f* motion_field, uint motion_field_sz, RansacParams params,
                                float* rmse, Vec2f* global_motion, std::vector<uint>* inlier_indices);

namespace svc {

void SeedGlobalMotionRng(uint32_t seed);

// std::minstd_rand0 (libstdc++'s std::default_random_engine) and the way libstdc++'s
// std::uniform_int_distribution<uint>(0, n) consumes it, restated so that the sample
// sequence does not depend on the standard library the host code is built with.
struct MinstdRand0 {
  uint32_t state;
  explicit MinstdRand0(uint32_t seed = 1u) { Seed(seed); }
  void Seed(uint32_t seed) {
    state = seed % 2147483647u;
    if (state == 0) state = 1;
  }
  uint32_t Next() {
    state = (uint32_t)(((uint64_t)state * 16807u) % 2147483647u);
    return state;
  }
  uint32_t UniformInclusive(uint32_t hi);  // in [0, hi]
};

// cv::RNG (multiply-with-carry), as used by cv::kmeans through cv::theRNG().
struct CvRng {
  uint64_t state;
  explicit CvRng(uint64_t s = 0xffffffffu) : state(s ? s : 0xffffffffu) {}
  uint32_t Next() {
    state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32);
    return (uint32_t)state;
  }
  double NextDouble() {
    const uint32_t t = Next();
    return (double)(((uint64_t)t << 32) | Next()) * 5.4210108624275221700372640043497e-20;
  }
};

void RansacGlobalMotion(const Vec2f* motion_field, uint n, const RansacParams& params, MinstdRand0& rng,
                        float* rmse, Vec2f* global_motion, std::vector<uint>* inliers);

enum MorphOp : uint { kErode = 0, kDilate = 1, kOpen = 2, kClose = 3 };  // cv::MorphTypes
void MorphologyEx(uint8_t* mask, uint w, uint h, uint op, uint rect_w, uint rect_h);

// returns cv::connectedComponents' return value (number of components + 1)
uint ConnectedComponents(const uint8_t* mask, uint w, uint h, uint connectivity, int32_t* labels);

// cv::kmeans with KMEANS_PP_CENTERS and TermCriteria(COUNT | EPS, max_iter, eps); returns the compactness
double KMeans(const float* data, int n, int dims, int k, int max_iter, double eps, int attempts, CvRng& rng,
              int32_t* labels, float* centers /* k x dims or null */);

struct SegmentConfig {  // EncoderConfig fields of these stages; defaults apps/encoder.cpp:28-58
  RansacParams ransac{1, 7.5f, 0.99f, 0.5f};
  uint morph_rect_w = 3, morph_rect_h = 3;
  KMeansParams kmeans{10, 3, 10, 1.0f};
  uint connected_components_connectivity = 4;
  uint mv_block_w = 16, mv_block_h = 16;
};

// "" when valid, else the message of the reference's Validate (libs/encoder.cpp:20-60, 86-101)
const char* ValidateSegmentConfig(const SegmentConfig& cfg);

// libs/encoder.cpp:491-624 for one motion field.  Not thread safe per object (scratch
// buffers); one object per worker thread.
class MotionSegmenter {
 public:
  MotionSegmenter(const SegmentConfig& cfg, uint mv_field_w, uint mv_field_h);
  void operator()(const Vec2f* mv_field, MinstdRand0& ransac_rng, CvRng& kmeans_rng, uint* block_types,
                  Vec2f* global_motion = nullptr);

 private:
  SegmentConfig cfg_;
  uint w_, h_;
  std::vector<uint> inliers_, fg_;
  std::vector<uint8_t> mask_, cluster_mask_;
  std::vector<float> features_;
  std::vector<int32_t> cluster_ids_, comp_ids_;
};

// The stage as svc::Encoder runs it: the motion fields of one GPU batch are labelled on
// worker threads.  Every encoded frame gets its own generator states derived from
// (seed, frame index), so the labels do not depend on the batch size, the thread count or
// the frame-range sharding (the reference's process-global generators make its labels
// differ run to run).
class BlockTypeStage {
 public:
  // seed 0: taken from std::random_device once (the reference's behaviour); threads 0: min(8, cores)
  BlockTypeStage(const SegmentConfig& cfg, uint mv_field_w, uint mv_field_h, uint64_t seed, uint threads);
  // mv_fields: n x (w*h) vectors of encoded frames [first_frame, first_frame + n); block_types: n x (w*h)
  void Run(const Vec2f* mv_fields, uint n, uint64_t first_frame, uint* block_types);
  uint64_t seed() const { return seed_; }
  static void FrameGenerators(uint64_t seed, uint64_t frame, MinstdRand0* ransac_rng, CvRng* kmeans_rng);

 private:
  SegmentConfig cfg_;
  uint w_, h_, threads_;
  uint64_t seed_;
  std::vector<MotionSegmenter> workers_;
};

}  // namespace svc

#endif  // SVC_B200_HOST_SEGMENT_HPP
