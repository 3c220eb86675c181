/* svc_segment.h -- C ABI of the block-type stages that consume the motion field
 * (SURVEY.md section 8f, rank 2): RANSAC global motion, foreground mask + morphology,
 * k-means over the foreground vectors, connected components per cluster.
 *
 * These are the CPU consumers of the GPU hot path's motion vectors in the reference
 * (libs/encoder.cpp:491-624, libs/motion.cpp:157-266); they stay on the host here too
 * (8 160 vectors per 1080p frame: tens of microseconds per stage), batched over the
 * frames of a GPU batch by svc::Encoder's worker threads.  Exported by
 * scalable-video-codec_b200/lib/libsvc_host.so (C++ face: host/segment.hpp).
 *
 * The reference draws its random numbers from process-global generators
 * (std::default_random_engine seeded by std::random_device, libs/motion.cpp:186-187;
 * cv::theRNG() inside cv::kmeans, libs/encoder.cpp:574-575).  Every entry point here
 * takes the generator state explicitly (in/out), so a caller can reproduce the
 * reference's call-after-call sequence from a known seed, or give each frame its own.
 *
 * Return value: 0 = ok, 1 = invalid argument (message: svc_seg_last_error()).
 */
#ifndef SVC_SEGMENT_H
#define SVC_SEGMENT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::MorphTypes values used by the reference (libs/encoder.cpp:519-522) */
#define SVC_MORPH_ERODE 0u
#define SVC_MORPH_DILATE 1u
#define SVC_MORPH_OPEN 2u
#define SVC_MORPH_CLOSE 3u

/* EncoderConfig fields of these stages (libs/encoder.hpp:25-37, libs/motion.hpp:60-80);
 * defaults apps/encoder.cpp:28-58. */
typedef struct svc_seg_config {
  uint32_t ransac_subset_sz;     /* 1 */
  float ransac_inlier_thresh;    /* 7.5 */
  float ransac_success_prob;     /* 0.99 */
  float ransac_inlier_ratio;     /* 0.5 */
  uint32_t morph_rect_w;         /* 3 */
  uint32_t morph_rect_h;         /* 3 */
  uint32_t kmeans_cluster_count; /* 10 */
  uint32_t kmeans_attempt_count; /* 3 */
  uint32_t kmeans_max_iter_count;/* 10 */
  float kmeans_epsilon;          /* 1 */
  uint32_t connected_components_connectivity; /* 4 (or 8) */
  uint32_t mv_block_w;           /* 16 */
  uint32_t mv_block_h;           /* 16 */
} svc_seg_config;

void svc_seg_default_config(svc_seg_config* cfg);

/* Validate(RansacParams) / Validate(KMeansParams) and the morphology / connectivity checks of
 * Validate(EncoderConfig), libs/encoder.cpp:20-142: 0 when valid, else 1 and the reference's message. */
int svc_seg_validate(const svc_seg_config* cfg);

/* EstimateGlobalMotionRansac, libs/motion.hpp:99-103 / libs/motion.cpp:182-266.
 * rng_state: state of std::minstd_rand0 (= std::default_random_engine of libstdc++); pass the
 * seed on the first call (0 is mapped to 1 like std::linear_congruential_engine::seed).
 * global_motion_xy is in/out (the reference reads its previous value on one degenerate path).
 * inliers: room for n indices. */
int svc_seg_ransac(const float* mv_xy, uint32_t n, uint32_t subset_sz, float inlier_thresh,
                   float success_prob, float inlier_ratio, uint32_t* rng_state, float* rmse,
                   float* global_motion_xy, uint32_t* inliers, uint32_t* n_inliers);

/* EstimateGlobalMotionAvg, libs/motion.hpp:40 / libs/motion.cpp:45-53 */
int svc_seg_global_motion_avg(const float* mv_xy, uint32_t n, float* global_motion_xy);

/* cv::morphologyEx(mask, mask, op, getStructuringElement(MORPH_RECT, {rect_w, rect_h})) with the
 * default anchor, one iteration and the default (ignored) border; in place. libs/encoder.cpp:519-522 */
int svc_seg_morphology(uint8_t* mask, uint32_t w, uint32_t h, uint32_t op, uint32_t rect_w,
                       uint32_t rect_h);

/* cv::connectedComponents(mask, labels, connectivity, CV_32S): labels 0 = background, components
 * numbered in OpenCV's order; *n_labels = the function's return value (components + 1).
 * libs/encoder.cpp:607-611 */
int svc_seg_connected_components(const uint8_t* mask, uint32_t w, uint32_t h, uint32_t connectivity,
                                 int32_t* labels, uint32_t* n_labels);

/* cv::kmeans(data (n x dims, f32), k, labels, TermCriteria(COUNT|EPS, max_iter, eps), attempts,
 * KMEANS_PP_CENTERS[, centers]); rng_state: state of cv::RNG (cv::theRNG() starts at 0xffffffff;
 * cv::setRNGSeed(s) sets it to s).  centers (k x dims) and compactness may be null.
 * libs/encoder.cpp:570-575 */
int svc_seg_kmeans(const float* data, uint32_t n, uint32_t dims, uint32_t k, uint32_t max_iter,
                   float eps, uint32_t attempts, uint64_t* rng_state, int32_t* labels, float* centers,
                   double* compactness);

/* The whole chain of libs/encoder.cpp:491-624 for one motion field: block_types receives
 * mv_field_w * mv_field_h labels (0 = BLOCK_TYPE_BACKGROUND, libs/codec.hpp:6).
 * global_motion_xy may be null. */
int svc_seg_block_types(const float* mv_xy, uint32_t mv_field_w, uint32_t mv_field_h,
                        const svc_seg_config* cfg, uint32_t* ransac_rng_state,
                        uint64_t* kmeans_rng_state, uint32_t* block_types, float* global_motion_xy);

/* Generator states svc::Encoder gives encoded frame `frame` (0-based) for a stream seed: labels
 * are then independent of batch size, worker threads and frame-range sharding. */
void svc_seg_frame_generators(uint64_t seed, uint64_t frame, uint32_t* ransac_rng_state,
                              uint64_t* kmeans_rng_state);

/* svc::BlockTypeStage: n motion fields (frames first_frame .. first_frame + n - 1) labelled on
 * `threads` worker threads (0 = default) with the per-frame generators above. */
int svc_seg_block_types_batch(const float* mv_xy, uint32_t n, uint32_t mv_field_w, uint32_t mv_field_h,
                              const svc_seg_config* cfg, uint64_t seed, uint64_t first_frame,
                              uint32_t threads, uint32_t* block_types);

const char* svc_seg_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* SVC_SEGMENT_H */
