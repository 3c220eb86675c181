// encoder.hpp -- C++ host driver of the device-resident hot path: the
// counterpart of the reference's Encoder functor (libs/encoder.hpp:52-95,
// libs/encoder.cpp:341-671) with its OpenCV buffers replaced by one
// svc_session on one GPU.
//
// Same contract towards the application (apps/encoder.cpp:172-228): frames
// are popped from an input queue until the producer is done, the first byte
// buffer pushed to the output queue is the 32-byte Header
// (libs/encoder.cpp:361-381), then one serialised frame per encoded frame
// (libs/encoder.cpp:647-652); the first input frame is tracked-only.  What
// differs, because the image has no C++ OpenCV: frames are raw interleaved
// 8-bit BGR buffers instead of cv::Mat3b, and the CPU stages that turn a
// motion field into block types (RANSAC .. connected components,
// libs/encoder.cpp:491-624) are the restatements of host/segment.hpp, run on
// worker threads over each GPU batch (EncoderConfig::segment; an application
// can substitute its own callback).  The labels are patched into the records
// the GPU already wrote (svc_patch_block_types), so the consumers downstream
// see the reference layout.
#ifndef SVC_B200_HOST_ENCODER_HPP
#define SVC_B200_HOST_ENCODER_HPP

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include <memory>

#include "motion.hpp"
#include "segment.hpp"

struct svc_session;

namespace svc {

enum class ErrorCode { kOk, kUnspecified, kInvalidParameter };  // libs/error.hpp:6

struct Status {  // libs/error.hpp:8-11 (Error{code, message})
  ErrorCode code;
  std::string message;
};

// Hot-path subset of EncoderConfig (libs/encoder.hpp:25-37); same field names.
struct EncoderConfig {
  uint mv_block_w = 16;         // apps/encoder.cpp:42-58 defaults
  uint mv_block_h = 16;
  uint mv_search_range = 8;
  uint pyr_lvl_count = 4;
  uint transform_block_w = 8;
  uint transform_block_h = 8;
  int device = 0;
  uint max_batch = 32;  // frames per kernel launch
  // Block-type stages (libs/encoder.cpp:491-624: RANSAC, morphology, k-means, connected components),
  // run on host worker threads over the motion fields of each GPU batch; fields as in the
  // reference's EncoderConfig (ransac, morph_rect_w/h, kmeans, connected_components_connectivity).
  bool segment = true;
  SegmentConfig seg;
  uint64_t seed = 0;          // 0: from std::random_device like the reference; else reproducible labels
  uint classify_threads = 0;  // 0: min(8, hardware threads)
};

struct VideoProperties {  // libs/encoder.hpp:46-50
  uint frame_w;
  uint frame_h;
  uint frame_count;
};

// Validate(const EncoderConfig&), libs/encoder.cpp:62-142 (hot-path fields).
Status Validate(const EncoderConfig&);

// Bounded blocking queue with the reference's CircularQueue semantics
// (libs/queue.hpp:13-83): Push blocks when full, Pop returns false once the
// queue is empty and the producer signalled completion.  Added for error paths
// (the reference has none): Close() cancels the queue -- it wakes both sides,
// Push returns false from then on and Pop returns false immediately, so a
// producer blocked on a full queue and a consumer blocked on an empty one both
// leave when the stage between them failed.
template <typename T>
class BoundedQueue {
 public:
  explicit BoundedQueue(size_t capacity) : cap_(capacity) {}
  bool Push(T v) {
    std::unique_lock<std::mutex> l(m_);
    not_full_.wait(l, [&] { return q_.size() < cap_ || closed_; });
    if (closed_) return false;
    q_.push_back(std::move(v));
    not_empty_.notify_one();
    return true;
  }
  bool Pop(T& out) {
    std::unique_lock<std::mutex> l(m_);
    not_empty_.wait(l, [&] { return !q_.empty() || done_ || closed_; });
    if (closed_ || q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    not_full_.notify_one();
    return true;
  }
  // Non-blocking variant used to top up a batch without stalling the GPU.
  bool TryPop(T& out) {
    std::lock_guard<std::mutex> l(m_);
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    not_full_.notify_one();
    return true;
  }
  void SignalProducerIsDone() {
    std::lock_guard<std::mutex> l(m_);
    done_ = true;
    not_empty_.notify_all();
  }
  void Close() {
    std::lock_guard<std::mutex> l(m_);
    closed_ = done_ = true;
    not_empty_.notify_all();
    not_full_.notify_all();
  }
  bool closed() {
    std::lock_guard<std::mutex> l(m_);
    return closed_;
  }

 private:
  size_t cap_;
  std::deque<T> q_;
  bool done_ = false, closed_ = false;
  std::mutex m_;
  std::condition_variable not_full_, not_empty_;
};

using Frame = std::vector<uchar>;  // frame_h * frame_w * 3 interleaved BGR
using Bytes = std::vector<uchar>;

// mv_field / min_mad: mv_field_w * mv_field_h entries of one encoded frame;
// block_types (out): same count, preset to BLOCK_TYPE_BACKGROUND (0).
using BlockTypeFn = std::function<void(const Vec2f* mv_field, const float* min_mad, uint mv_field_w,
                                       uint mv_field_h, uint* block_types)>;

class Encoder {
 public:
  Encoder(const EncoderConfig& cfg, const VideoProperties& vidprops, BoundedQueue<Frame>& in_queue,
          BoundedQueue<Bytes>& out_queue, BlockTypeFn classify = nullptr);
  ~Encoder();
  Encoder(const Encoder&) = delete;
  Encoder& operator=(const Encoder&) = delete;
  void operator()();

  uint padded_frame_w() const { return padded_frame_w_; }
  uint padded_frame_h() const { return padded_frame_h_; }
  uint mv_field_w() const { return mv_field_w_; }
  uint mv_field_h() const { return mv_field_h_; }
  uint64_t frames_encoded() const { return frames_encoded_; }

 private:
  EncoderConfig cfg_;
  VideoProperties vidprops_;
  BoundedQueue<Frame>& in_queue_;
  BoundedQueue<Bytes>& out_queue_;
  BlockTypeFn classify_;
  std::unique_ptr<BlockTypeStage> stage_;
  struct SessionDeleter { void operator()(svc_session* s) const; };
  struct PinnedDeleter { void operator()(void* p) const; };
  template <class T> using Pinned = std::unique_ptr<T, PinnedDeleter>;
  // RAII members: a constructor that throws half-way releases what it already acquired
  std::unique_ptr<svc_session, SessionDeleter> session_;
  uint padded_frame_w_ = 0, padded_frame_h_ = 0, mv_field_w_ = 0, mv_field_h_ = 0;
  uint64_t frame_stream_bytes_ = 0, frame_in_bytes_ = 0;
  uint64_t frames_encoded_ = 0;
  // pinned staging owned by the encoder (svc_host_alloc)
  Pinned<uchar> h_in_;
  // two sets of output staging: a post thread labels, patches and pushes batch k out of one set
  // while the GPU encodes batch k+1 into the other
  Pinned<uchar> h_stream_[2];
  Pinned<float> h_mv_[2];
  Pinned<float> h_mad_[2];
};

}  // namespace svc

#endif  // SVC_B200_HOST_ENCODER_HPP
