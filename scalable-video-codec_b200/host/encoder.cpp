// encoder.cpp -- see encoder.hpp.  Per-batch flow (one GPU):
//   pop up to max_batch frames -> svc_session_encode (H2D | K3+K1 | K2 | D2H)
//   -> per encoded frame: block-type callback on the motion field, patch the
//   4-byte block types into the GPU-written records, push the bytes.
#include "encoder.hpp"

#include <cstring>
#include <exception>
#include <thread>

#include "../../include/svc_b200.h"

namespace svc {

Status Validate(const EncoderConfig& cfg) {
  // messages as in libs/encoder.cpp:62-142
  if (cfg.mv_block_w < 1) return {ErrorCode::kInvalidParameter, "invalid mv block width: must be > 0"};
  if (cfg.mv_block_h < 1) return {ErrorCode::kInvalidParameter, "invalid mv block height: must be > 0"};
  if (cfg.pyr_lvl_count < 1)
    return {ErrorCode::kInvalidParameter, "invalid pyramid level count: must be > 0"};
  if (cfg.pyr_lvl_count > SVC_MAX_LEVELS)
    return {ErrorCode::kInvalidParameter, "invalid pyramid level count: must be <= 8"};
  if (cfg.mv_search_range / (1u << (cfg.pyr_lvl_count - 1)) == 0)
    return {ErrorCode::kInvalidParameter,
            "invalid mv search and pyramid level count: the quotient from dividing the mv search "
            "range by the pyramid level reduction factor must be > 0"};
  if (cfg.segment) {  // Validate(ransac), Validate(kmeans), connectivity: libs/encoder.cpp:86-101
    SegmentConfig sc = cfg.seg;
    const char* m = ValidateSegmentConfig(sc);
    if (*m) return {ErrorCode::kInvalidParameter, m};
  }
  if (cfg.transform_block_w < 1)
    return {ErrorCode::kInvalidParameter, "invalid transform block width: must be > 0"};
  if (cfg.transform_block_h < 1)
    return {ErrorCode::kInvalidParameter, "invalid transform block height: must be > 0"};
  if (cfg.transform_block_w > cfg.mv_block_w)
    return {ErrorCode::kInvalidParameter,
            "invalid transform block width and mv block width: transform block width must be <= mv "
            "block width"};
  if (cfg.transform_block_h > cfg.mv_block_h)
    return {ErrorCode::kInvalidParameter,
            "invalid transform block height and mv block height: transform block height must be <= "
            "mv block height"};
  if (cfg.mv_block_w % cfg.transform_block_w != 0)
    return {ErrorCode::kInvalidParameter,
            "invalid mv block width and transform block width: mv block width must be divisible by "
            "transform block width"};
  if (cfg.mv_block_h % cfg.transform_block_h != 0)
    return {ErrorCode::kInvalidParameter,
            "invalid mv block height and transform block height: mv block height must be divisible "
            "by transform block height"};
  return {ErrorCode::kOk, ""};
}

static void check(int rc) {
  if (rc != SVC_OK) throw Error(rc, svc_last_error());
}

void Encoder::SessionDeleter::operator()(svc_session* s) const { svc_session_destroy(s); }
void Encoder::PinnedDeleter::operator()(void* p) const { svc_host_free(p); }

Encoder::Encoder(const EncoderConfig& cfg, const VideoProperties& vidprops,
                 BoundedQueue<Frame>& in_queue, BoundedQueue<Bytes>& out_queue, BlockTypeFn classify)
    : cfg_(cfg), vidprops_(vidprops), in_queue_(in_queue), out_queue_(out_queue),
      classify_(std::move(classify)) {
  svc_session_config c{};
  c.struct_size = sizeof(c);
  c.frame_w = vidprops.frame_w;
  c.frame_h = vidprops.frame_h;
  c.mv_block_w = cfg.mv_block_w;
  c.mv_block_h = cfg.mv_block_h;
  c.mv_search_range = cfg.mv_search_range;
  c.pyr_lvl_count = cfg.pyr_lvl_count;
  c.transform_block_w = cfg.transform_block_w;
  c.transform_block_h = cfg.transform_block_h;
  c.device = cfg.device;
  c.max_batch = cfg.max_batch;
  {
    svc_session* raw = nullptr;
    check(svc_session_create(&c, &raw));
    session_.reset(raw);
  }
  svc_session_info info{};
  check(svc_session_info_get(session_.get(), &info));
  padded_frame_w_ = info.padded_w;  // libs/encoder.cpp:165-175
  padded_frame_h_ = info.padded_h;
  mv_field_w_ = info.mv_field_w;
  mv_field_h_ = info.mv_field_h;
  frame_stream_bytes_ = info.frame_stream_bytes;
  frame_in_bytes_ = info.frame_in_bytes;
  const size_t B = info.max_batch, mvn = (size_t)mv_field_w_ * mv_field_h_;
  h_in_.reset(static_cast<uchar*>(svc_host_alloc(B * frame_in_bytes_)));
  if (!h_in_) throw Error(SVC_ERR_CUDA, "pinned allocation failed");
  for (int k = 0; k < 2; ++k) {
    h_stream_[k].reset(static_cast<uchar*>(svc_host_alloc(B * frame_stream_bytes_)));
    h_mv_[k].reset(static_cast<float*>(svc_host_alloc(B * mvn * 2 * sizeof(float))));
    h_mad_[k].reset(static_cast<float*>(svc_host_alloc(B * mvn * sizeof(float))));
    if (!h_stream_[k] || !h_mv_[k] || !h_mad_[k]) throw Error(SVC_ERR_CUDA, "pinned allocation failed");
  }
  if (!classify_ && cfg_.segment) {
    SegmentConfig sc = cfg_.seg;
    sc.mv_block_w = cfg_.mv_block_w;
    sc.mv_block_h = cfg_.mv_block_h;
    stage_.reset(new BlockTypeStage(sc, mv_field_w_, mv_field_h_, cfg_.seed, cfg_.classify_threads));
  }
}

Encoder::~Encoder() = default;  // pinned buffers, then the session (members in reverse order)

void Encoder::operator()() {
  // Whatever happens below, the consumer of the output queue is released: a normal return
  // signals completion (libs/encoder.cpp:666); an exception closes BOTH queues, so a reader
  // blocked in Push on a full input queue and a writer blocked in Pop both wake up.
  struct Finish {
    BoundedQueue<Frame>& in;
    BoundedQueue<Bytes>& out;
    bool ok = false;
    ~Finish() {
      if (ok) {
        out.SignalProducerIsDone();
      } else {
        in.Close();
        out.Close();
      }
    }
  } finish{in_queue_, out_queue_};
  // The reference pops the first frame before it emits anything and returns without output
  // when the input is empty (libs/encoder.cpp:344-348).
  Frame frame;
  if (!in_queue_.Pop(frame)) {
    finish.ok = true;
    return;
  }
  bool have_first = true;
  // Header next (libs/encoder.cpp:361-381): frame_count excludes the tracked-only first frame
  {
    Bytes hdr(32);
    check(svc_write_header(vidprops_.frame_count, vidprops_.frame_w, vidprops_.frame_h, padded_frame_w_,
                           padded_frame_h_, cfg_.transform_block_w, cfg_.transform_block_h, 3,
                           hdr.data()));
    out_queue_.Push(std::move(hdr));
  }
  const size_t mvn = (size_t)mv_field_w_ * mv_field_h_;
  std::vector<uint> block_types(mvn);
  std::vector<uint> batch_types;
  svc_session_info info{};
  check(svc_session_info_get(session_.get(), &info));
  std::thread post;  // at most one alive: batch k's labels / patching / pushes, in frame order
  std::exception_ptr post_error;
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } joiner{post};
  for (uint k = 0;; ++k) {
    // block for one frame, then take whatever else is already queued (<= max_batch)
    if (!have_first && !in_queue_.Pop(frame)) break;
    have_first = false;
    uint n = 0;
    do {
      if (frame.size() != frame_in_bytes_) throw Error(SVC_ERR_INVALID_ARG, "frame has the wrong size");
      std::memcpy(h_in_.get() + (size_t)n * frame_in_bytes_, frame.data(), frame_in_bytes_);
      ++n;
    } while (n < info.max_batch && in_queue_.TryPop(frame));
    const int set = (int)(k & 1);
    uchar* h_stream = h_stream_[set].get();
    float* h_mv = h_mv_[set].get();
    float* h_mad = h_mad_[set].get();
    uint n_enc = 0;
    check(svc_session_encode(session_.get(), h_in_.get(), n, h_mv, h_mad, h_stream, nullptr, &n_enc));
    // batch k-1 is out of the other set before batch k+1 may overwrite it, and frames leave in order
    if (post.joinable()) post.join();
    if (post_error) std::rethrow_exception(post_error);
    const uint64_t first_frame = frames_encoded_;
    frames_encoded_ += n_enc;
    post = std::thread([&, n_enc, first_frame, h_stream, h_mv, h_mad] {
      try {
        if (stage_ && n_enc) {  // libs/encoder.cpp:491-624 for the whole batch, on worker threads
          batch_types.resize((size_t)n_enc * mvn);
          stage_->Run(reinterpret_cast<const Vec2f*>(h_mv), n_enc, first_frame, batch_types.data());
        }
        for (uint i = 0; i < n_enc; ++i) {
          uchar* rec = h_stream + (size_t)i * frame_stream_bytes_;
          if (stage_) {
            check(svc_patch_block_types(rec, vidprops_.frame_w, vidprops_.frame_h, cfg_.transform_block_w,
                                        cfg_.transform_block_h, 3, cfg_.mv_block_w, cfg_.mv_block_h,
                                        mv_field_w_, batch_types.data() + (size_t)i * mvn));
          } else if (classify_) {
            std::fill(block_types.begin(), block_types.end(), 0u);  // BLOCK_TYPE_BACKGROUND, libs/encoder.cpp:549-551
            classify_(reinterpret_cast<const Vec2f*>(h_mv + (size_t)i * mvn * 2), h_mad + (size_t)i * mvn,
                      mv_field_w_, mv_field_h_, block_types.data());
            check(svc_patch_block_types(rec, vidprops_.frame_w, vidprops_.frame_h, cfg_.transform_block_w,
                                        cfg_.transform_block_h, 3, cfg_.mv_block_w, cfg_.mv_block_h,
                                        mv_field_w_, block_types.data()));
          }
          if (!out_queue_.Push(Bytes(rec, rec + frame_stream_bytes_)))
            throw Error(SVC_ERR_INVALID_ARG, "output queue closed by its consumer");
        }
      } catch (...) {
        post_error = std::current_exception();
      }
    });
  }
  if (post.joinable()) post.join();
  if (post_error) std::rethrow_exception(post_error);
  finish.ok = true;  // -> SignalProducerIsDone, libs/encoder.cpp:666
}

}  // namespace svc
