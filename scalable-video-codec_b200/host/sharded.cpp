// sharded.cpp -- see sharded.hpp.
#include "sharded.hpp"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>

#include "../../include/svc_b200.h"

namespace svc {

std::vector<ShardRange> ShardFrameRanges(uint n_input_frames, uint world) {
  std::vector<ShardRange> out;
  const uint n_enc = n_input_frames ? n_input_frames - 1 : 0;
  const uint base = world ? n_enc / world : 0, rem = world ? n_enc % world : 0;
  uint t = 1;
  for (uint r = 0; r < world; ++r) {
    const uint k = base + (r < rem ? 1 : 0);
    if (k == 0) out.push_back({t - 1, t - 1, t, t});
    else out.push_back({t - 1, t + k, t, t + k});
    t += k;
  }
  return out;
}

namespace {

struct Pinned {
  void* p = nullptr;
  explicit Pinned(size_t n) : p(svc_host_alloc(n)) {
    if (!p) throw Error(SVC_ERR_CUDA, "pinned allocation failed");
  }
  ~Pinned() { svc_host_free(p); }
};

void check(int rc) {
  if (rc != SVC_OK) throw Error(rc, svc_last_error());
}

void pwrite_all(int fd, const uchar* p, size_t n, uint64_t off) {
  while (n) {
    const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
    if (w <= 0) throw Error(SVC_ERR_STATE, "Failed to write bytes.");
    p += w;
    n -= (size_t)w;
    off += (uint64_t)w;
  }
}

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// times[]: set-up, read, encode (+ block types), write
void run_shard(const EncoderConfig& cfg, const VideoProperties& vp, const std::string& in_path, int out_fd,
               int device, ShardRange r, const BlockTypeFn& classify, uint64_t* encoded,
               std::array<double, 4>* times) {
  if (r.enc_hi <= r.enc_lo) return;
  double t0 = now();
  svc_session_config c{};
  c.struct_size = sizeof(c);
  c.frame_w = vp.frame_w; c.frame_h = vp.frame_h;
  c.mv_block_w = cfg.mv_block_w; c.mv_block_h = cfg.mv_block_h;
  c.mv_search_range = cfg.mv_search_range; c.pyr_lvl_count = cfg.pyr_lvl_count;
  c.transform_block_w = cfg.transform_block_w; c.transform_block_h = cfg.transform_block_h;
  c.device = device; c.max_batch = cfg.max_batch;
  svc_session* s = nullptr;
  check(svc_session_create(&c, &s));
  struct Guard { svc_session* s; ~Guard() { svc_session_destroy(s); } } guard{s};
  svc_session_info info{};
  check(svc_session_info_get(s, &info));
  const size_t B = info.max_batch, mvn = (size_t)info.mv_field_w * info.mv_field_h;
  // Two sets of output staging: while the GPU encodes batch k+1 into one set, a post thread labels
  // the motion fields of batch k (block-type stages, libs/encoder.cpp:491-624), patches the labels
  // into its records and writes them out of the other set.
  Pinned h_in(B * info.frame_in_bytes);
  struct OutSet {
    Pinned st, mv, mad;
    std::thread post;
    std::exception_ptr error;
    OutSet(size_t b, size_t fst, size_t mvn) : st(b * fst), mv(b * mvn * 8), mad(b * mvn * 4) {}
  };
  OutSet sets[2] = {OutSet(B, info.frame_stream_bytes, mvn), OutSet(B, info.frame_stream_bytes, mvn)};
  std::mutex post_mu;  // post sections run one at a time, in batch order (the stage's scratch is shared)
  std::vector<uint> bt(mvn), batch_types;
  std::unique_ptr<BlockTypeStage> stage;
  if (!classify && cfg.segment) {
    SegmentConfig sc = cfg.seg;
    sc.mv_block_w = cfg.mv_block_w;
    sc.mv_block_h = cfg.mv_block_h;
    stage.reset(new BlockTypeStage(sc, info.mv_field_w, info.mv_field_h, cfg.seed, cfg.classify_threads));
  }
  FILE* in = std::fopen(in_path.c_str(), "rb");
  if (!in) throw Error(SVC_ERR_INVALID_ARG, "Failed to open " + in_path);
  struct FGuard { FILE* f; ~FGuard() { std::fclose(f); } } fguard{in};
  if (fseeko(in, (off_t)((uint64_t)r.in_lo * info.frame_in_bytes), SEEK_SET) != 0)
    throw Error(SVC_ERR_INVALID_ARG, "seek failed");
  uint next_in = r.in_lo, next_enc = r.enc_lo;
  (*times)[0] = now() - t0;
  std::atomic<uint64_t> post_ns{0}, write_ns{0};
  struct Joiner {  // declared last: on an exception the post threads are joined before anything they use dies
    OutSet* s;
    ~Joiner() {
      for (int i = 0; i < 2; ++i)
        if (s[i].post.joinable()) s[i].post.join();
    }
  } joiner{sets};
  for (uint k = 0; next_in < r.in_hi; ++k) {
    OutSet& o = sets[k & 1];
    if (o.post.joinable()) o.post.join();  // this set's previous batch is on disk
    if (o.error) std::rethrow_exception(o.error);
    const uint n = (uint)std::min<size_t>(B, r.in_hi - next_in);
    t0 = now();
    if (std::fread(h_in.p, info.frame_in_bytes, n, in) != n) throw Error(SVC_ERR_INVALID_ARG, "short read");
    (*times)[1] += now() - t0;
    t0 = now();
    uint n_enc = 0;
    check(svc_session_encode(s, static_cast<const uint8_t*>(h_in.p), n, static_cast<float*>(o.mv.p),
                             static_cast<float*>(o.mad.p), static_cast<uint8_t*>(o.st.p), nullptr, &n_enc));
    (*times)[2] += now() - t0;
    const uint first_enc = next_enc;
    o.post = std::thread([&, n_enc, first_enc, po = &o] {
      try {
        std::lock_guard<std::mutex> lock(post_mu);
        const double p0 = now();
        if (stage && n_enc) {  // per-frame generators depend on the global frame index only
          batch_types.resize((size_t)n_enc * mvn);
          stage->Run(reinterpret_cast<const Vec2f*>(po->mv.p), n_enc, first_enc - 1, batch_types.data());
        }
        for (uint i = 0; i < n_enc; ++i) {
          uchar* rec = static_cast<uchar*>(po->st.p) + (size_t)i * info.frame_stream_bytes;
          if (stage) {
            check(svc_patch_block_types(rec, vp.frame_w, vp.frame_h, cfg.transform_block_w, cfg.transform_block_h, 3,
                                        cfg.mv_block_w, cfg.mv_block_h, info.mv_field_w,
                                        batch_types.data() + (size_t)i * mvn));
          } else if (classify) {
            std::fill(bt.begin(), bt.end(), 0u);
            classify(reinterpret_cast<const Vec2f*>(static_cast<float*>(po->mv.p) + (size_t)i * mvn * 2),
                     static_cast<float*>(po->mad.p) + (size_t)i * mvn, info.mv_field_w, info.mv_field_h, bt.data());
            check(svc_patch_block_types(rec, vp.frame_w, vp.frame_h, cfg.transform_block_w, cfg.transform_block_h, 3,
                                        cfg.mv_block_w, cfg.mv_block_h, info.mv_field_w, bt.data()));
          }
        }
        const double p1 = now();
        // encoded frame t (anchor = input frame t) lives at 32 + (t-1) * frame_stream_bytes
        pwrite_all(out_fd, static_cast<uchar*>(po->st.p), (size_t)n_enc * info.frame_stream_bytes,
                   32 + (uint64_t)(first_enc - 1) * info.frame_stream_bytes);
        post_ns += (uint64_t)((p1 - p0) * 1e9);
        write_ns += (uint64_t)((now() - p1) * 1e9);
      } catch (...) {
        po->error = std::current_exception();
      }
    });
    next_in += n;
    next_enc += n_enc;
    *encoded += n_enc;
  }
  for (auto& o : sets) {
    if (o.post.joinable()) o.post.join();
    if (o.error) std::rethrow_exception(o.error);
  }
  (*times)[2] += post_ns.load() * 1e-9;   // block-type stages + patching (overlapped with the GPU batches)
  (*times)[3] += write_ns.load() * 1e-9;  // (overlapped too)
}

}  // namespace

ShardedStats EncodeFileSharded(const EncoderConfig& cfg, const VideoProperties& vidprops,
                               const std::string& in_path, const std::string& out_path,
                               const std::vector<int>& devices, BlockTypeFn classify) {
  if (devices.empty()) throw Error(SVC_ERR_INVALID_ARG, "no devices");
  const Status st = Validate(cfg);
  if (st.code != ErrorCode::kOk) throw Error(SVC_ERR_INVALID_ARG, st.message);
  const uint pw = svc_padded_dim(vidprops.frame_w, cfg.mv_block_w, cfg.pyr_lvl_count);
  const uint ph = svc_padded_dim(vidprops.frame_h, cfg.mv_block_h, cfg.pyr_lvl_count);
  const int fd = ::open(out_path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0) throw Error(SVC_ERR_INVALID_ARG, "Failed to open " + out_path);
  struct FdGuard { int fd; ~FdGuard() { ::close(fd); } } fdg{fd};
  uchar hdr[32];
  check(svc_write_header(vidprops.frame_count, vidprops.frame_w, vidprops.frame_h, pw, ph,
                         cfg.transform_block_w, cfg.transform_block_h, 3, hdr));
  pwrite_all(fd, hdr, 32, 0);
  const auto ranges = ShardFrameRanges(vidprops.frame_count, (uint)devices.size());
  std::vector<std::thread> threads;
  std::vector<std::exception_ptr> errors(devices.size());
  std::vector<uint64_t> encoded(devices.size(), 0);
  std::vector<std::array<double, 4>> times(devices.size(), std::array<double, 4>{0, 0, 0, 0});
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t g = 0; g < devices.size(); ++g)
    threads.emplace_back([&, g] {
      try {
        run_shard(cfg, vidprops, in_path, fd, devices[g], ranges[g], classify, &encoded[g], &times[g]);
      } catch (...) {
        errors[g] = std::current_exception();
      }
    });
  for (auto& t : threads) t.join();
  ShardedStats stats;
  stats.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto& e : errors)
    if (e) std::rethrow_exception(e);
  for (auto n : encoded) stats.frames_encoded += n;
  for (const auto& t : times) {
    stats.setup_seconds = std::max(stats.setup_seconds, t[0]);
    stats.read_seconds = std::max(stats.read_seconds, t[1]);
    stats.encode_seconds = std::max(stats.encode_seconds, t[2]);
    stats.write_seconds = std::max(stats.write_seconds, t[3]);
  }
  return stats;
}

}  // namespace svc
