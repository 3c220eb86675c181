// apps/encoder.cpp -- the encoder application on the GPU hot path: counterpart of the
// reference's apps/encoder.cpp (three threads: reader -> Encoder -> writer, bounded
// queues of capacity 10, :172-173, :223-228; byte stream on stdout, :154-169).
//
// The reference decodes a video file with cv::VideoCapture (:192-204); this image has
// no C++ OpenCV, so the frame source is raw interleaved 8-bit BGR (file or stdin), e.g.
//   ffmpeg -i in.mp4 -f rawvideo -pix_fmt bgr24 - | svc_encoder --width W --height H - > out.svc
// Options keep the reference's names (apps/encoder.cpp:75-104) for the hot-path fields.
// Options keep the reference's names for the block-type stages too (RANSAC, morphology,
// k-means, connected components: host/segment.hpp); --segment 0 leaves every block
// BLOCK_TYPE_BACKGROUND, --seed makes the labels reproducible.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include <vector>

#include "../host/encoder.hpp"
#include "../host/sharded.hpp"

static void usage() {
  std::fprintf(stderr,
               "usage: svc_encoder --width W --height H [--frames N] [--mv-search-range R]\n"
               "                   [--pyr-lvl-count L] [--mv-block-w B] [--mv-block-h B]\n"
               "                   [--transform-block-w T] [--transform-block-h T] [--device D]\n"
               "                   [--ransac-subset-sz N] [--ransac-inlier-thresh T] [--ransac-success-prob P]\n"
               "                   [--ransac-inlier-ratio W] [--morph-rect-w M] [--morph-rect-h M]\n"
               "                   [--kmeans-cluster-count K] [--kmeans-attempt-count A] [--kmeans-max-iter-count I]\n"
               "                   [--kmeans-epsilon E] [--connected-components-connectivity 4|8]\n"
               "                   [--segment 0|1] [--seed S] [--classify-threads T]\n"
               "                   [--batch K] [--verbose 0|1] <raw-bgr-file | ->\n"
               "       svc_encoder ... --devices 0,1,2,3 --out stream.svc <raw-bgr-file>\n"
               "         (frame-range sharded over several GPUs, one host thread per GPU)\n");
}

int main(int argc, char** argv) {
  svc::EncoderConfig cfg;  // defaults of apps/encoder.cpp:42-58
  unsigned width = 0, height = 0, frames = 0;
  int verbose = 1;
  const char* path = nullptr;
  const char* out_path = nullptr;
  std::vector<int> devices;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&](unsigned& out) {
      if (i + 1 >= argc) { usage(); std::exit(EXIT_FAILURE); }
      out = (unsigned)std::strtoul(argv[++i], nullptr, 10);
    };
    auto fval = [&](float& out) {
      if (i + 1 >= argc) { usage(); std::exit(EXIT_FAILURE); }
      out = std::strtof(argv[++i], nullptr);
    };
    unsigned tmp;
    if (a == "--width") val(width);
    else if (a == "--height") val(height);
    else if (a == "--frames") val(frames);
    else if (a == "--mv-search-range") val(cfg.mv_search_range);
    else if (a == "--pyr-lvl-count") val(cfg.pyr_lvl_count);
    else if (a == "--mv-block-w") val(cfg.mv_block_w);
    else if (a == "--mv-block-h") val(cfg.mv_block_h);
    else if (a == "--transform-block-w") val(cfg.transform_block_w);
    else if (a == "--transform-block-h") val(cfg.transform_block_h);
    else if (a == "--ransac-subset-sz") val(cfg.seg.ransac.subset_sz);
    else if (a == "--ransac-inlier-thresh") fval(cfg.seg.ransac.inlier_thresh);
    else if (a == "--ransac-success-prob") fval(cfg.seg.ransac.success_prob);
    else if (a == "--ransac-inlier-ratio") fval(cfg.seg.ransac.inlier_ratio);
    else if (a == "--morph-rect-w") val(cfg.seg.morph_rect_w);
    else if (a == "--morph-rect-h") val(cfg.seg.morph_rect_h);
    else if (a == "--kmeans-cluster-count") val(cfg.seg.kmeans.cluster_count);
    else if (a == "--kmeans-attempt-count") val(cfg.seg.kmeans.attempt_count);
    else if (a == "--kmeans-max-iter-count") val(cfg.seg.kmeans.max_iter_count);
    else if (a == "--kmeans-epsilon") fval(cfg.seg.kmeans.epsilon);
    else if (a == "--connected-components-connectivity") val(cfg.seg.connected_components_connectivity);
    else if (a == "--segment") { val(tmp); cfg.segment = tmp != 0; }
    else if (a == "--seed") { if (i + 1 >= argc) { usage(); return EXIT_FAILURE; } cfg.seed = std::strtoull(argv[++i], nullptr, 10); }
    else if (a == "--classify-threads") val(cfg.classify_threads);
    else if (a == "--batch") val(cfg.max_batch);
    else if (a == "--device") { val(tmp); cfg.device = (int)tmp; }
    else if (a == "--verbose") { val(tmp); verbose = (int)tmp; }
    else if (a == "--out") { if (i + 1 >= argc) { usage(); return EXIT_FAILURE; } out_path = argv[++i]; }
    else if (a == "--devices") {
      if (i + 1 >= argc) { usage(); return EXIT_FAILURE; }
      for (const char* q = argv[++i]; *q;) {
        devices.push_back((int)std::strtol(q, const_cast<char**>(&q), 10));
        if (*q == ',') ++q;
      }
    }
    else if (a == "-" || a[0] != '-') path = argv[i];
    else { usage(); return EXIT_FAILURE; }
  }
  if (!path || !width || !height) { usage(); return EXIT_FAILURE; }
  const svc::Status st = svc::Validate(cfg);
  if (st.code != svc::ErrorCode::kOk) {  // apps/encoder.cpp:185-190
    std::fprintf(stderr, "Invalid encoder configuration: %s\n", st.message.c_str());
    return EXIT_FAILURE;
  }
  FILE* in = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
  if (!in) { std::fprintf(stderr, "Failed to open %s\n", path); return EXIT_FAILURE; }
  const size_t fbytes = (size_t)width * height * 3;
  if (!frames && in != stdin) {  // frame count from the file size (the header needs it up front)
    std::fseek(in, 0, SEEK_END);
    frames = (unsigned)(std::ftell(in) / (long)fbytes);
    std::fseek(in, 0, SEEK_SET);
  }
  if (!frames) { std::fprintf(stderr, "--frames is required when reading stdin\n"); return EXIT_FAILURE; }
  if (verbose) std::fprintf(stderr, "frame width: %u\nframe height: %u\nframe count: %u\n", width, height, frames);

  if (!devices.empty()) {  // sharded multi-GPU mode: seekable input, file output
    if (in == stdin || !out_path) {
      std::fprintf(stderr, "--devices needs a seekable input file and --out\n");
      return EXIT_FAILURE;
    }
    std::fclose(in);
    try {
      const svc::ShardedStats s = svc::EncodeFileSharded(cfg, svc::VideoProperties{width, height, frames}, path,
                                                         out_path, devices);
      if (verbose)
        std::fprintf(stderr,
                     "encoded %llu frames on %zu device(s) in %.3f s (%.1f frames/s overall)\n"
                     "  slowest shard: set-up %.3f s, then %.1f frames/s; busy time: read %.3f s, GPU encode + "
                     "block types %.3f s, write %.3f s (block types and writes overlap the next batch)\n",
                     (unsigned long long)s.frames_encoded, devices.size(), s.seconds,
                     s.seconds > 0 ? s.frames_encoded / s.seconds : 0.0, s.setup_seconds,
                     s.seconds > s.setup_seconds ? s.frames_encoded / (s.seconds - s.setup_seconds) : 0.0,
                     s.read_seconds, s.encode_seconds, s.write_seconds);
    } catch (const std::exception& e) {
      std::fprintf(stderr, "svc_encoder: %s\n", e.what());
      return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
  }

  svc::BoundedQueue<svc::Frame> in_queue(10);
  svc::BoundedQueue<svc::Bytes> out_queue(10);
  int rc = EXIT_SUCCESS;
  try {
    svc::Encoder encoder(cfg, svc::VideoProperties{width, height, frames}, in_queue, out_queue);
    std::thread reader([&] {
      svc::Frame f(fbytes);
      for (unsigned i = 0; i < frames; ++i) {
        if (std::fread(f.data(), 1, fbytes, in) != fbytes) break;
        if (!in_queue.Push(f)) break;  // closed: the encoder failed
      }
      in_queue.SignalProducerIsDone();
    });
    std::thread writer([&] {
      svc::Bytes b;
      while (out_queue.Pop(b)) {
        if (std::fwrite(b.data(), 1, b.size(), stdout) != b.size()) {
          std::fprintf(stderr, "Failed to write bytes.\n");  // apps/encoder.cpp:163-167
          rc = EXIT_FAILURE;
        }
      }
      std::fflush(stdout);
    });
    try {
      encoder();
    } catch (const std::exception& e) {
      std::fprintf(stderr, "svc_encoder: %s\n", e.what());
      rc = EXIT_FAILURE;
      in_queue.Close();  // (the encoder closed both already; harmless to repeat)
      out_queue.Close();
    }
    reader.join();
    writer.join();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "svc_encoder: %s\n", e.what());
    rc = EXIT_FAILURE;
  }
  if (in != stdin) std::fclose(in);
  return rc;
}
