// apps/decoder.cpp -- stream decoder on the GPU block path: counterpart of the reference's
// apps/decoder.cpp + Decoder::operator() (libs/decoder.cpp:151-218) without the GUI.  Reads the
// codec stream (32-byte Header, then 772-byte records) from a file or stdin, dequantises and
// inverse-transforms every frame on the device (svc_decode_frames_device: ParseBlock + DecodeBlock,
// libs/decoder.cpp:102-149) and writes raw 8-bit BGR frames:
//   svc_encoder --width W --height H in.bgr | svc_decoder - > out.bgr
// Options keep the reference's names (apps/decoder.cpp:34-40).  What differs:
//   * no window: the gaze position is an option (--gaze-x/--gaze-y, in original-frame pixels; none by
//     default) instead of the mouse, and frames are written instead of shown.  The reference also
//     rescales the padded frame to the original size for display (cv::resize, libs/decoder.cpp:209);
//     here the padded frame is cropped to W x H (--padded writes it whole);
//   * the reference reader expects padded_h / tbh block rows per frame while its encoder writes
//     ceil(h / tbh) (136 vs 135 at 1080p, the streams desynchronise there).  This decoder
//     follows what the encoder wrote and leaves the missing block rows black.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/svc_b200.h"

static void usage() {
  std::fprintf(stderr,
               "usage: svc_decoder [--foreground-quant-step Q] [--background-quant-step Q]\n"
               "                   [--max-gaze-rect-w W] [--max-gaze-rect-h H] [--gaze-x X --gaze-y Y]\n"
               "                   [--device D] [--batch K] [--padded 0|1] [--verbose 0|1] <stream-file | ->\n");
}

static bool read_all(FILE* f, void* dst, size_t n) { return std::fread(dst, 1, n, f) == n; }

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != SVC_OK) {                                                \
      std::fprintf(stderr, "svc_decoder: %s\n", svc_last_error());         \
      return EXIT_FAILURE;                                                 \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  unsigned fg_q = 1, bg_q = 640, gaze_w = 64, gaze_h = 64;  // apps/decoder.cpp:20-25
  int gaze_x = -1, gaze_y = -1, device = 0, verbose = 1, padded = 0;
  unsigned batch = 8;
  const char* path = nullptr;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* {
      if (i + 1 >= argc) { usage(); std::exit(EXIT_FAILURE); }
      return argv[++i];
    };
    if (a == "--foreground-quant-step") fg_q = (unsigned)std::strtoul(next(), nullptr, 10);
    else if (a == "--background-quant-step") bg_q = (unsigned)std::strtoul(next(), nullptr, 10);
    else if (a == "--max-gaze-rect-w") gaze_w = (unsigned)std::strtoul(next(), nullptr, 10);
    else if (a == "--max-gaze-rect-h") gaze_h = (unsigned)std::strtoul(next(), nullptr, 10);
    else if (a == "--gaze-x") gaze_x = std::atoi(next());
    else if (a == "--gaze-y") gaze_y = std::atoi(next());
    else if (a == "--device") device = std::atoi(next());
    else if (a == "--batch") batch = std::max(1u, (unsigned)std::strtoul(next(), nullptr, 10));
    else if (a == "--padded") padded = std::atoi(next());
    else if (a == "--verbose") verbose = std::atoi(next());
    else if (a == "-" || a[0] != '-') path = argv[i];
    else { usage(); return EXIT_FAILURE; }
  }
  if (!path) { usage(); return EXIT_FAILURE; }
  if (fg_q == 0) {  // Validate(DecoderConfig), libs/decoder.cpp:35-47
    std::fprintf(stderr, "Invalid decoder configuration: invalid foreground quantization step: must be > 0\n");
    return EXIT_FAILURE;
  }
  if (bg_q == 0) {
    std::fprintf(stderr, "Invalid decoder configuration: invalid background quantization step: must be > 0\n");
    return EXIT_FAILURE;
  }
  FILE* in = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
  if (!in) { std::fprintf(stderr, "Failed to open %s\n", path); return EXIT_FAILURE; }
  uint8_t hdr[32];
  if (!read_all(in, hdr, 32)) { std::fprintf(stderr, "failed to read the stream header\n"); return EXIT_FAILURE; }
  svc_stream_layout lay{};
  CHECK(svc_stream_layout_from_header(hdr, &lay));
  if (lay.tbw != lay.tbh || (lay.tbw != 8 && lay.tbw != 16 && lay.tbw != 4) || lay.channels != 3) {
    std::fprintf(stderr, "svc_decoder: only 8x8, 16x16 and 4x4 transform blocks of 3 channels are supported\n");
    return EXIT_FAILURE;
  }
  const uint32_t pw = lay.padded_w, ph = lay.padded_h, w = lay.frame_w, h = lay.frame_h;
  if ((w + lay.tbw - 1) / lay.tbw != pw / lay.tbw) {
    // with horizontal padding the encoder's rows hold fewer records than the decoder's, and its
    // serializer read the planes with the unpadded stride: such a stream has no decoding
    std::fprintf(stderr, "svc_decoder: horizontally padded stream (%u -> %u): not decodable\n", w, pw);
    return EXIT_FAILURE;
  }
  const size_t rec_enc = (size_t)lay.encoder_records_per_frame * lay.record_bytes;
  const size_t rec_dec = (size_t)lay.decoder_records_per_frame * lay.record_bytes;
  if (verbose) {
    std::fprintf(stderr, "frame width: %u\nframe height: %u\nframe count: %u\n", w, h, lay.frame_count);
    if (!lay.consistent)
      std::fprintf(stderr,
                   "note: the encoder wrote %llu records per frame, the reference decoder would read %llu; the "
                   "missing block rows stay black\n",
                   (unsigned long long)lay.encoder_records_per_frame,
                   (unsigned long long)lay.decoder_records_per_frame);
  }
  svc_rect gaze{};
  const bool has_gaze = gaze_x >= 0 && gaze_y >= 0;
  if (has_gaze) CHECK(svc_gaze_rect((uint32_t)gaze_x, (uint32_t)gaze_y, gaze_w, gaze_h, w, h, pw, ph, &gaze));

  const size_t out_px = (size_t)pw * ph * 3;
  uint8_t* h_rec = static_cast<uint8_t*>(svc_host_alloc(batch * rec_dec));
  float* h_out = static_cast<float*>(svc_host_alloc(batch * out_px * sizeof(float)));
  uint8_t* d_rec = static_cast<uint8_t*>(svc_device_alloc(device, batch * rec_dec));
  float* d_out = static_cast<float*>(svc_device_alloc(device, batch * out_px * sizeof(float)));
  if (!h_rec || !h_out || !d_rec || !d_out) {
    std::fprintf(stderr, "svc_decoder: allocation failed (no CUDA device? there is no CPU fallback)\n");
    return EXIT_FAILURE;
  }
  const uint32_t ow = padded ? pw : w, oh = padded ? ph : h;
  std::vector<uint8_t> frame8((size_t)ow * oh * 3);
  int rc = EXIT_SUCCESS;
  for (uint32_t done = 0; done < lay.frame_count && rc == EXIT_SUCCESS;) {
    const uint32_t n = std::min<uint32_t>(batch, lay.frame_count - done);
    std::memset(h_rec, 0, (size_t)n * rec_dec);  // block rows the encoder did not write: type 0, zero coefficients
    uint32_t got = 0;
    for (; got < n; ++got)
      if (!read_all(in, h_rec + (size_t)got * rec_dec, std::min(rec_enc, rec_dec))) break;
    if (got < n) {
      std::fprintf(stderr, "failed to read all expected blocks\n");  // apps/decoder.cpp:74-77
      rc = EXIT_FAILURE;
    }
    if (got == 0) break;
    CHECK(svc_memcpy_h2d(device, d_rec, h_rec, (size_t)got * rec_dec));
    CHECK(svc_decode_frames_device(device, nullptr, d_rec, got, pw, ph, lay.tbw, lay.tbh, fg_q, bg_q,
                                   has_gaze ? &gaze : nullptr, d_out));
    CHECK(svc_memcpy_d2h(device, h_out, d_out, (size_t)got * out_px * sizeof(float)));
    for (uint32_t f = 0; f < got; ++f) {
      const float* src = h_out + (size_t)f * out_px;
      for (uint32_t y = 0; y < oh; ++y) {
        const float* srow = src + (size_t)y * pw * 3;
        uint8_t* drow = frame8.data() + (size_t)y * ow * 3;
        for (uint32_t x = 0; x < ow * 3; ++x) {  // cv::Mat3f -> 8 bit: saturate_cast<uchar>(round(v))
          const float v = std::nearbyint(srow[x]);
          drow[x] = (uint8_t)(v < 0.f ? 0.f : (v > 255.f ? 255.f : v));
        }
      }
      if (std::fwrite(frame8.data(), 1, frame8.size(), stdout) != frame8.size()) {
        std::fprintf(stderr, "Failed to write bytes.\n");
        rc = EXIT_FAILURE;
        break;
      }
    }
    done += got;
  }
  std::fflush(stdout);
  svc_host_free(h_rec);
  svc_host_free(h_out);
  svc_device_free(device, d_rec);
  svc_device_free(device, d_out);
  if (in != stdin) std::fclose(in);
  return rc;
}
