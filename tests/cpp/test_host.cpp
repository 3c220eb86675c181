// test_host.cpp -- exercises the C++ host layer (host/motion.hpp, host/encoder.hpp)
// against the C oracle (oracle/svc_oracle.c; test infrastructure).
//   test_host            full run, needs a GPU
//   test_host --no-gpu   host-only checks (validation text, hard failure without a device)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../scalable-video-codec_b200/host/encoder.hpp"
#include "../../scalable-video-codec_b200/host/motion.hpp"
#include "../../scalable-video-codec_b200/host/sharded.hpp"
#include "../../include/svc_b200.h"

extern "C" {
unsigned orc_padded_dim(unsigned a, unsigned x, unsigned y);
void orc_y_pyramid(const uint8_t* bgr, unsigned w, unsigned h, unsigned pw, unsigned ph, unsigned levels,
                   uint8_t* const* out_levels);
void orc_hbma(const uint8_t* const* t, const uint8_t* const* a, unsigned levels, unsigned fw, unsigned fh,
              unsigned range, unsigned bw, unsigned bh, float* mv, float* mad);
void orc_ebma(const uint8_t* t, const uint8_t* a, unsigned fw, unsigned fh, unsigned r, unsigned bw,
              unsigned bh, float* mv, float* mad);
int orc_dct_planar(const uint8_t* bgr, unsigned w, unsigned h, unsigned pw, unsigned ph, unsigned tbw,
                   unsigned tbh, float* const* planes);
void orc_header(unsigned n, unsigned w, unsigned h, unsigned ew, unsigned eh, unsigned tbw, unsigned tbh,
                unsigned ch, uint8_t* out32);
uint64_t orc_serialized_frame_bytes(unsigned w, unsigned h, unsigned tbw, unsigned tbh, unsigned ch);
void orc_serialize_frame(const float* const* planes, uint64_t plane_elems, unsigned channels,
                         const uint32_t* bt, unsigned w, unsigned h, unsigned tbw, unsigned tbh,
                         unsigned mvw, unsigned mbw, unsigned mbh, uint8_t* out);
}

#define CHECK(c)                                                      \
  do {                                                                \
    if (!(c)) {                                                       \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);        \
      return 1;                                                       \
    }                                                                 \
  } while (0)

static std::vector<std::vector<uchar>> make_frames(uint w, uint h, uint n) {
  // smooth-ish texture translating by (2,-1) px per frame plus a little noise
  std::vector<std::vector<uchar>> f(n, std::vector<uchar>((size_t)w * h * 3));
  uint32_t s = 12345;
  auto rnd = [&] { s = s * 1664525u + 1013904223u; return s >> 24; };
  std::vector<uchar> base((size_t)(w + 64) * (h + 64) * 3);
  for (auto& b : base) b = (uchar)rnd();
  for (int pass = 0; pass < 2; ++pass)
    for (size_t i = 3; i + 3 < base.size(); ++i) base[i] = (uchar)((base[i - 3] + 2 * base[i] + base[i + 3]) / 4);
  for (uint k = 0; k < n; ++k)
    for (uint y = 0; y < h; ++y)
      for (uint x = 0; x < w; ++x)
        for (uint c = 0; c < 3; ++c) {
          const size_t si = ((size_t)(y + 32 - k) * (w + 64) + (x + 32 + 2 * k)) * 3 + c;
          f[k][((size_t)y * w + x) * 3 + c] = (uchar)std::min(255u, base[si] + (rnd() & 3u));
        }
  return f;
}

static void classify(const Vec2f* mv, const float*, uint mw, uint mh, uint* bt) {
  for (uint i = 0; i < mw * mh; ++i) bt[i] = (mv[i].x != 0.0f || mv[i].y != 0.0f) ? 1u + (i % 5u) : 0u;
}

int main(int argc, char** argv) {
  const bool no_gpu = argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0;
  // ---- validation mirrors the reference's Validate(EncoderConfig) -----------------
  {
    svc::EncoderConfig c;
    CHECK(svc::Validate(c).code == svc::ErrorCode::kOk);
    c.mv_search_range = 7;
    CHECK(svc::Validate(c).code == svc::ErrorCode::kInvalidParameter);
    CHECK(svc::Validate(c).message.find("quotient") != std::string::npos);
    c = svc::EncoderConfig();
    c.transform_block_w = 32;
    CHECK(svc::Validate(c).message.find("transform block width must be <= mv block width") != std::string::npos);
  }
  // ---- frame-range sharding: partition of the encoded frames, one overlap frame -------------
  for (uint n : {0u, 1u, 2u, 5u, 300u, 601u})
    for (uint world : {1u, 2u, 3u, 8u}) {
      const auto r = svc::ShardFrameRanges(n, world);
      CHECK(r.size() == world);
      uint next = 1;
      for (const auto& s : r) {
        CHECK(s.enc_lo == next && s.enc_hi >= s.enc_lo);
        if (s.enc_hi > s.enc_lo) CHECK(s.in_lo == s.enc_lo - 1 && s.in_hi == s.enc_hi);
        next = s.enc_hi;
      }
      CHECK(next == (n ? n : 1));
    }
  // ---- BoundedQueue: Close() releases a producer blocked on a full queue and a consumer blocked on
  // an empty one (the error path of the encoder application) ------------------------------------
  {
    svc::BoundedQueue<int> q(2);
    CHECK(q.Push(1) && q.Push(2));
    bool push_result = true;
    std::thread producer([&] { push_result = q.Push(3); });  // blocks: the queue is full
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    q.Close();
    producer.join();
    CHECK(!push_result);
    int v = 0;
    CHECK(!q.Pop(v) && !q.Push(4) && q.closed());
    svc::BoundedQueue<int> e(2);
    bool pop_result = true;
    std::thread consumer([&] { int x; pop_result = e.Pop(x); });  // blocks: the queue is empty
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    e.Close();
    consumer.join();
    CHECK(!pop_result);
    svc::BoundedQueue<int> d(2);  // SignalProducerIsDone keeps the reference semantics: drain, then false
    CHECK(d.Push(7));
    d.SignalProducerIsDone();
    CHECK(d.Pop(v) && v == 7 && !d.Pop(v));
  }
  int ndev = 0;
  svc_device_count(&ndev);
  if (no_gpu || ndev == 0) {
    if (ndev == 0) {
      // no CUDA device: the host layer must fail loudly, never fall back
      std::vector<uchar> z(64 * 64, 0);
      const uchar* lv[4] = {z.data(), z.data(), z.data(), z.data()};
      std::vector<Vec2f> mv(16);
      std::vector<float> mad(16);
      bool threw = false;
      try {
        EstimateMotionHierarchical16x16Sse2(lv, lv, 64, 64, 8, mv.data(), mad.data());
      } catch (const svc::Error& e) {
        threw = e.code == SVC_ERR_CUDA;
      }
      CHECK(threw);
    }
    std::printf("PASS (host-only)\n");
    return 0;
  }

  const uint w = 208, h = 120, n = 7, L = 4, R = 8;
  const uint pw = orc_padded_dim(w, 16, 8), ph = orc_padded_dim(h, 16, 8);
  auto frames = make_frames(w, h, n);

  // ---- oracle pyramids ---------------------------------------------------------------
  std::vector<std::vector<std::vector<uchar>>> pyr(n, std::vector<std::vector<uchar>>(L));
  for (uint i = 0; i < n; ++i) {
    uint8_t* lv[4];
    for (uint l = 0; l < L; ++l) {
      pyr[i][l].resize((size_t)(pw >> l) * (ph >> l));
      lv[l] = pyr[i][l].data();
    }
    orc_y_pyramid(frames[i].data(), w, h, pw, ph, L, lv);
  }
  const uint mw = pw / 16, mh = ph / 16, mvn = mw * mh;

  // ---- motion.hpp drop-ins vs oracle ------------------------------------------------------
  {
    const uchar* t[4] = {pyr[0][0].data(), pyr[0][1].data(), pyr[0][2].data(), pyr[0][3].data()};
    const uchar* a[4] = {pyr[1][0].data(), pyr[1][1].data(), pyr[1][2].data(), pyr[1][3].data()};
    std::vector<Vec2f> mv(mvn), mv2(mvn);
    std::vector<float> mad(mvn), mad2(mvn), emv(2 * mvn), emad(mvn);
    EstimateMotionHierarchical16x16Sse2(t, a, pw, ph, R, mv.data(), mad.data());
    EstimateMotionHierarchical(t, a, L, pw, ph, R, 16, 16, mv2.data(), mad2.data());
    orc_hbma(t, a, L, pw, ph, R, 16, 16, emv.data(), emad.data());
    CHECK(std::memcmp(mv.data(), emv.data(), sizeof(float) * 2 * mvn) == 0);
    CHECK(std::memcmp(mv2.data(), emv.data(), sizeof(float) * 2 * mvn) == 0);
    CHECK(std::memcmp(mad.data(), emad.data(), sizeof(float) * mvn) == 0);
    CHECK(std::memcmp(mad2.data(), emad.data(), sizeof(float) * mvn) == 0);
    const uint ew = pw / 8, eh = ph / 8;
    std::vector<Vec2f> em(ew * eh);
    std::vector<float> emd(ew * eh), oem(2 * ew * eh), oemd(ew * eh);
    EstimateMotionExhaustiveSearch(t[0], a[0], pw, ph, 3, 8, 8, em.data(), emd.data());
    orc_ebma(t[0], a[0], pw, ph, 3, 8, 8, oem.data(), oemd.data());
    CHECK(std::memcmp(em.data(), oem.data(), sizeof(float) * 2 * ew * eh) == 0);
    CHECK(std::memcmp(emd.data(), oemd.data(), sizeof(float) * ew * eh) == 0);
    bool threw = false;
    try {
      EstimateMotionHierarchical(t, a, L, pw, ph, 4, 16, 16, mv.data(), mad.data());  // R < 2^(L-1)
    } catch (const svc::Error& e) {
      threw = e.code == SVC_ERR_INVALID_ARG;
    }
    CHECK(threw);
  }

  // ---- Encoder functor: reader thread -> encoder -> writer thread ------------------------------
  svc::EncoderConfig cfg;
  cfg.max_batch = 3;  // several uneven batches
  svc::VideoProperties vp{w, h, n};
  svc::BoundedQueue<svc::Frame> in_q(10);   // apps/encoder.cpp:172-173
  svc::BoundedQueue<svc::Bytes> out_q(10);
  std::vector<svc::Bytes> got;
  svc::Encoder enc(cfg, vp, in_q, out_q, classify);
  CHECK(enc.padded_frame_w() == pw && enc.padded_frame_h() == ph);
  std::thread reader([&] {
    for (auto& f : frames) in_q.Push(f);
    in_q.SignalProducerIsDone();
  });
  std::thread writer([&] {
    svc::Bytes b;
    while (out_q.Pop(b)) got.push_back(std::move(b));
  });
  enc();
  reader.join();
  writer.join();
  CHECK(got.size() == n);  // header + n-1 frames
  CHECK(enc.frames_encoded() == n - 1);
  uint8_t hdr[32];
  orc_header(n, w, h, pw - w, ph - h, 8, 8, 3, hdr);
  CHECK(got[0].size() == 32 && std::memcmp(got[0].data(), hdr, 32) == 0);
  const uint64_t fbytes = orc_serialized_frame_bytes(w, h, 8, 8, 3);
  double max_err = 0;
  for (uint i = 1; i < n; ++i) {
    const uchar* t[4] = {pyr[i - 1][0].data(), pyr[i - 1][1].data(), pyr[i - 1][2].data(), pyr[i - 1][3].data()};
    const uchar* a[4] = {pyr[i][0].data(), pyr[i][1].data(), pyr[i][2].data(), pyr[i][3].data()};
    std::vector<float> emv(2 * mvn), emad(mvn);
    orc_hbma(t, a, L, pw, ph, R, 16, 16, emv.data(), emad.data());
    std::vector<uint> bt(mvn);
    classify(reinterpret_cast<const Vec2f*>(emv.data()), emad.data(), mw, mh, bt.data());
    std::vector<float> planes((size_t)3 * pw * ph);
    float* pl[3] = {planes.data(), planes.data() + (size_t)pw * ph, planes.data() + (size_t)2 * pw * ph};
    CHECK(orc_dct_planar(frames[i].data(), w, h, pw, ph, 8, 8, pl) == 0);
    std::vector<uint8_t> exp(fbytes);
    orc_serialize_frame(pl, (uint64_t)pw * ph, 3, bt.data(), w, h, 8, 8, mw, 16, 16, exp.data());
    CHECK(got[i].size() == fbytes);
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(got[i].data());
    const uint32_t* ew = reinterpret_cast<const uint32_t*>(exp.data());
    for (uint64_t k = 0; k < fbytes / 4; ++k) {
      if (k % 193 == 0) {
        CHECK(gw[k] == ew[k]);  // block type (from the bit-exact motion field)
      } else {
        float g, e;
        std::memcpy(&g, &gw[k], 4);
        std::memcpy(&e, &ew[k], 4);
        max_err = std::max(max_err, (double)std::fabs(g - e));
      }
    }
  }
  CHECK(max_err <= 1e-3);  // DCT tolerance (absolute, coefficients up to 2040)
  std::printf("PASS frames=%u max_dct_err=%.3g\n", n - 1, max_err);
  return 0;
}
