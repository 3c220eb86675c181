// svc_api.cu -- the C ABI of libsvc_b200.so (include/svc_b200.h): argument
// validation, device memory/stream management and the launch sequence of the
// three hot-path kernels.  No CPU implementation of any stage exists here: a
// missing device or a CUDA failure is reported, never worked around.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stddef.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/svc_b200.h"
#pragma GCC visibility pop
#include "common.cuh"

using namespace svc;

namespace {

thread_local std::string g_err;
thread_local int g_device = 0;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  cudaGetLastError();  // clear sticky-less errors
  return SVC_ERR_CUDA;
}

#define CU(call)                                       \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// once-per-device kernel attributes; sessions on several host threads may get here concurrently
std::mutex g_prepare_mutex;
bool g_prepared[64] = {};

int prepare_device(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (dev < 0 || dev >= n) return fail(SVC_ERR_CUDA, "no such CUDA device: " + std::to_string(dev));
  CU(cudaSetDevice(dev));
  std::lock_guard<std::mutex> lock(g_prepare_mutex);
  if (dev >= 64 || !g_prepared[dev]) {
    CU(prepare_dct_kernels());
    if (dev < 64) g_prepared[dev] = true;
  }
  return SVC_OK;
}

// The stateless entry points run on a per-thread, per-device non-blocking stream: two host threads
// calling them do not serialise on the legacy default stream (nor synchronise with it).
struct ThreadStreams {
  cudaStream_t st[64] = {};
  ~ThreadStreams() {
    for (int d = 0; d < 64; ++d)
      if (st[d] && cudaSetDevice(d) == cudaSuccess) cudaStreamDestroy(st[d]);
  }
};
thread_local ThreadStreams g_streams;

int thread_stream(int dev, cudaStream_t* out) {
  if (dev < 0 || dev >= 64) {
    *out = cudaStreamPerThread;
    return SVC_OK;
  }
  if (!g_streams.st[dev]) CU(cudaStreamCreateWithFlags(&g_streams.st[dev], cudaStreamNonBlocking));
  *out = g_streams.st[dev];
  return SVC_OK;
}

// Preconditions of EstimateMotionHierarchical (libs/motion.cpp:417-433) plus
// the divisibility the reference silently assumes (block / 2^(L-1) != 0).
int check_hbma_args(uint32_t levels, uint32_t fw, uint32_t fh, uint32_t range,
                    uint32_t bw, uint32_t bh, bool allow_zero_range) {
  if (levels < 1 || levels > SVC_MAX_LEVELS)
    return fail(SVC_ERR_INVALID_ARG, "level_count must be in [1, 8]");
  if (fw == 0 || fh == 0 || bw == 0 || bh == 0)
    return fail(SVC_ERR_INVALID_ARG, "frame and block dimensions must be > 0");
  if (fw % bw || fh % bh)
    return fail(SVC_ERR_INVALID_ARG, "frame dimensions must be divisible by the block dimensions");
  const uint32_t red = 1u << (levels - 1);
  if (bw % red || bh % red)
    return fail(SVC_ERR_INVALID_ARG, "block dimensions must be divisible by 2^(level_count-1)");
  if (!allow_zero_range && range < red)
    return fail(SVC_ERR_INVALID_ARG, "search_range must be >= 2^(level_count-1)");
  if (bw > 128 || bh > 128)
    return fail(SVC_ERR_UNSUPPORTED, "block dimensions above 128 are not supported");
  if (fw > 32768 || fh > 32768)
    return fail(SVC_ERR_UNSUPPORTED, "frame dimensions above 32768 are not supported");
  if (range > 4096) return fail(SVC_ERR_UNSUPPORTED, "search_range above 4096 is not supported");
  return SVC_OK;
}

// Device scratch of the stateless entry points: a small per-thread, per-device cache of
// grow-only allocations, so that a host calling e.g. svc_estimate_motion_hierarchical once
// per frame (drop-in 1) does not pay cudaMalloc/cudaFree on every call.
// A DevBuf borrows one cache slot for its lifetime (calls on one thread do not nest).
struct ScratchSlot {
  void* p = nullptr;
  size_t cap = 0;
  int device = -1;
  bool busy = false;
};
constexpr int kScratchSlots = 6;
struct ScratchCache {
  ScratchSlot slot[kScratchSlots];
  ~ScratchCache() {
    for (auto& s : slot)
      if (s.p && cudaSetDevice(s.device) == cudaSuccess) cudaFree(s.p);
  }
};
thread_local ScratchCache g_scratch;

struct DevBuf {
  void* p = nullptr;
  ScratchSlot* slot = nullptr;
  ~DevBuf() {
    if (slot) slot->busy = false;
  }
  cudaError_t alloc(size_t n) {
    if (n == 0) n = 1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    ScratchSlot* pick = nullptr;
    for (auto& s : g_scratch.slot)  // smallest free slot on this device that is large enough
      if (!s.busy && s.device == dev && s.cap >= n && (!pick || s.cap < pick->cap)) pick = &s;
    if (!pick) {
      for (auto& s : g_scratch.slot)  // else recycle the smallest free slot (or an empty one)
        if (!s.busy && (!pick || s.cap < pick->cap)) pick = &s;
      if (!pick) return cudaErrorMemoryAllocation;
      if (pick->p) {
        if (cudaSetDevice(pick->device) == cudaSuccess) cudaFree(pick->p);
        cudaSetDevice(dev);
        pick->p = nullptr;
        pick->cap = 0;
      }
      e = cudaMalloc(&pick->p, n);
      if (e != cudaSuccess) {
        pick->p = nullptr;
        return e;
      }
      pick->cap = n;
      pick->device = dev;
    }
    pick->busy = true;
    slot = pick;
    p = pick->p;
    return cudaSuccess;
  }
  template <class T>
  T* as() { return static_cast<T*>(p); }
};

constexpr size_t kSlack = 256;  // kernels may read one aligned word past a row

}  // namespace

// =============================================================================
extern "C" {

const char* svc_last_error(void) { return g_err.c_str(); }
const char* svc_version(void) { return "svc_b200 0.1 (sm_100a)"; }

int svc_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (count) *count = n;
  return SVC_OK;
}

int svc_set_device(int device) {
  if (device < 0) return fail(SVC_ERR_INVALID_ARG, "device must be >= 0");
  g_device = device;
  return SVC_OK;
}

uint32_t svc_padded_dim(uint32_t a, uint32_t mv_block, uint32_t levels) {
  if (mv_block == 0 || levels == 0 || levels > SVC_MAX_LEVELS) return 0;
  const uint32_t l = std::lcm(mv_block, 1u << (levels - 1));
  return (a + l - 1) / l * l;
}

uint64_t svc_serialized_frame_bytes(uint32_t w, uint32_t h, uint32_t tbw,
                                    uint32_t tbh, uint32_t channels) {
  if (!tbw || !tbh) return 0;
  const uint64_t nx = (w + tbw - 1) / tbw, ny = (h + tbh - 1) / tbh;
  return nx * ny * (4 + (uint64_t)channels * tbw * tbh * 4);
}

int svc_write_header(uint32_t n_input_frames, uint32_t frame_w, uint32_t frame_h,
                     uint32_t padded_w, uint32_t padded_h, uint32_t tbw,
                     uint32_t tbh, uint32_t channels, uint8_t out32[32]) {
  if (!out32) return fail(SVC_ERR_INVALID_ARG, "out32 is null");
  if (padded_w < frame_w || padded_h < frame_h)
    return fail(SVC_ERR_INVALID_ARG, "padded dimensions smaller than the frame");
  const uint32_t hdr[8] = {n_input_frames ? n_input_frames - 1 : 0, frame_w, frame_h,
                           padded_w - frame_w, padded_h - frame_h, tbw, tbh, channels};
  memcpy(out32, hdr, 32);
  return SVC_OK;
}

int svc_patch_block_types(uint8_t* frame_stream, uint32_t frame_w, uint32_t frame_h,
                          uint32_t tbw, uint32_t tbh, uint32_t channels,
                          uint32_t mv_block_w, uint32_t mv_block_h,
                          uint32_t mv_field_w, const uint32_t* block_types) {
  if (!frame_stream || !block_types) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  if (!tbw || !tbh || !mv_block_w || !mv_block_h) return fail(SVC_ERR_INVALID_ARG, "zero block size");
  const size_t rec = 4 + (size_t)channels * tbw * tbh * 4;
  uint8_t* p = frame_stream;
  for (uint32_t y = 0; y < frame_h; y += tbh)
    for (uint32_t x = 0; x < frame_w; x += tbw, p += rec)
      memcpy(p, &block_types[(y / mv_block_h) * mv_field_w + x / mv_block_w], 4);
  return SVC_OK;
}

// ---- stateless drop-ins ---------------------------------------------------------

static int hbma_host(const uint8_t* const* tracked, const uint8_t* const* anchor,
                     uint32_t levels, uint32_t fw, uint32_t fh, uint32_t range,
                     uint32_t bw, uint32_t bh, float* mv, float* mad, bool ebma) {
  if (!tracked || !anchor || !mv || !mad) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  for (uint32_t l = 0; l < levels && l < SVC_MAX_LEVELS; ++l)
    if (!tracked[l] || !anchor[l]) return fail(SVC_ERR_INVALID_ARG, "null pyramid level");
  int rc = check_hbma_args(levels, fw, fh, range, bw, bh, ebma);
  if (rc) return rc;
  rc = prepare_device(g_device);
  if (rc) return rc;
  const PyrLayout lay = make_pyr_layout(fw, fh, levels);
  const uint32_t mvw = fw / bw, mvh = fh / bh;
  DevBuf pyr, dmv, dmad;
  CU(pyr.alloc(2 * lay.slot_bytes + kSlack));
  CU(dmv.alloc(sizeof(float2) * mvw * mvh));
  CU(dmad.alloc(sizeof(float) * mvw * mvh));
  cudaStream_t st = nullptr;
  rc = thread_stream(g_device, &st);
  if (rc) return rc;
  for (uint32_t l = 0; l < levels; ++l) {
    CU(cudaMemcpy2DAsync(pyr.as<uint8_t>() + lay.off[l], lay.pitch[l], tracked[l], lay.w[l],
                         lay.w[l], lay.h[l], cudaMemcpyHostToDevice, st));
    CU(cudaMemcpy2DAsync(pyr.as<uint8_t>() + lay.slot_bytes + lay.off[l], lay.pitch[l], anchor[l],
                         lay.w[l], lay.w[l], lay.h[l], cudaMemcpyHostToDevice, st));
  }
  HbmaParams p{};
  p.pyr = pyr.as<uint8_t>();
  p.lay = lay;
  p.bw = bw;
  p.bh = bh;
  p.r = range >> (levels - 1);
  p.mvw = mvw;
  p.mvh = mvh;
  p.mv = dmv.as<float2>();
  p.mad = dmad.as<float>();
  p.n_frames = 1;
  CU(launch_hbma(p, st, nullptr));
  CU(cudaMemcpyAsync(mv, dmv.p, sizeof(float2) * mvw * mvh, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(mad, dmad.p, sizeof(float) * mvw * mvh, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SVC_OK;
}

int svc_estimate_motion_hierarchical(const uint8_t* const* tracked_pyr,
                                     const uint8_t* const* anchor_pyr,
                                     uint32_t level_count, uint32_t frame_w,
                                     uint32_t frame_h, uint32_t search_range,
                                     uint32_t block_w, uint32_t block_h,
                                     float* motion_field_xy, float* min_mad) {
  return hbma_host(tracked_pyr, anchor_pyr, level_count, frame_w, frame_h, search_range,
                   block_w, block_h, motion_field_xy, min_mad, false);
}

int svc_estimate_motion_hierarchical_16x16(const uint8_t* const* tracked_pyr,
                                           const uint8_t* const* anchor_pyr,
                                           uint32_t frame_w, uint32_t frame_h,
                                           uint32_t search_range, float* mv_field_xy,
                                           float* min_mad) {
  return hbma_host(tracked_pyr, anchor_pyr, 4, frame_w, frame_h, search_range, 16, 16,
                   mv_field_xy, min_mad, false);
}

int svc_estimate_motion_exhaustive(const uint8_t* tracked_frame, const uint8_t* anchor_frame,
                                   uint32_t frame_w, uint32_t frame_h, uint32_t search_range,
                                   uint32_t block_w, uint32_t block_h,
                                   float* motion_field_xy, float* min_mad) {
  const uint8_t* t[1] = {tracked_frame};
  const uint8_t* a[1] = {anchor_frame};
  return hbma_host(t, a, 1, frame_w, frame_h, search_range, block_w, block_h,
                   motion_field_xy, min_mad, true);
}

int svc_y_pyramid(const uint8_t* bgr, uint32_t frame_w, uint32_t frame_h, uint32_t padded_w,
                  uint32_t padded_h, uint32_t level_count, uint8_t* const* out_levels) {
  if (!bgr || !out_levels) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  if (level_count < 1 || level_count > SVC_MAX_LEVELS)
    return fail(SVC_ERR_INVALID_ARG, "level_count must be in [1, 8]");
  for (uint32_t l = 0; l < level_count; ++l)
    if (!out_levels[l]) return fail(SVC_ERR_INVALID_ARG, "null output level");
  if (!frame_w || !frame_h || padded_w < frame_w || padded_h < frame_h)
    return fail(SVC_ERR_INVALID_ARG, "bad frame / padded dimensions");
  const uint32_t red = 1u << (level_count - 1);
  if (padded_w % red || padded_h % red)
    return fail(SVC_ERR_INVALID_ARG, "padded dimensions must be divisible by 2^(level_count-1)");
  if (padded_w > 32768 || padded_h > 32768) return fail(SVC_ERR_UNSUPPORTED, "frame too large");
  int rc = prepare_device(g_device);
  if (rc) return rc;
  const PyrLayout lay = make_pyr_layout(padded_w, padded_h, level_count);
  DevBuf in, pyr;
  const size_t in_bytes = (size_t)frame_w * frame_h * 3;
  CU(in.alloc(in_bytes));
  CU(pyr.alloc(lay.slot_bytes + kSlack));
  cudaStream_t st = nullptr;
  rc = thread_stream(g_device, &st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(in.p, bgr, in_bytes, cudaMemcpyHostToDevice, st));
  CU(launch_bgr_to_y(in.as<uint8_t>(), frame_w, frame_h, pyr.as<uint8_t>(), lay, 0, 1, st));
  CU(launch_pyr_levels(pyr.as<uint8_t>(), lay, 0, 1, st, nullptr));
  for (uint32_t l = 0; l < level_count; ++l)
    CU(cudaMemcpy2DAsync(out_levels[l], lay.w[l], pyr.as<uint8_t>() + lay.off[l], lay.pitch[l],
                         lay.w[l], lay.h[l], cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SVC_OK;
}

static int check_dct_args(uint32_t w, uint32_t h, uint32_t pw, uint32_t ph, uint32_t tbw, uint32_t tbh) {
  if (!w || !h || pw < w || ph < h) return fail(SVC_ERR_INVALID_ARG, "bad frame / padded dimensions");
  if (!tbw || !tbh) return fail(SVC_ERR_INVALID_ARG, "transform block dimensions must be > 0");
  if (pw % tbw || ph % tbh)
    return fail(SVC_ERR_INVALID_ARG, "padded dimensions must be divisible by the transform block");
  if (tbw > 32 || tbh > 32) return fail(SVC_ERR_UNSUPPORTED, "transform blocks above 32x32 are not supported");
  if (pw > 32768 || ph > 32768) return fail(SVC_ERR_UNSUPPORTED, "frame too large");
  return SVC_OK;
}

int svc_dct_planar(const uint8_t* bgr, uint32_t frame_w, uint32_t frame_h, uint32_t padded_w,
                   uint32_t padded_h, uint32_t tbw, uint32_t tbh, float* const* planes) {
  if (!bgr || !planes || !planes[0] || !planes[1] || !planes[2])
    return fail(SVC_ERR_INVALID_ARG, "null pointer");
  int rc = check_dct_args(frame_w, frame_h, padded_w, padded_h, tbw, tbh);
  if (rc) return rc;
  rc = prepare_device(g_device);
  if (rc) return rc;
  const size_t in_bytes = (size_t)frame_w * frame_h * 3;
  const size_t plane_elems = (size_t)padded_w * padded_h;
  DevBuf in, out, tmp;
  CU(in.alloc(in_bytes));
  CU(out.alloc(plane_elems * 3 * sizeof(float)));
  CU(tmp.alloc(plane_elems * 3 * sizeof(float)));
  cudaStream_t st = nullptr;
  rc = thread_stream(g_device, &st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(in.p, bgr, in_bytes, cudaMemcpyHostToDevice, st));
  DctParams p{};
  p.bgr = in.as<uint8_t>();
  p.w = frame_w; p.h = frame_h; p.pw = padded_w; p.ph = padded_h;
  p.tbw = tbw; p.tbh = tbh;
  p.n_frames = 1;
  p.planes = out.as<float>();
  p.scratch_planes = tmp.as<float>();
  p.scratch_frames = 1;
  CU(launch_dct(p, st, nullptr));
  for (int c = 0; c < 3; ++c)
    CU(cudaMemcpyAsync(planes[c], out.as<float>() + c * plane_elems, plane_elems * sizeof(float),
                       cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SVC_OK;
}

int svc_encode_frame_stream(const uint8_t* bgr, uint32_t frame_w, uint32_t frame_h,
                            uint32_t padded_w, uint32_t padded_h, uint32_t tbw, uint32_t tbh,
                            uint32_t mv_block_w, uint32_t mv_block_h,
                            const uint32_t* block_types, uint8_t* out) {
  if (!bgr || !out) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  int rc = check_dct_args(frame_w, frame_h, padded_w, padded_h, tbw, tbh);
  if (rc) return rc;
  if (!mv_block_w || !mv_block_h || padded_w % mv_block_w || padded_h % mv_block_h)
    return fail(SVC_ERR_INVALID_ARG, "padded dimensions must be divisible by the mv block");
  rc = prepare_device(g_device);
  if (rc) return rc;
  const size_t in_bytes = (size_t)frame_w * frame_h * 3;
  const size_t plane_elems = (size_t)padded_w * padded_h;
  const uint64_t sbytes = svc_serialized_frame_bytes(frame_w, frame_h, tbw, tbh, 3);
  const uint32_t mvw = padded_w / mv_block_w, mvh = padded_h / mv_block_h;
  DevBuf in, stream, scratch, bt;
  CU(in.alloc(in_bytes));
  CU(stream.alloc(sbytes));
  CU(scratch.alloc(plane_elems * 6 * sizeof(float)));
  cudaStream_t st = nullptr;
  rc = thread_stream(g_device, &st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(in.p, bgr, in_bytes, cudaMemcpyHostToDevice, st));
  if (block_types) {
    CU(bt.alloc(sizeof(uint32_t) * mvw * mvh));
    CU(cudaMemcpyAsync(bt.p, block_types, sizeof(uint32_t) * mvw * mvh, cudaMemcpyHostToDevice, st));
  }
  DctParams p{};
  p.bgr = in.as<uint8_t>();
  p.w = frame_w; p.h = frame_h; p.pw = padded_w; p.ph = padded_h;
  p.tbw = tbw; p.tbh = tbh;
  p.n_frames = 1;
  p.stream = stream.as<uint8_t>();
  p.frame_stream_bytes = sbytes;
  p.block_types = block_types ? bt.as<uint32_t>() : nullptr;
  p.mv_block_w = mv_block_w; p.mv_block_h = mv_block_h;
  p.mv_field_w = mvw; p.mv_field_h = mvh;
  p.scratch_planes = scratch.as<float>();
  p.scratch_frames = 1;
  CU(launch_dct(p, st, nullptr));
  CU(cudaMemcpyAsync(out, stream.p, sbytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SVC_OK;
}

int svc_stream_layout_from_header(const uint8_t header32[32], svc_stream_layout* out) {
  if (!header32 || !out) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  uint32_t h[8];
  memcpy(h, header32, 32);  // libs/codec.hpp:8-17
  if (!h[5] || !h[6] || !h[7] || !h[1] || !h[2]) return fail(SVC_ERR_INVALID_ARG, "malformed header");
  svc_stream_layout L{};
  L.frame_count = h[0]; L.frame_w = h[1]; L.frame_h = h[2];
  L.padded_w = h[1] + h[3]; L.padded_h = h[2] + h[4];
  L.tbw = h[5]; L.tbh = h[6]; L.channels = h[7];
  L.record_bytes = 4 + L.channels * L.tbw * L.tbh * 4;
  L.encoder_records_per_frame = (uint64_t)((L.frame_w + L.tbw - 1) / L.tbw) * ((L.frame_h + L.tbh - 1) / L.tbh);
  L.decoder_records_per_frame = (uint64_t)((L.padded_w + L.tbw - 1) / L.tbw) * ((L.padded_h + L.tbh - 1) / L.tbh);
  L.encoder_stream_bytes = 32 + (uint64_t)L.frame_count * L.encoder_records_per_frame * L.record_bytes;
  L.consistent = L.encoder_records_per_frame == L.decoder_records_per_frame;
  *out = L;
  return SVC_OK;
}

// ---- decoder block path ---------------------------------------------------------------

int svc_gaze_rect(uint32_t gaze_x, uint32_t gaze_y, uint32_t max_w, uint32_t max_h, uint32_t frame_w,
                  uint32_t frame_h, uint32_t padded_w, uint32_t padded_h, svc_rect* out) {
  if (!out) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  if (!frame_w || !frame_h || gaze_x >= frame_w || gaze_y >= frame_h)
    return fail(SVC_ERR_INVALID_ARG, "gaze position outside the frame");
  // CalcWithinFrameRectFromCenter, libs/decoder.cpp:66-98
  uint32_t half_w = (max_w + 1) / 2;
  if (gaze_x + half_w >= frame_w) half_w = frame_w - gaze_x - 1;
  if (gaze_x < half_w) half_w = gaze_x;
  uint32_t half_h = (max_h + 1) / 2;
  if (gaze_y + half_h >= frame_h) half_h = frame_h - gaze_y - 1;
  if (gaze_y < half_h) half_h = gaze_y;
  // scaling to the padded frame, libs/decoder.cpp:172-189 (RoundFloatToInt = std::round)
  const float wr = (float)padded_w / (float)frame_w, hr = (float)padded_h / (float)frame_h;
  out->x = (uint32_t)(int)roundf((float)(gaze_x - half_w) * wr);
  out->y = (uint32_t)(int)roundf((float)(gaze_y - half_h) * hr);
  out->w = (uint32_t)(int)roundf((float)(2 * half_w) * wr);
  out->h = (uint32_t)(int)roundf((float)(2 * half_h) * hr);
  return SVC_OK;
}

static int check_decode_args(uint32_t pw, uint32_t ph, uint32_t tbw, uint32_t tbh, uint32_t fg_q,
                             uint32_t bg_q) {
  if (fg_q == 0) return fail(SVC_ERR_INVALID_ARG, "invalid foreground quantization step: must be > 0");
  if (bg_q == 0) return fail(SVC_ERR_INVALID_ARG, "invalid background quantization step: must be > 0");
  if (tbw != tbh || (tbw != 8 && tbw != 16 && tbw != 4))
    return fail(SVC_ERR_UNSUPPORTED, "the decoder block path supports 8x8, 16x16 and 4x4 transform blocks only");
  if (!pw || !ph || pw % tbw || ph % tbh)
    return fail(SVC_ERR_INVALID_ARG, "padded dimensions must be multiples of the transform block");
  if (pw > 32768 || ph > 32768) return fail(SVC_ERR_UNSUPPORTED, "frame too large");
  return SVC_OK;
}

int svc_decode_frames_device(int device, void* cuda_stream, const uint8_t* d_records, uint32_t n_frames,
                             uint32_t padded_w, uint32_t padded_h, uint32_t tbw, uint32_t tbh,
                             uint32_t fg_quant_step, uint32_t bg_quant_step, const svc_rect* gaze,
                             float* d_out_bgr) {
  if (!d_records || !d_out_bgr) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  int rc = check_decode_args(padded_w, padded_h, tbw, tbh, fg_quant_step, bg_quant_step);
  if (rc) return rc;
  if (reinterpret_cast<uintptr_t>(d_out_bgr) & 15u) return fail(SVC_ERR_INVALID_ARG, "output must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(d_records) & 3u) return fail(SVC_ERR_INVALID_ARG, "records must be 4-byte aligned");
  rc = prepare_device(device);
  if (rc) return rc;
  DecodeParams p{};
  p.records = d_records;
  p.frame_record_bytes = (uint64_t)(padded_w / tbw) * (padded_h / tbh) * (4u + 12u * tbw * tbh);
  p.pw = padded_w; p.ph = padded_h;
  p.tb = tbw;
  p.n_frames = n_frames;
  p.fg_q = fg_quant_step; p.bg_q = bg_quant_step;
  if (gaze) { p.has_gaze = 1; p.gaze_x = gaze->x; p.gaze_y = gaze->y; p.gaze_w = gaze->w; p.gaze_h = gaze->h; }
  p.out = d_out_bgr;
  CU(launch_decode(p, static_cast<cudaStream_t>(cuda_stream)));
  return SVC_OK;
}

int svc_decode_frame_blocks(const uint8_t* frame_records, uint32_t padded_w, uint32_t padded_h,
                            uint32_t tbw, uint32_t tbh, uint32_t fg_quant_step, uint32_t bg_quant_step,
                            const svc_rect* gaze, float* out_bgr) {
  if (!frame_records || !out_bgr) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  int rc = check_decode_args(padded_w, padded_h, tbw, tbh, fg_quant_step, bg_quant_step);
  if (rc) return rc;
  rc = prepare_device(g_device);
  if (rc) return rc;
  const size_t in_bytes = (size_t)(padded_w / tbw) * (padded_h / tbh) * (4u + 12u * tbw * tbh);
  const size_t out_bytes = (size_t)padded_w * padded_h * 3 * sizeof(float);
  DevBuf in, out;
  CU(in.alloc(in_bytes));
  CU(out.alloc(out_bytes));
  cudaStream_t st = nullptr;
  rc = thread_stream(g_device, &st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(in.p, frame_records, in_bytes, cudaMemcpyHostToDevice, st));
  rc = svc_decode_frames_device(g_device, st, in.as<uint8_t>(), 1, padded_w, padded_h, tbw, tbh,
                                fg_quant_step, bg_quant_step, gaze, out.as<float>());
  if (rc) return rc;
  CU(cudaMemcpyAsync(out_bgr, out.p, out_bytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SVC_OK;
}

int svc_selftest_dequant(int device, uint32_t q_lo, uint32_t q_hi, uint64_t* mismatches) {
  if (!mismatches || q_lo == 0 || q_hi < q_lo) return fail(SVC_ERR_INVALID_ARG, "bad arguments");
  int rc = prepare_device(device);
  if (rc) return rc;
  cudaStream_t st = nullptr;
  rc = thread_stream(device, &st);
  if (rc) return rc;
  DevBuf cnt;
  CU(cnt.alloc(2 * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(unsigned long long), st));
  CU(run_dequant_selftest(q_lo, q_hi, cnt.as<unsigned long long>(), st));
  unsigned long long h[2] = {0, 0};
  CU(cudaMemcpyAsync(h, cnt.p, sizeof h, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *mismatches = h[0];
  if (h[0]) {
    char msg[128];
    snprintf(msg, sizeof msg, "dequantiser self-test: %llu mismatches, e.g. step %u, coefficient bits 0x%08x",
             h[0], (unsigned)(h[1] >> 32), (unsigned)(h[1] & 0xffffffffu));
    g_err = msg;
  }
  return SVC_OK;
}

// ---- memory helpers -------------------------------------------------------------

void* svc_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void svc_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
void* svc_device_alloc(int device, size_t bytes) {
  void* p = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void svc_device_free(int device, void* p) {
  if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p);
}
int svc_memcpy_h2d(int device, void* dst, const void* src, size_t bytes) {
  CU(cudaSetDevice(device));
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return SVC_OK;
}
int svc_memcpy_d2h(int device, void* dst, const void* src, size_t bytes) {
  CU(cudaSetDevice(device));
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return SVC_OK;
}

}  // extern "C"

// =============================================================================
// Session
// =============================================================================
struct svc_session {
  svc_session_config cfg{};
  svc_session_info info{};
  PyrLayout lay{};
  int device = 0;
  cudaStream_t stream = nullptr;  // compute
  bool own_stream = false;
  cudaStream_t s_in = nullptr, s_out = nullptr;  // host path copy streams
  // Two pyramid slot arrays (max_batch + 1 slots each), used by alternate batches so that
  // K3 of batch k+1 (HBM-bound, main stream) overlaps K1b/K2 of batch k (ALU/latency-bound,
  // motion stream).  d_pyr aliases array 0 (stage entry points, work counters).
  uint8_t* d_pyrs[2] = {nullptr, nullptr};
  uint8_t* d_pyr = nullptr;
  cudaStream_t s_aux = nullptr;            // motion stream: pyrDown + HBMA
  cudaEvent_t ev_y = nullptr;              // level-0 luma of the current batch written
  cudaEvent_t ev_copy = nullptr;           // previous frame's pyramid handed over
  cudaEvent_t ev_motion[2] = {nullptr, nullptr};  // motion of the batch that used array i done
  bool motion_pending[2] = {false, false};
  bool copy_pending = false;
  uint32_t batch_idx = 0;
  uint32_t prev_array = 0, prev_slot = 0;  // where the previous frame's pyramid lives
  bool have_prev = false;
  float* d_scratch = nullptr;  // generic DCT path only
  uint32_t scratch_frames = 0;
  // host path staging (double buffered), allocated on first use
  static constexpr int kInRing = 8;  // input staging slots: the upload runs up to 8 chunks ahead of the kernels (24 measured: no gain)
  uint8_t* d_in[kInRing] = {};
  cudaEvent_t ev_in_ring[kInRing] = {}, ev_used[kInRing] = {};
  uint32_t in_ring = 2;
  float* d_mv[2] = {nullptr, nullptr};
  float* d_mad[2] = {nullptr, nullptr};
  uint8_t* d_st[2] = {nullptr, nullptr};
  uint32_t* d_bt[kInRing] = {};  // (one per input slot)
  cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};
  bool staging = false;
  uint32_t host_chunk = 16;  // frames per pipeline stage of svc_session_encode
  uint64_t launches = 0;
};

namespace {

// Scratch planes of the generic transform / serializer path, allocated on first use.
int ensure_scratch(svc_session* s, DctParams& dp) {
  if (!dct_needs_scratch(dp)) return SVC_OK;
  if (!s->d_scratch) {
    s->scratch_frames = std::min<uint32_t>(s->info.max_batch, 16);
    CU(cudaMalloc(&s->d_scratch, (size_t)s->scratch_frames * 6 * dp.pw * dp.ph * sizeof(float)));
  }
  dp.scratch_planes = s->d_scratch;
  dp.scratch_frames = s->scratch_frames;
  return SVC_OK;
}

// One batch (<= max_batch frames), everything on the device.  K3 (+ fused luma) runs on
// s->stream, the pyramid levels and the motion search on s->s_aux; the caller joins
// (join_motion) before anything that consumes motion vectors.
int encode_batch_device(svc_session* s, const uint8_t* d_frames, uint32_t m, float* d_mv,
                        float* d_mad, uint8_t* d_stream, const uint32_t* d_bt,
                        uint32_t* n_enc_out) {
  const uint32_t first_slot = s->have_prev ? 1u : 0u;
  const uint32_t n_enc = s->have_prev ? m : m - 1;
  const uint32_t cur = s->batch_idx & 1u;
  uint8_t* pyr = s->d_pyrs[cur];
  int nl = 0;
  DctParams dp{};
  if (n_enc && d_stream) {
    dp.bgr = d_frames + (size_t)(m - n_enc) * s->info.frame_in_bytes;
    dp.w = s->cfg.frame_w; dp.h = s->cfg.frame_h;
    dp.pw = s->info.padded_w; dp.ph = s->info.padded_h;
    dp.tbw = s->cfg.transform_block_w; dp.tbh = s->cfg.transform_block_h;
    dp.n_frames = n_enc;
    dp.stream = d_stream;
    dp.frame_stream_bytes = s->info.frame_stream_bytes;
    dp.block_types = d_bt;
    dp.mv_block_w = s->cfg.mv_block_w; dp.mv_block_h = s->cfg.mv_block_h;
    dp.mv_field_w = s->info.mv_field_w; dp.mv_field_h = s->info.mv_field_h;
    dp.scratch_planes = s->d_scratch;
    dp.scratch_frames = s->scratch_frames;
  }
  if (n_enc && d_stream) {  // generic transform blocks, padded widths, unaligned frame pointers
    int rcs = ensure_scratch(s, dp);
    if (rcs) return rcs;
  }
  const bool fuse_y = n_enc && d_stream && dct_can_fuse_y(dp);

  // ---- motion stream, part 1: hand the previous frame's pyramid to slot 0 of this array
  // (libs/encoder.cpp:661-663 ping-pong).  Ordered after the previous batch's pyramid
  // build and after the last search that read this array, both on s_aux.
  // The copy issued by the PREVIOUS batch read the array this batch is about to overwrite.
  if (s->copy_pending) CU(cudaStreamWaitEvent(s->stream, s->ev_copy, 0));
  if (s->have_prev) {
    CU(launch_copy_slot(pyr, s->d_pyrs[s->prev_array] + (size_t)s->prev_slot * s->lay.slot_bytes,
                        s->lay.slot_bytes, s->s_aux));
    nl += 1;
    CU(cudaEventRecord(s->ev_copy, s->s_aux));
    s->copy_pending = true;
  }
  // ---- main stream: level-0 luma (fused into K3 when possible)
  if (s->motion_pending[cur]) CU(cudaStreamWaitEvent(s->stream, s->ev_motion[cur], 0));  // array free
  if (fuse_y) {
    // K3 also emits the level-0 luma of every encoded frame (slots 1..n_enc); a
    // tracked-only first frame (slot 0) still needs the stand-alone conversion.
    if (!s->have_prev) {
      CU(launch_bgr_to_y(d_frames, s->cfg.frame_w, s->cfg.frame_h, pyr, s->lay, 0, 1, s->stream));
      nl += 1;
    }
    dp.y_l0 = pyr + s->lay.off[0];
    dp.y_slot_bytes = s->lay.slot_bytes;
    dp.y_first_slot = 1;
    dp.y_pitch = s->lay.pitch[0];
    CU(launch_dct(dp, s->stream, &nl));
  } else {
    CU(launch_bgr_to_y(d_frames, s->cfg.frame_w, s->cfg.frame_h, pyr, s->lay, first_slot, m, s->stream));
    nl += 1;
  }
  CU(cudaEventRecord(s->ev_y, s->stream));
  if (n_enc && d_stream && !fuse_y) CU(launch_dct(dp, s->stream, &nl));

  // ---- motion stream, part 2: pyramid levels 1.. and the search
  CU(cudaStreamWaitEvent(s->s_aux, s->ev_y, 0));
  CU(launch_pyr_levels(pyr, s->lay, first_slot, m, s->s_aux, &nl));
  if (n_enc && (d_mv || d_mad)) {
    HbmaParams p{};
    p.pyr = pyr;
    p.lay = s->lay;
    p.bw = s->cfg.mv_block_w;
    p.bh = s->cfg.mv_block_h;
    p.r = s->cfg.mv_search_range >> (s->cfg.pyr_lvl_count - 1);
    p.mvw = s->info.mv_field_w;
    p.mvh = s->info.mv_field_h;
    p.mv = reinterpret_cast<float2*>(d_mv);
    p.mad = d_mad;
    p.n_frames = n_enc;
    p.family = s->cfg.hbma_kernel_family;
    CU(launch_hbma(p, s->s_aux, &nl));
  }
  CU(cudaEventRecord(s->ev_motion[cur], s->s_aux));
  s->motion_pending[cur] = true;
  s->prev_array = cur;
  s->prev_slot = first_slot + m - 1;
  s->have_prev = true;
  s->batch_idx += 1;
  s->launches += (uint64_t)nl;
  if (n_enc_out) *n_enc_out = n_enc;
  return SVC_OK;
}

// Make `st` wait for every motion search issued so far.
int join_motion(svc_session* s, cudaStream_t st) {
  for (int i = 0; i < 2; ++i)
    if (s->motion_pending[i]) CU(cudaStreamWaitEvent(st, s->ev_motion[i], 0));
  return SVC_OK;
}

// Order the motion stream after everything already queued on the main stream (results of
// an earlier call may still be read there).
int fork_motion(svc_session* s) {
  CU(cudaEventRecord(s->ev_y, s->stream));
  CU(cudaStreamWaitEvent(s->s_aux, s->ev_y, 0));
  return SVC_OK;
}

void free_staging(svc_session* s) {
  for (int b = 0; b < svc_session::kInRing; ++b) {
    cudaFree(s->d_in[b]);
    cudaFree(s->d_bt[b]);
    s->d_in[b] = nullptr;
    s->d_bt[b] = nullptr;
    if (s->ev_in_ring[b]) cudaEventDestroy(s->ev_in_ring[b]);
    if (s->ev_used[b]) cudaEventDestroy(s->ev_used[b]);
    s->ev_in_ring[b] = s->ev_used[b] = nullptr;
  }
  for (int b = 0; b < 2; ++b) {
    cudaFree(s->d_mv[b]);
    cudaFree(s->d_mad[b]);
    cudaFree(s->d_st[b]);
    s->d_st[b] = nullptr;
    s->d_mv[b] = s->d_mad[b] = nullptr;
    if (s->ev_in[b]) cudaEventDestroy(s->ev_in[b]);
    if (s->ev_comp[b]) cudaEventDestroy(s->ev_comp[b]);
    if (s->ev_out[b]) cudaEventDestroy(s->ev_out[b]);
    s->ev_in[b] = s->ev_comp[b] = s->ev_out[b] = nullptr;
  }
  if (s->s_in) cudaStreamDestroy(s->s_in);
  if (s->s_out) cudaStreamDestroy(s->s_out);
  s->s_in = s->s_out = nullptr;
  s->staging = false;
}

int alloc_staging(svc_session* s) {
  const size_t B = std::min(s->info.max_batch, s->host_chunk);  // svc_session_encode moves host_chunk frames per stage
  const size_t mvn = (size_t)s->info.mv_field_w * s->info.mv_field_h;
  CU(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
  // input ring: as many slots as fit 4 GB (at least 2): inputs are a fifth of the traffic, and an upload
  // that runs ahead keeps host-memory reads out of the way of the record writes that bound the path
  s->in_ring = (uint32_t)std::max<size_t>(2, std::min<size_t>(svc_session::kInRing, (4ull << 30) / std::max<size_t>(1, B * s->info.frame_in_bytes)));
  for (uint32_t b = 0; b < s->in_ring; ++b) {
    CU(cudaMalloc(&s->d_in[b], B * s->info.frame_in_bytes));
    CU(cudaMalloc(&s->d_bt[b], B * mvn * sizeof(uint32_t)));
    CU(cudaEventCreateWithFlags(&s->ev_in_ring[b], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->ev_used[b], cudaEventDisableTiming));
  }
  for (int b = 0; b < 2; ++b) {
    CU(cudaMalloc(&s->d_mv[b], B * mvn * sizeof(float2)));
    CU(cudaMalloc(&s->d_mad[b], B * mvn * sizeof(float)));
    CU(cudaMalloc(&s->d_st[b], B * s->info.frame_stream_bytes));
    CU(cudaEventCreateWithFlags(&s->ev_in[b], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->ev_comp[b], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->ev_out[b], cudaEventDisableTiming));
  }
  return SVC_OK;
}

// Host-path staging, created on the first svc_session_encode; a failure half-way releases what
// was created, so a retry neither leaks nor reuses a partial set.
int ensure_staging(svc_session* s) {
  if (s->staging) return SVC_OK;
  const int rc = alloc_staging(s);
  if (rc != SVC_OK) {
    const std::string keep = g_err;
    free_staging(s);
    cudaGetLastError();
    g_err = keep;
    return rc;
  }
  s->staging = true;
  return SVC_OK;
}

}  // namespace

extern "C" {

int svc_session_create(const svc_session_config* cfg_in, svc_session** out) {
  if (!cfg_in || !out) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  // struct_size versions the struct: fields added after `cuda_stream` default to zero for older callers
  if (cfg_in->struct_size < offsetof(svc_session_config, hbma_kernel_family) ||
      cfg_in->struct_size > sizeof(svc_session_config))
    return fail(SVC_ERR_INVALID_ARG, "svc_session_config.struct_size mismatch");
  svc_session_config cfg_copy{};
  memcpy(&cfg_copy, cfg_in, cfg_in->struct_size);
  cfg_copy.struct_size = sizeof(svc_session_config);
  const svc_session_config* cfg = &cfg_copy;
  if (cfg->hbma_kernel_family > SVC_HBMA_FAMILY_TILE)
    return fail(SVC_ERR_INVALID_ARG, "hbma_kernel_family must be one of SVC_HBMA_FAMILY_*");
  *out = nullptr;
  // Validate(EncoderConfig), libs/encoder.cpp:62-142 (hot-path fields)
  if (cfg->mv_block_w < 1) return fail(SVC_ERR_INVALID_ARG, "invalid mv block width: must be > 0");
  if (cfg->mv_block_h < 1) return fail(SVC_ERR_INVALID_ARG, "invalid mv block height: must be > 0");
  if (cfg->pyr_lvl_count < 1) return fail(SVC_ERR_INVALID_ARG, "invalid pyramid level count: must be > 0");
  if (cfg->pyr_lvl_count > SVC_MAX_LEVELS) return fail(SVC_ERR_UNSUPPORTED, "pyramid level count above 8");
  const uint32_t red = 1u << (cfg->pyr_lvl_count - 1);
  if (cfg->mv_search_range / red == 0)
    return fail(SVC_ERR_INVALID_ARG,
                "invalid mv search and pyramid level count: the quotient from dividing the mv "
                "search range by the pyramid level reduction factor must be > 0");
  if (cfg->transform_block_w < 1) return fail(SVC_ERR_INVALID_ARG, "invalid transform block width: must be > 0");
  if (cfg->transform_block_h < 1) return fail(SVC_ERR_INVALID_ARG, "invalid transform block height: must be > 0");
  if (cfg->transform_block_w > cfg->mv_block_w)
    return fail(SVC_ERR_INVALID_ARG, "transform block width must be <= mv block width");
  if (cfg->transform_block_h > cfg->mv_block_h)
    return fail(SVC_ERR_INVALID_ARG, "transform block height must be <= mv block height");
  if (cfg->mv_block_w % cfg->transform_block_w)
    return fail(SVC_ERR_INVALID_ARG, "mv block width must be divisible by transform block width");
  if (cfg->mv_block_h % cfg->transform_block_h)
    return fail(SVC_ERR_INVALID_ARG, "mv block height must be divisible by transform block height");
  if (cfg->frame_w == 0 || cfg->frame_h == 0) return fail(SVC_ERR_INVALID_ARG, "frame dimensions must be > 0");
  const uint32_t pw = svc_padded_dim(cfg->frame_w, cfg->mv_block_w, cfg->pyr_lvl_count);
  const uint32_t ph = svc_padded_dim(cfg->frame_h, cfg->mv_block_h, cfg->pyr_lvl_count);
  int rc = check_hbma_args(cfg->pyr_lvl_count, pw, ph, cfg->mv_search_range, cfg->mv_block_w,
                           cfg->mv_block_h, false);
  if (rc) return rc;
  rc = check_dct_args(cfg->frame_w, cfg->frame_h, pw, ph, cfg->transform_block_w, cfg->transform_block_h);
  if (rc) return rc;
  rc = prepare_device(cfg->device);
  if (rc) return rc;

  svc_session* s = new svc_session();
  s->cfg = *cfg;
  s->device = cfg->device;
  s->info.padded_w = pw;
  s->info.padded_h = ph;
  s->info.mv_field_w = pw / cfg->mv_block_w;
  s->info.mv_field_h = ph / cfg->mv_block_h;
  s->info.frame_in_bytes = (uint64_t)cfg->frame_w * cfg->frame_h * 3;
  s->info.frame_stream_bytes = svc_serialized_frame_bytes(cfg->frame_w, cfg->frame_h,
                                                          cfg->transform_block_w,
                                                          cfg->transform_block_h, 3);
  s->info.record_bytes = 4 + 3 * cfg->transform_block_w * cfg->transform_block_h * 4;
  s->info.max_batch = cfg->max_batch ? cfg->max_batch : 32;
  s->lay = make_pyr_layout(pw, ph, cfg->pyr_lvl_count);
  auto bail = [&](int code) {
    svc_session_destroy(s);
    return code;
  };
  cudaError_t e;
  if (cfg->cuda_stream) {
    s->stream = static_cast<cudaStream_t>(cfg->cuda_stream);
  } else {
    e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate"));
    s->own_stream = true;
  }
  const size_t pyr_bytes = (size_t)(s->info.max_batch + 1) * s->lay.slot_bytes + kSlack;
  for (int i = 0; i < 2; ++i) {
    e = cudaMalloc(&s->d_pyrs[i], pyr_bytes);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(pyramids)"));
    e = cudaMemsetAsync(s->d_pyrs[i], 0, pyr_bytes, s->stream);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemset(pyramids)"));
    e = cudaEventCreateWithFlags(&s->ev_motion[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
  }
  s->d_pyr = s->d_pyrs[0];
  {
    // highest priority: as K3's CTAs retire, the pending pyrDown / HBMA CTAs of the previous
    // batch are placed first, so the ALU-bound motion work co-runs with the HBM-bound K3
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    e = cudaStreamCreateWithPriority(&s->s_aux, cudaStreamNonBlocking, prio_hi);
  }
  if (e != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate(motion)"));
  e = cudaEventCreateWithFlags(&s->ev_y, cudaEventDisableTiming);
  if (e != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
  e = cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming);
  if (e != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
  e = cudaStreamSynchronize(s->stream);  // the memsets, before the motion stream may touch the arrays
  if (e != cudaSuccess) return bail(cuda_fail(e, "cudaStreamSynchronize"));
  // (the scratch planes of the generic transform path are allocated on first use, see
  // encode_batch_device / svc_session_run_stage: the fused stream kernels never touch them)
  if (cfg->host_chunk_frames) s->host_chunk = cfg->host_chunk_frames;
  *out = s;
  return SVC_OK;
}

void svc_session_destroy(svc_session* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->s_in) cudaStreamSynchronize(s->s_in);
  if (s->s_out) cudaStreamSynchronize(s->s_out);
  free_staging(s);
  if (s->s_aux) cudaStreamSynchronize(s->s_aux);
  cudaFree(s->d_pyrs[0]);
  cudaFree(s->d_pyrs[1]);
  if (s->ev_y) cudaEventDestroy(s->ev_y);
  if (s->ev_copy) cudaEventDestroy(s->ev_copy);
  for (int i = 0; i < 2; ++i)
    if (s->ev_motion[i]) cudaEventDestroy(s->ev_motion[i]);
  if (s->s_aux) cudaStreamDestroy(s->s_aux);
  cudaFree(s->d_scratch);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  cudaGetLastError();
  delete s;
}

int svc_session_info_get(const svc_session* s, svc_session_info* out) {
  if (!s || !out) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  *out = s->info;
  return SVC_OK;
}

int svc_session_reset(svc_session* s) {
  if (!s) return fail(SVC_ERR_INVALID_ARG, "null session");
  s->have_prev = false;
  return SVC_OK;
}

uint64_t svc_session_launch_count(const svc_session* s) { return s ? s->launches : 0; }

int svc_sad_peak(int device, double* absdiffs_per_s) {
  if (!absdiffs_per_s) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  int rc = prepare_device(device);
  if (rc) return rc;
  CU(measure_sad_peak(0, absdiffs_per_s));
  return SVC_OK;
}

int svc_session_hbma_work(svc_session* s, uint32_t n_frames, uint64_t* candidates, uint64_t* absdiffs) {
  if (!s || !candidates || !absdiffs) return fail(SVC_ERR_INVALID_ARG, "null pointer");
  if (n_frames == 0 || n_frames > s->info.max_batch)
    return fail(SVC_ERR_INVALID_ARG, "n_frames must be in [1, max_batch]");
  CU(cudaSetDevice(s->device));
  DevBuf cnt;
  CU(cnt.alloc(2 * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(unsigned long long), s->stream));
  HbmaParams p{};
  p.pyr = s->d_pyr;
  p.lay = s->lay;
  p.bw = s->cfg.mv_block_w;
  p.bh = s->cfg.mv_block_h;
  p.r = s->cfg.mv_search_range >> (s->cfg.pyr_lvl_count - 1);
  p.mvw = s->info.mv_field_w;
  p.mvh = s->info.mv_field_h;
  p.n_frames = n_frames;
  p.family = s->cfg.hbma_kernel_family;
  p.counters = cnt.as<unsigned long long>();
  int nl = 0;
  CU(launch_hbma(p, s->stream, &nl));
  unsigned long long h[2] = {0, 0};
  CU(cudaMemcpyAsync(h, cnt.p, sizeof h, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  *candidates = h[0];
  *absdiffs = h[1];
  s->launches += (uint64_t)nl;
  return SVC_OK;
}

int svc_session_synchronize(svc_session* s) {
  if (!s) return fail(SVC_ERR_INVALID_ARG, "null session");
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaStreamSynchronize(s->s_aux));
  return SVC_OK;
}

int svc_session_encode_device(svc_session* s, const uint8_t* d_frames, uint32_t n_frames,
                              float* d_mv, float* d_mad, uint8_t* d_stream,
                              const uint32_t* d_bt, uint32_t* n_encoded) {
  if (!s) return fail(SVC_ERR_INVALID_ARG, "null session");
  if (n_encoded) *n_encoded = 0;
  if (n_frames == 0) return SVC_OK;
  if (!d_frames) return fail(SVC_ERR_INVALID_ARG, "null frames");
  CU(cudaSetDevice(s->device));
  const size_t mvn = (size_t)s->info.mv_field_w * s->info.mv_field_h;
  uint32_t done_in = 0, done_enc = 0;
  int rcf = fork_motion(s);
  if (rcf) return rcf;
  while (done_in < n_frames) {
    const uint32_t m = std::min(s->info.max_batch, n_frames - done_in);
    uint32_t ne = 0;
    int rc = encode_batch_device(
        s, d_frames + (size_t)done_in * s->info.frame_in_bytes, m,
        d_mv ? d_mv + done_enc * mvn * 2 : nullptr, d_mad ? d_mad + done_enc * mvn : nullptr,
        d_stream ? d_stream + (size_t)done_enc * s->info.frame_stream_bytes : nullptr,
        d_bt ? d_bt + done_enc * mvn : nullptr, &ne);
    if (rc) return rc;
    done_in += m;
    done_enc += ne;
  }
  // everything the call produced is complete once the session stream gets here
  int rcj = join_motion(s, s->stream);
  if (rcj) return rcj;
  if (n_encoded) *n_encoded = done_enc;
  return SVC_OK;
}

int svc_session_encode(svc_session* s, const uint8_t* frames, uint32_t n_frames, float* mv,
                       float* mad, uint8_t* stream, const uint32_t* block_types,
                       uint32_t* n_encoded) {
  if (!s) return fail(SVC_ERR_INVALID_ARG, "null session");
  if (n_encoded) *n_encoded = 0;
  if (n_frames == 0) return SVC_OK;
  if (!frames) return fail(SVC_ERR_INVALID_ARG, "null frames");
  CU(cudaSetDevice(s->device));
  int rc = ensure_staging(s);
  if (rc) return rc;
  const size_t mvn = (size_t)s->info.mv_field_w * s->info.mv_field_h;
  const size_t fin = s->info.frame_in_bytes, fst = s->info.frame_stream_bytes;
  uint32_t done_in = 0, done_enc = 0, chunk = 0;
  // 3-stage pipeline over chunks: H2D (s_in) | kernels (stream) | D2H (s_out).  Outputs are double
  // buffered (slot b); inputs go through a ring of in_ring slots (slot bi), so the upload -- a fifth of
  // the traffic -- runs well ahead of the kernels instead of in lock step with the record download.
  while (done_in < n_frames) {
    const int b = chunk & 1;
    const uint32_t bi = chunk % s->in_ring;
    // PCIe-bound: small chunks keep H2D | kernels | D2H overlapped and the exposed head
    // and tail of the pipeline short.  The record download (25 MB per 1080p frame) is the critical
    // path and cannot start before the first chunk is uploaded and encoded, so the first chunks ramp up
    // (base/8, base/4, base/2, base frames): the download of chunk c (4x the bytes of an upload per
    // frame) still covers the upload of the twice larger chunk c+1.
    const uint32_t base = std::min(s->info.max_batch, s->host_chunk);
    const uint32_t ramp = chunk < 3 ? std::max(2u, base >> (3 - chunk)) : base;
    const uint32_t m = std::min(std::min(base, ramp), n_frames - done_in);
    const uint32_t ne = s->have_prev ? m : m - 1;
    if (chunk >= s->in_ring) CU(cudaStreamWaitEvent(s->s_in, s->ev_used[bi], 0));  // d_in[bi] consumed
    if (chunk >= 2) CU(cudaStreamWaitEvent(s->stream, s->ev_out[b], 0));           // outputs[b] drained
    CU(cudaMemcpyAsync(s->d_in[bi], frames + (size_t)done_in * fin, (size_t)m * fin,
                       cudaMemcpyHostToDevice, s->s_in));
    if (block_types && ne)
      CU(cudaMemcpyAsync(s->d_bt[bi], block_types + done_enc * mvn, ne * mvn * sizeof(uint32_t),
                         cudaMemcpyHostToDevice, s->s_in));
    CU(cudaEventRecord(s->ev_in_ring[bi], s->s_in));
    CU(cudaStreamWaitEvent(s->stream, s->ev_in_ring[bi], 0));
    uint32_t ne2 = 0;
    rc = encode_batch_device(s, s->d_in[bi], m, (mv || mad) ? s->d_mv[b] : nullptr,
                             mad ? s->d_mad[b] : nullptr, stream ? s->d_st[b] : nullptr,
                             block_types ? s->d_bt[bi] : nullptr, &ne2);
    if (rc) return rc;
    CU(cudaEventRecord(s->ev_used[bi], s->stream));  // (K3 is the only reader of the frames and block types)
    CU(cudaEventRecord(s->ev_comp[b], s->stream));
    CU(cudaStreamWaitEvent(s->s_out, s->ev_comp[b], 0));
    rc = join_motion(s, s->s_out);  // motion vectors come from the motion stream
    if (rc) return rc;
    if (ne2) {
      if (mv)
        CU(cudaMemcpyAsync(mv + done_enc * mvn * 2, s->d_mv[b], ne2 * mvn * sizeof(float2),
                           cudaMemcpyDeviceToHost, s->s_out));
      if (mad)
        CU(cudaMemcpyAsync(mad + done_enc * mvn, s->d_mad[b], ne2 * mvn * sizeof(float),
                           cudaMemcpyDeviceToHost, s->s_out));
      if (stream)
        CU(cudaMemcpyAsync(stream + (size_t)done_enc * fst, s->d_st[b], (size_t)ne2 * fst,
                           cudaMemcpyDeviceToHost, s->s_out));
    }
    CU(cudaEventRecord(s->ev_out[b], s->s_out));
    done_in += m;
    done_enc += ne2;
    ++chunk;
  }
  CU(cudaStreamSynchronize(s->s_in));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaStreamSynchronize(s->s_aux));
  CU(cudaStreamSynchronize(s->s_out));
  if (n_encoded) *n_encoded = done_enc;
  return SVC_OK;
}

int svc_session_run_stage(svc_session* s, int stage, const uint8_t* d_frames, uint32_t n_frames,
                          float* d_mv, float* d_mad, uint8_t* d_stream) {
  if (!s) return fail(SVC_ERR_INVALID_ARG, "null session");
  if (n_frames == 0 || n_frames > s->info.max_batch)
    return fail(SVC_ERR_INVALID_ARG, "n_frames must be in [1, max_batch]");
  CU(cudaSetDevice(s->device));
  int nl = 0;
  switch (stage) {
    case SVC_STAGE_Y_PYRAMID: {
      if (!d_frames) return fail(SVC_ERR_INVALID_ARG, "null frames");
      s->have_prev = false;  // slots of array 0 are overwritten: the encode sequence starts over
      CU(launch_bgr_to_y(d_frames, s->cfg.frame_w, s->cfg.frame_h, s->d_pyr, s->lay, 1, n_frames, s->stream));
      nl += 1;
      CU(launch_pyr_levels(s->d_pyr, s->lay, 1, n_frames, s->stream, &nl));
      break;
    }
    case SVC_STAGE_HBMA: {
      HbmaParams p{};
      p.pyr = s->d_pyr;
      p.lay = s->lay;
      p.bw = s->cfg.mv_block_w;
      p.bh = s->cfg.mv_block_h;
      p.r = s->cfg.mv_search_range >> (s->cfg.pyr_lvl_count - 1);
      p.mvw = s->info.mv_field_w;
      p.mvh = s->info.mv_field_h;
      p.mv = reinterpret_cast<float2*>(d_mv);
      p.mad = d_mad;
      p.n_frames = n_frames;
      p.family = s->cfg.hbma_kernel_family;
      CU(launch_hbma(p, s->stream, &nl));
      break;
    }
    case SVC_STAGE_DCT_STREAM: {
      if (!d_frames || !d_stream) return fail(SVC_ERR_INVALID_ARG, "null pointer");
      DctParams p{};
      p.bgr = d_frames;
      p.w = s->cfg.frame_w; p.h = s->cfg.frame_h;
      p.pw = s->info.padded_w; p.ph = s->info.padded_h;
      p.tbw = s->cfg.transform_block_w; p.tbh = s->cfg.transform_block_h;
      p.n_frames = n_frames;
      p.stream = d_stream;
      p.frame_stream_bytes = s->info.frame_stream_bytes;
      p.mv_block_w = s->cfg.mv_block_w; p.mv_block_h = s->cfg.mv_block_h;
      p.mv_field_w = s->info.mv_field_w; p.mv_field_h = s->info.mv_field_h;
      {
        int rcs = ensure_scratch(s, p);
        if (rcs) return rcs;
      }
      if (dct_can_fuse_y(p)) {  // as inside a step: level-0 luma of slots 1..n rides along
        s->have_prev = false;   // (overwrites slots of array 0)
        p.y_l0 = s->d_pyr + s->lay.off[0];
        p.y_slot_bytes = s->lay.slot_bytes;
        p.y_first_slot = 1;
        p.y_pitch = s->lay.pitch[0];
      }
      CU(launch_dct(p, s->stream, &nl));
      break;
    }
    case SVC_STAGE_PYR_DOWN: {
      s->have_prev = false;
      CU(launch_pyr_levels(s->d_pyr, s->lay, 1, n_frames, s->stream, &nl));
      break;
    }
    default:
      return fail(SVC_ERR_INVALID_ARG, "unknown stage");
  }
  s->launches += (uint64_t)nl;
  return SVC_OK;
}

}  // extern "C"
