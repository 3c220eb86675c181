// common.cuh -- shared declarations of the sm_100a kernels and their launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svc {

constexpr int kMaxLevels = 8;
constexpr int kNumSms = 148;  // B200

// ---- programmatic dependent launch (the motion stream's chain hand-over -> pyramid -> search) ------
// A kernel launched with launch_dependent() may be scheduled while the tail of its predecessor in
// the stream still runs; it must call grid_dependency_wait() before it reads anything the
// predecessor wrote and before its first global write (the predecessor's results are complete and
// visible after it).  grid_dependency_release() lets the NEXT kernel of the chain be scheduled early; it
// is called right AFTER the wait, so that a kernel never starts before the predecessor of its
// predecessor has completed.  Both are no-ops for a kernel launched the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// Device layout of one Y pyramid "slot" (one frame): levels are stored one
// after another, each with a 128-byte-multiple row pitch (TMA needs 16-byte
// strides; the reference keeps tightly packed Mats, libs/encoder.cpp:197-219,
// the stateless entry points repack at the PCIe boundary).
struct PyrLayout {
  uint32_t levels;
  uint32_t w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
  uint64_t off[kMaxLevels];
  uint64_t slot_bytes;  // multiple of 256
};

inline PyrLayout make_pyr_layout(uint32_t pw, uint32_t ph, uint32_t levels) {
  PyrLayout L{};
  L.levels = levels;
  uint64_t off = 0;
  for (uint32_t l = 0; l < levels; ++l) {
    L.w[l] = pw >> l;
    L.h[l] = ph >> l;
    L.pitch[l] = (L.w[l] + 127u) & ~127u;
    L.off[l] = off;
    off += (uint64_t)L.pitch[l] * L.h[l];
    off = (off + 255u) & ~(uint64_t)255u;
  }
  L.slot_bytes = off;
  return L;
}

// ---- K1: zero pad + BGR -> Y (level 0) and pyrDown (levels 1..) -------------
// frames: n x (h*w*3) interleaved BGR; slot s of the pyramid array receives
// frame s - first_slot.
cudaError_t launch_bgr_to_y(const uint8_t* d_bgr, uint32_t w, uint32_t h,
                            uint8_t* d_pyr, const PyrLayout& lay,
                            uint32_t first_slot, uint32_t n_frames,
                            cudaStream_t st);
cudaError_t launch_pyr_down(uint8_t* d_pyr, const PyrLayout& lay,
                            uint32_t src_level, uint32_t first_slot,
                            uint32_t n_frames, cudaStream_t st);

// device-to-device copy of one pyramid slot on the SMs (not on a copy engine)
cudaError_t launch_copy_slot(uint8_t* dst, const uint8_t* src, size_t bytes, cudaStream_t st);
// every level above 0 of n_frames slots (level 0 -> 1, then the smaller levels two per launch)
cudaError_t launch_pyr_levels(uint8_t* d_pyr, const PyrLayout& lay, uint32_t first_slot,
                              uint32_t n_frames, cudaStream_t st, int* n_launches);

// ---- K2: hierarchical block matching ---------------------------------------
// svc_session_config.hbma_kernel_family (include/svc_b200.h: SVC_HBMA_FAMILY_*)
enum : uint32_t { kHbmaAuto = 0, kHbmaGeneric = 1, kHbmaPool = 2, kHbmaWindow = 3, kHbmaTile = 4 };
struct HbmaParams {
  const uint8_t* pyr;  // slot array; frame i: tracked = slot i, anchor = slot i+1
  PyrLayout lay;
  uint32_t bw, bh;  // base-level block size
  uint32_t r;       // search_range / 2^(levels-1), used at every level
  uint32_t mvw, mvh;
  float2* mv;   // n_frames x mvh x mvw, may be null
  float* mad;   // n_frames x mvh x mvw, may be null
  uint32_t n_frames;
  uint32_t family;  // kHbma*: restrict the dispatcher to one kernel family (test hook; 0 = automatic)
  // optional exact work counters: [0] += candidates, [1] += byte-absdiffs
  unsigned long long* counters;
};
cudaError_t launch_hbma(const HbmaParams& p, cudaStream_t st, int* n_launches);
// dependency-free VABSDIFF4.ACC loop: measured packed-byte SAD peak of the device (absdiffs/s)
cudaError_t measure_sad_peak(cudaStream_t st, double* absdiffs_per_s);

// ---- K3: block DCT + stream layout -------------------------------------------
struct DctParams {
  const uint8_t* bgr;  // n x (h*w*3)
  uint32_t w, h, pw, ph;
  uint32_t tbw, tbh;
  uint32_t n_frames;
  // planar output: 3 planes per frame, each ph x pw floats
  float* planes;          // n x 3 x ph x pw (may be null)
  // stream output
  uint8_t* stream;        // n x frame_stream_bytes (may be null)
  uint64_t frame_stream_bytes;
  const uint32_t* block_types;  // n x mvw*mvh or null
  uint32_t mv_block_w, mv_block_h, mv_field_w, mv_field_h;
  float* scratch_planes;  // >= min(n, scratch_frames) x 3 x ph x pw, for the generic path
  uint32_t scratch_frames;
  // optional fused level-0 luma output (only honoured when dct_can_fuse_y(p)):
  // frame f of this launch -> pyramid slot y_first_slot + f
  uint8_t* y_l0;  // slot array base + level-0 offset, or null
  uint64_t y_slot_bytes;
  uint32_t y_first_slot, y_pitch;
};
bool dct_can_fuse_y(const DctParams& p);
// true when launch_dct(p) will need p.scratch_planes (generic transform / serializer path)
bool dct_needs_scratch(const DctParams& p);
cudaError_t launch_dct(const DctParams& p, cudaStream_t st, int* n_launches);
// once per device (opt-in shared memory sizes etc.)
cudaError_t prepare_dct_kernels();

// ---- decoder block path: dequantise + 8x8 IDCT + merge ------------------------------
struct DecodeParams {
  const uint8_t* records;        // n_frames x frame_record_bytes (records of 4 + 12 tb^2 bytes, raster order)
  uint64_t frame_record_bytes;
  uint32_t pw, ph;               // padded frame (multiples of the transform block)
  uint32_t tb;                   // square transform block: 8 (0 = 8), 16 or 4
  uint32_t n_frames;
  uint32_t fg_q, bg_q;
  uint32_t has_gaze, gaze_x, gaze_y, gaze_w, gaze_h;
  float* out;                    // n_frames x ph x pw x 3 interleaved BGR (16-byte aligned)
};
cudaError_t launch_decode(const DecodeParams& p, cudaStream_t st);
// exhaustive check of the decoder's division-free quotient against __fdiv_rn (see k_idct.cu)
cudaError_t run_dequant_selftest(uint32_t q_lo, uint32_t q_hi, unsigned long long* d_mismatches, cudaStream_t st);

}  // namespace svc
