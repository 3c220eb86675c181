/*
 * svc_b200.h -- C ABI of libsvc_b200.so: the B200-native (sm_100a) encoder hot
 * path of the scalable video codec (Y pyramid, hierarchical block-matching
 * motion estimation, block DCT, stream layout).
 *
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference repository, fonzcastellanos/scalable-video-codec).  All
 * pointers are plain host pointers unless the name says "device"; the caller
 * owns every buffer, exactly as in the reference (libs/encoder.cpp:174-219,
 * 353-356).  There is no CPU fallback: without a CUDA device every compute
 * entry point returns SVC_ERR_CUDA.
 *
 * Return value: 0 (SVC_OK) on success, otherwise an SVC_ERR_* code;
 * svc_last_error() returns a thread-local human readable message.  The
 * reference motion functions return void and assert() their preconditions
 * (libs/motion.cpp:417-433, 701-712); here a violated precondition is
 * SVC_ERR_INVALID_ARG.
 */
#ifndef SVC_B200_H
#define SVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVC_OK 0
#define SVC_ERR_INVALID_ARG 1 /* violated precondition (reference: assert / Validate) */
#define SVC_ERR_CUDA 2        /* CUDA runtime/driver failure, or no device */
#define SVC_ERR_UNSUPPORTED 3 /* legal for the reference, outside this build's limits */
#define SVC_ERR_STATE 4       /* session misuse */

#define SVC_MAX_LEVELS 8

const char* svc_last_error(void);
const char* svc_version(void);
/* Number of visible CUDA devices (0 when there is none). */
int svc_device_count(int* count);

/* ------------------------------------------------------------------------
 * Geometry helpers (pure host arithmetic)
 * ---------------------------------------------------------------------- */

/* ClosestLargerDivisible(a, block, 2^(levels-1)) -- libs/math.hpp:276-283 as
 * used by Encoder::Encoder, libs/encoder.cpp:165-169. */
uint32_t svc_padded_dim(uint32_t a, uint32_t mv_block, uint32_t levels);

/* Bytes of one serialised frame: ceil(w/tbw) * ceil(h/tbh) records of
 * 4 + channels*tbw*tbh*4 bytes (libs/encoder.cpp:243-266). */
uint64_t svc_serialized_frame_bytes(uint32_t frame_w, uint32_t frame_h,
                                    uint32_t tbw, uint32_t tbh,
                                    uint32_t channels);

/* The 32-byte stream header, libs/codec.hpp:8-17, emitted at
 * libs/encoder.cpp:361-381 (frame_count = n_input_frames - 1). */
int svc_write_header(uint32_t n_input_frames, uint32_t frame_w,
                     uint32_t frame_h, uint32_t padded_w, uint32_t padded_h,
                     uint32_t tbw, uint32_t tbh, uint32_t channels,
                     uint8_t out32[32]);

/* ------------------------------------------------------------------------
 * Stateless drop-ins (host buffers in, host buffers out; default device 0,
 * see svc_set_device).  These bind 1:1 to the reference free functions.
 * ---------------------------------------------------------------------- */

/* Device used by the stateless entry points of the calling thread. */
int svc_set_device(int device);

/* EstimateMotionHierarchical -- libs/motion.hpp:134-138, libs/motion.cpp:412-465.
 * tracked_pyr/anchor_pyr: level_count pointers to tightly packed 8-bit planes,
 * level l is (frame_w >> l) x (frame_h >> l).  motion_field_xy: {x,y} float
 * pairs (the Vec2f ABI, libs/math.hpp:177-181), row-major
 * (frame_w/block_w) x (frame_h/block_h); min_mad likewise.
 * Preconditions: dims > 0, frame_w % block_w == frame_h % block_h == 0,
 * search_range >= 2^(level_count-1), block_w and block_h divisible by
 * 2^(level_count-1) (the reference divides by zero otherwise). */
int svc_estimate_motion_hierarchical(const uint8_t* const* tracked_pyr,
                                     const uint8_t* const* anchor_pyr,
                                     uint32_t level_count, uint32_t frame_w,
                                     uint32_t frame_h, uint32_t search_range,
                                     uint32_t block_w, uint32_t block_h,
                                     float* motion_field_xy, float* min_mad);

/* EstimateMotionHierarchical16x16Sse2 -- libs/motion.hpp:148-152,
 * libs/motion.cpp:691-749 (4 levels, 16x16 blocks). */
int svc_estimate_motion_hierarchical_16x16(const uint8_t* const* tracked_pyr,
                                           const uint8_t* const* anchor_pyr,
                                           uint32_t frame_w, uint32_t frame_h,
                                           uint32_t search_range,
                                           float* mv_field_xy, float* min_mad);

/* EstimateMotionExhaustiveSearch -- libs/motion.hpp:106-110,
 * libs/motion.cpp:268-340. */
int svc_estimate_motion_exhaustive(const uint8_t* tracked_frame,
                                   const uint8_t* anchor_frame,
                                   uint32_t frame_w, uint32_t frame_h,
                                   uint32_t search_range, uint32_t block_w,
                                   uint32_t block_h, float* motion_field_xy,
                                   float* min_mad);

/* copyMakeBorder + cvtColor(BGR2YUV) + extractChannel(0) + buildPyramid --
 * libs/encoder.cpp:459-470 (first frame :447-451).  bgr: frame_h x frame_w x 3
 * interleaved; out_levels[l]: tightly packed (padded_w >> l) x (padded_h >> l).
 * padded dims must be >= frame dims and divisible by 2^(level_count-1). */
int svc_y_pyramid(const uint8_t* bgr, uint32_t frame_w, uint32_t frame_h,
                  uint32_t padded_w, uint32_t padded_h, uint32_t level_count,
                  uint8_t* const* out_levels);

/* convertTo(CV_32FC3) + Dct -- libs/encoder.cpp:638-640, 323-339.
 * planes[c], c = B,G,R: padded_h x padded_w float, every tbh x tbw block
 * replaced by its orthonormal 2-D DCT-II (zero padding on bottom/right). */
int svc_dct_planar(const uint8_t* bgr, uint32_t frame_w, uint32_t frame_h,
                   uint32_t padded_w, uint32_t padded_h, uint32_t tbw,
                   uint32_t tbh, float* const* planes);

/* Dct + SerializeEncodedFrame -- libs/encoder.cpp:638-650, 222-269: one
 * frame's records written to `out` (svc_serialized_frame_bytes bytes).
 * block_types: mv_field_w*mv_field_h u32 (NULL = all BLOCK_TYPE_BACKGROUND). */
int svc_encode_frame_stream(const uint8_t* bgr, uint32_t frame_w,
                            uint32_t frame_h, uint32_t padded_w,
                            uint32_t padded_h, uint32_t tbw, uint32_t tbh,
                            uint32_t mv_block_w, uint32_t mv_block_h,
                            const uint32_t* block_types, uint8_t* out);

/* Overwrite the 4-byte block-type field of every record of one serialised
 * frame (host memory) -- lets the CPU stages that follow motion estimation
 * (RANSAC .. connected components, libs/encoder.cpp:491-624) label a stream
 * the GPU has already written.  Pure host byte patching. */
int svc_patch_block_types(uint8_t* frame_stream, uint32_t frame_w,
                          uint32_t frame_h, uint32_t tbw, uint32_t tbh,
                          uint32_t channels, uint32_t mv_block_w,
                          uint32_t mv_block_h, uint32_t mv_field_w,
                          const uint32_t* block_types);

/* ------------------------------------------------------------------------
 * Session: the device-resident fast path.  One session = one Encoder
 * (libs/encoder.hpp:52-95) on one GPU; it keeps the previous frame's Y pyramid
 * resident and ping-pongs it like libs/encoder.cpp:661-663.  Single caller
 * thread per session; sessions on different GPUs may run concurrently.
 * ---------------------------------------------------------------------- */

typedef struct svc_session svc_session;

typedef struct svc_session_config {
  uint32_t struct_size; /* = sizeof(svc_session_config) */
  /* VideoProperties, libs/encoder.hpp:46-50 */
  uint32_t frame_w, frame_h;
  /* EncoderConfig hot-path fields, libs/encoder.hpp:25-37
   * (defaults apps/encoder.cpp:42-58: 16,16,8,4,8,8) */
  uint32_t mv_block_w, mv_block_h;
  uint32_t mv_search_range;
  uint32_t pyr_lvl_count;
  uint32_t transform_block_w, transform_block_h;
  int32_t device;     /* CUDA ordinal */
  uint32_t max_batch; /* frames per launch batch (0 = default 32) */
  void* cuda_stream;  /* cudaStream_t to run on; NULL = session-owned stream */
  /* Test hook (the only one; nothing in the library reads the environment): restrict the motion-search
   * dispatcher to one kernel family so that parity tests can reach kernels the dispatcher would not
   * pick for a configuration.  SVC_HBMA_FAMILY_AUTO (0) = fastest kernel for the configuration.
   * A caller compiled against the older struct (struct_size without this field) gets AUTO. */
  uint32_t hbma_kernel_family;
  /* Host path (svc_session_encode): frames per pipeline stage of H2D | kernels | D2H (0 = default 16).
   * PCIe-bound: small stages keep the three engines overlapped and the exposed head / tail short. */
  uint32_t host_chunk_frames;
} svc_session_config;

#define SVC_HBMA_FAMILY_AUTO 0u
#define SVC_HBMA_FAMILY_GENERIC 1u /* hbma_generic_kernel: any block shape / level count / range */
#define SVC_HBMA_FAMILY_POOL 2u    /* single-launch pooled kernel (16x16 blocks, r = 5..64) */
#define SVC_HBMA_FAMILY_WINDOW 3u  /* per-block TMA window kernels (16x16 blocks) */
#define SVC_HBMA_FAMILY_TILE 4u    /* bounded-reach tile kernel also where the strip kernel would be picked */

typedef struct svc_session_info {
  uint32_t padded_w, padded_h;       /* libs/encoder.cpp:166-169 */
  uint32_t mv_field_w, mv_field_h;   /* libs/encoder.cpp:174-175 */
  uint64_t frame_in_bytes;           /* frame_w*frame_h*3 */
  uint64_t frame_stream_bytes;       /* one serialised frame */
  uint32_t record_bytes;             /* 4 + 3*tbw*tbh*4 */
  uint32_t max_batch;
} svc_session_info;

/* Validates like Validate(EncoderConfig) (libs/encoder.cpp:62-142) for the
 * hot-path fields plus the preconditions listed above. */
int svc_session_create(const svc_session_config* cfg, svc_session** out);
void svc_session_destroy(svc_session* s);
int svc_session_info_get(const svc_session* s, svc_session_info* out);
/* Forget the previous frame (the next frame pushed is a tracked-only first
 * frame again, libs/encoder.cpp:447-451). */
int svc_session_reset(svc_session* s);

/* Push n_frames consecutive input frames (host memory, frame_h x frame_w x 3
 * BGR each).  Encoded frames produced: n_frames - 1 if the session has no
 * previous frame, else n_frames (written to *n_encoded).  Per encoded frame:
 *   mv_xy   mv_field_h*mv_field_w*2 floats   (NULL = not wanted)
 *   min_mad mv_field_h*mv_field_w floats     (NULL = not wanted)
 *   stream  frame_stream_bytes bytes, block types 0 or `block_types`
 *           (n_encoded * mv_field_w*mv_field_h u32, may be NULL)
 * Host buffers from svc_host_alloc are pinned and let the copies overlap the
 * kernels; any other host memory works too.  Blocking. */
int svc_session_encode(svc_session* s, const uint8_t* frames_bgr,
                       uint32_t n_frames, float* mv_xy, float* min_mad,
                       uint8_t* stream, const uint32_t* block_types,
                       uint32_t* n_encoded);

/* Same, with every buffer in the session GPU's memory; asynchronous on the
 * session stream (call svc_session_synchronize, or synchronise the stream
 * handed in through cuda_stream). */
int svc_session_encode_device(svc_session* s, const uint8_t* d_frames_bgr,
                              uint32_t n_frames, float* d_mv_xy,
                              float* d_min_mad, uint8_t* d_stream,
                              const uint32_t* d_block_types,
                              uint32_t* n_encoded);
int svc_session_synchronize(svc_session* s);

/* Device-side stage entry points of a session (device pointers, async on the
 * session stream); used for per-kernel measurement and parity tests.
 * Stage ids for svc_session_run_stage. */
#define SVC_STAGE_Y_PYRAMID 1 /* K1: frames -> pyramid slots 1..n */
#define SVC_STAGE_HBMA 2      /* K2: slots i,i+1 -> mv/mad of frame i */
#define SVC_STAGE_DCT_STREAM 3 /* K3: frames -> stream records (+ level-0 luma of slots 1..n
                                  when the fused path applies: square 8x8, 16x16 or 4x4 blocks, W == padded W) */
#define SVC_STAGE_PYR_DOWN 4   /* K1b: level 0 of slots 1..n -> levels 1.. */
int svc_session_run_stage(svc_session* s, int stage,
                          const uint8_t* d_frames_bgr, uint32_t n_frames,
                          float* d_mv_xy, float* d_min_mad, uint8_t* d_stream);
/* Exact SAD work of K2 on pyramid slots 0..n_frames (the pairs run_stage(HBMA) would
 * process): candidate evaluations and byte-absdiffs, border clamping included, counted
 * on the device while searching.  Blocking. */
int svc_session_hbma_work(svc_session* s, uint32_t n_frames, uint64_t* candidates,
                          uint64_t* absdiffs);
/* Measured packed-byte SAD issue peak of the device (byte-absdiffs per second of a
 * dependency-free VABSDIFF4.ACC loop): the integer roofline of the range sweep. */
int svc_sad_peak(int device, double* absdiffs_per_s);
/* Number of kernel launches issued by this session so far. */
uint64_t svc_session_launch_count(const svc_session* s);

/* ------------------------------------------------------------------------
 * Decoder block path (the inverse of K3)
 * ---------------------------------------------------------------------- */
typedef struct svc_rect { uint32_t x, y, w, h; } svc_rect;

/* CalcWithinFrameRectFromCenter + the scaling of the gaze rectangle to the padded
 * frame -- libs/decoder.cpp:66-98, 172-189 (pure host arithmetic). */
int svc_gaze_rect(uint32_t gaze_x, uint32_t gaze_y, uint32_t max_gaze_rect_w,
                  uint32_t max_gaze_rect_h, uint32_t frame_w, uint32_t frame_h,
                  uint32_t padded_w, uint32_t padded_h, svc_rect* out);

/* ParseBlock + DecodeBlock for every record of one frame -- libs/decoder.cpp:102-149 over
 * the loop at :191-213.  frame_records: (padded_w/tbw)*(padded_h/tbh) records in raster
 * order; quantisation step 1 for blocks whose top-left corner is inside `gaze` (may be
 * NULL), background step for block type 0, foreground step otherwise (DecoderConfig,
 * libs/decoder.hpp:12-17; defaults 1 / 640, apps/decoder.cpp:20-25).  out_bgr: padded_h x
 * padded_w x 3 interleaved float -- the `upscaled_frame` before the division by 255 and
 * the resize.  Square 8x8, 16x16 and 4x4 transform blocks. */
int svc_decode_frame_blocks(const uint8_t* frame_records, uint32_t padded_w, uint32_t padded_h,
                            uint32_t tbw, uint32_t tbh, uint32_t fg_quant_step,
                            uint32_t bg_quant_step, const svc_rect* gaze, float* out_bgr);
/* Same for n_frames frames resident on `device`; asynchronous on `cuda_stream`. */
int svc_decode_frames_device(int device, void* cuda_stream, const uint8_t* d_records,
                             uint32_t n_frames, uint32_t padded_w, uint32_t padded_h,
                             uint32_t tbw, uint32_t tbh, uint32_t fg_quant_step,
                             uint32_t bg_quant_step, const svc_rect* gaze, float* d_out_bgr);

/* Self-test of the decoder's dequantiser (round(c / q) * q, libs/decoder.cpp:137-144): the kernels
 * replace the IEEE division by a reciprocal + two fused multiply-adds for steps <= 4096 and
 * |c| <= 2^18.  Compares that quotient with the IEEE division for EVERY float of that range (both
 * signs) and every step in [q_lo, q_hi]; *mismatches must come back 0. */
int svc_selftest_dequant(int device, uint32_t q_lo, uint32_t q_hi, uint64_t* mismatches);

/* Stream validator: record geometry implied by a 32-byte header on the encoder side
 * (SerializeEncodedFrame iterates the UNPADDED frame, libs/encoder.cpp:243-244) and on
 * the decoder side (Decoder::operator() iterates the PADDED frame, libs/decoder.cpp:
 * 191-192; the reader thread likewise, apps/decoder.cpp:66-71).  The two agree only when
 * ceil(h/tbh) == padded_h/tbh and ceil(w/tbw) == padded_w/tbw -- true for 960x540 and
 * 3840x2160, false for 1920x1080 (135 vs 136 block rows): a reference inconsistency that
 * this library preserves and reports instead of hiding. */
typedef struct svc_stream_layout {
  uint32_t frame_count, frame_w, frame_h, padded_w, padded_h, tbw, tbh, channels;
  uint32_t record_bytes;
  uint64_t encoder_records_per_frame; /* what the encoder writes */
  uint64_t decoder_records_per_frame; /* what the reference decoder will try to read */
  uint64_t encoder_stream_bytes;      /* 32 + frame_count * encoder_records * record_bytes */
  int32_t consistent;                 /* 1 when both conventions agree */
} svc_stream_layout;
int svc_stream_layout_from_header(const uint8_t header32[32], svc_stream_layout* out);

/* Memory helpers so that non-CUDA hosts (ctypes, cgo, JNI) need no runtime. */
void* svc_host_alloc(size_t bytes); /* pinned */
void svc_host_free(void* p);
void* svc_device_alloc(int device, size_t bytes);
void svc_device_free(int device, void* p);
int svc_memcpy_h2d(int device, void* dst_device, const void* src_host, size_t bytes);
int svc_memcpy_d2h(int device, void* dst_host, const void* src_device, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SVC_B200_H */
