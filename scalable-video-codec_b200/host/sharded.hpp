// sharded.hpp -- frame-range sharded encode of a raw BGR file over several GPUs of one
// box: one svc_session + one host thread per GPU, contiguous ranges of
// encoded frames with one overlap input frame, no collective; every shard writes its
// records straight to their final offset of the output file (pwrite), so the "host
// gather" is the file system.  Motion for frame t only needs input frames t-1 and t
// (libs/encoder.cpp:472-476, 661-663), the DCT only frame t (:638-640).
#ifndef SVC_B200_HOST_SHARDED_HPP
#define SVC_B200_HOST_SHARDED_HPP

#include <array>
#include <string>
#include <vector>

#include "encoder.hpp"

namespace svc {

struct ShardRange {  // input frames [in_lo, in_hi), encoded frames [enc_lo, enc_hi)
  uint in_lo, in_hi, enc_lo, enc_hi;
};

// Same partition as svc_b200/shard.py:shard_frame_ranges.
std::vector<ShardRange> ShardFrameRanges(uint n_input_frames, uint world);

struct ShardedStats {
  uint64_t frames_encoded = 0;
  double seconds = 0;        // wall clock of the parallel section (set-up included)
  double setup_seconds = 0;  // slowest shard: CUDA context, session, pinned staging buffers
  double read_seconds = 0;   // slowest shard: reading input frames
  // the next two run on a post thread, overlapped with the following batch's read + encode
  double encode_seconds = 0; // slowest shard: svc_session_encode (H2D | kernels | D2H) + block-type stages
  double write_seconds = 0;  // slowest shard: writing the records
};

// Encodes `in_path` (vidprops.frame_count raw BGR frames) into `out_path` (header +
// records, the reference stream layout) using one session per entry of `devices`
// (a device may be listed more than once).  Throws svc::Error on failure.
ShardedStats EncodeFileSharded(const EncoderConfig& cfg, const VideoProperties& vidprops,
                               const std::string& in_path, const std::string& out_path,
                               const std::vector<int>& devices, BlockTypeFn classify = nullptr);

}  // namespace svc
#endif  // SVC_B200_HOST_SHARDED_HPP
