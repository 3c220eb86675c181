"""svc_b200 -- python (ctypes) host binding of libsvc_b200.so.

The product is the C-ABI shared library (include/svc_b200.h) and the sm_100a
kernels behind it; this package only marshals numpy / raw device pointers
across that boundary for the tests and the benchmark, mirrors the reference's
function names (libs/motion.hpp, libs/encoder.cpp) and holds the host-side
frame-range sharding logic.  There is no CPU implementation here: importing
works without a GPU, every compute call raises SvcError without one.
"""
from .binding import (  # noqa: F401
    SvcError, lib, lib_path, device_count, sad_peak, selftest_dequant, padded_dim, serialized_frame_bytes,
    write_header, EstimateMotionHierarchical, EstimateMotionHierarchical16x16Sse2,
    EstimateMotionExhaustiveSearch, y_pyramid, dct_planar, encode_frame_stream,
    patch_block_types, stream_layout, gaze_rect, decode_frame_blocks, decode_frames_device, Session, SessionConfig, PinnedBuffer, DeviceBuffer,
    STAGE_Y_PYRAMID, STAGE_HBMA, STAGE_DCT_STREAM, STAGE_PYR_DOWN,
    HBMA_FAMILY_AUTO, HBMA_FAMILY_GENERIC, HBMA_FAMILY_POOL, HBMA_FAMILY_WINDOW, HBMA_FAMILY_TILE,
)
from .shard import shard_frame_ranges, gather_streams  # noqa: F401
from .synth import SyntheticSequence  # noqa: F401
