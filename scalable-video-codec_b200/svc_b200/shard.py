"""Host-side frame-range sharding.

Motion for encoded frame t uses only the ORIGINAL input frames t-1 and t
(libs/encoder.cpp:472-476, pyramids ping-ponged at :661-663) and the DCT of
frame t only frame t (:638-640), so a sequence splits into contiguous ranges
of encoded frames; each shard is fed one extra input frame in front (its
tracked-only first frame).  No collective is needed: the host concatenates the
shard outputs in rank order behind the single 32-byte header.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_frame_ranges(n_input_frames: int, world: int) -> List[Tuple[int, int, int, int]]:
    """Split input frames 0..n-1 (encoded frames 1..n-1) over `world` ranks.

    Returns per rank (in_lo, in_hi, enc_lo, enc_hi): the rank reads input frames
    [in_lo, in_hi) and produces encoded frames [enc_lo, enc_hi) where encoded
    frame t is the one whose anchor is input frame t.  in_lo == enc_lo - 1 (the
    overlap frame).  Ranks with no work get empty ranges.
    """
    if world < 1:
        raise ValueError("world must be >= 1")
    n_enc = max(0, n_input_frames - 1)
    base, rem = divmod(n_enc, world)
    out = []
    t = 1
    for r in range(world):
        k = base + (1 if r < rem else 0)
        if k == 0:
            out.append((t - 1, t - 1, t, t))
        else:
            out.append((t - 1, t + k, t, t + k))
        t += k
    return out


def gather_streams(header: np.ndarray, shard_streams: Sequence[np.ndarray]) -> np.ndarray:
    """Header followed by the shards' records in rank (= frame) order."""
    parts = [np.asarray(header, dtype=np.uint8).ravel()]
    parts += [np.asarray(s, dtype=np.uint8).ravel() for s in shard_streams]
    return np.concatenate(parts)
