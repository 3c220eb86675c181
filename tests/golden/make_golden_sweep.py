"""Golden motion fields for the SAD-bound rows of the 1080p range / level sweep (BASELINE config 3).

    python tests/golden/make_golden_sweep.py        # needs oracle/_ref (the compiled reference)

The UNMODIFIED reference (EstimateMotionHierarchical / EstimateMotionExhaustiveSearch through
oracle/_ref/libref_motion.so, libs/motion.cpp:268-465) is run once on ONE 1080p frame pair of the
seeded synthetic sequence the sweep uses (tools/sweep_hbma.py: SyntheticSequence(1920, 1080, n,
seed=1234) with n = 45 = the sweep's --frames-small, frames 1 and 2) for the wide-range configurations that take the scalar reference tens
of seconds each; tests/test_gpu_parity.py and tools/sweep_hbma.py compare the CUDA kernels against
this file, so the wide-range rows of the sweep carry a parity bit without re-running the reference.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

from oracle import oracle as O  # noqa: E402
from svc_b200.synth import SyntheticSequence  # noqa: E402

CONFIGS = [(32, 1), (64, 1), (64, 2), (64, 3), (32, 2)]  # (search range R, pyramid levels L)
W, H, SEED, NFRAMES = 1920, 1080, 1234, 45  # the generator's later draws depend on the frame count


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
    seq = SyntheticSequence(W, H, NFRAMES, seed=SEED)
    f1, f2 = seq.frame(1), seq.frame(2)
    out = {"width": W, "height": H, "seed": SEED, "n_frames": NFRAMES, "pair": np.array([1, 2])}
    for R, L in CONFIGS:
        pw, ph = O.padded_dim(W, 16, L), O.padded_dim(H, 16, L)
        p0, p1 = O.y_pyramid(f1, pw, ph, L), O.y_pyramid(f2, pw, ph, L)
        t0 = time.perf_counter()
        mv, mad = O.hbma(p0, p1, R, impl="ref")
        print(f"R={R} L={L}: reference took {time.perf_counter() - t0:.1f} s", flush=True)
        assert np.all(mv == np.round(mv)) and np.abs(mv).max() <= 127 * 32
        out[f"mv_R{R}_L{L}"] = mv.astype(np.int16)  # integer valued (libs/motion.cpp:326-327)
        out[f"mad_R{R}_L{L}"] = mad
    np.savez_compressed(os.path.join(HERE, "sweep_1080p.npz"), **out)


if __name__ == "__main__":
    main()
