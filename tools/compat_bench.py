#!/usr/bin/env python
"""Drop-in 1 of INTEGRATION.md: the stateless, host-buffer entry points called once per frame
(pyramids cross PCIe both ways) next to the reference's own CPU code on the same inputs."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

import numpy as np  # noqa: E402


def t_of(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    import svc_b200 as svc
    from oracle import oracle as O
    w, h = 1920, 1080
    pw, ph = svc.padded_dim(w, 16, 4), svc.padded_dim(h, 16, 4)
    fr = svc.SyntheticSequence(w, h, 2, seed=1234).frames()
    p0, p1 = O.y_pyramid(fr[0], pw, ph, 4), O.y_pyramid(fr[1], pw, ph, 4)
    res = {}
    res["svc_estimate_motion_hierarchical_16x16 ms"] = t_of(
        lambda: svc.EstimateMotionHierarchical16x16Sse2(p0, p1, pw, ph, 8), 20)
    if O.have_ref():
        res["reference EstimateMotionHierarchical16x16Sse2 ms (1 thread)"] = t_of(
            lambda: O.hbma(p0, p1, 8, impl="ref_sse2"), 5)
        res["reference EstimateMotionHierarchical ms (1 thread)"] = t_of(lambda: O.hbma(p0, p1, 8, impl="ref"), 2)
    res["svc_y_pyramid ms"] = t_of(lambda: svc.y_pyramid(fr[1], pw, ph, 4), 20)
    res["oracle y_pyramid (C port of the OpenCV calls) ms"] = t_of(lambda: O.y_pyramid(fr[1], pw, ph, 4), 3)
    res["svc_dct_planar ms"] = t_of(lambda: svc.dct_planar(fr[1], pw, ph), 10)
    res["oracle dct_planar (C port) ms"] = t_of(lambda: O.dct_planar(fr[1], pw, ph), 2)
    res["svc_encode_frame_stream ms"] = t_of(lambda: svc.encode_frame_stream(fr[1], pw, ph), 10)
    # "as shipped" comparators: the OpenCV calls the reference makes, one thread
    try:
        import cv2
        cv2.setNumThreads(1)
        bgr = fr[1]

        def cv_k1():
            padded = cv2.copyMakeBorder(bgr, 0, ph - h, 0, pw - w, cv2.BORDER_CONSTANT, value=0)  # libs/encoder.cpp:459-461
            y = cv2.extractChannel(cv2.cvtColor(padded, cv2.COLOR_BGR2YUV), 0)                     # :468-469
            pyr = [y]
            for _ in range(3):                                                                     # :470 (buildPyramid)
                pyr.append(cv2.pyrDown(pyr[-1]))
            return pyr

        def cv_dct():
            planes = cv2.split(cv2.copyMakeBorder(bgr, 0, ph - h, 0, pw - w, cv2.BORDER_CONSTANT, value=0)
                               .astype(np.float32))                                                # :638, :328
            for pl in planes:                                                                      # :329-337
                for by in range(0, ph, 8):
                    for bx in range(0, pw, 8):
                        pl[by:by + 8, bx:bx + 8] = cv2.dct(pl[by:by + 8, bx:bx + 8])
            return planes

        res["cv2 copyMakeBorder+cvtColor+extractChannel+3x pyrDown ms (1 thread)"] = t_of(cv_k1, 5)
        t0 = time.perf_counter()
        cv_dct()
        res["cv2.dct per 8x8 block, 97 920 calls ms (1 thread, python call overhead included)"] = \
            (time.perf_counter() - t0) * 1e3
        res["opencv"] = cv2.__version__
    except ImportError:
        pass
    res["host_cpus"] = os.cpu_count()
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "compat_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
