"""Generates tests/golden/segment.npz: fixtures for the block-type stages
(libs/encoder.cpp:491-624; product: scalable-video-codec_b200/host/segment.cpp).

Run in the authoring container (needs /root/reference compiled into oracle/_ref by
`make -C oracle`, and python cv2):

    python tests/golden/make_golden_segment.py

Sources of truth (the reference repository has no tests or golden vectors of its own):
  * RANSAC inliers / global motion / rmse: the UNMODIFIED reference
    EstimateGlobalMotionRansac (libs/motion.cpp:182-266) through oracle/_ref, its
    function-static engine seeded by interposing std::random_device (oracle/ref_shim.cpp);
  * morphology, k-means labels, connected components: python cv2 -- the calls the reference
    makes at libs/encoder.cpp:519-522, 574-575, 607-611 (cv2.setRNGSeed pins cv::theRNG());
  * block types of the whole chain: oracle.block_types_cv2 (cv2) on the reference's inliers.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def synth_mv_field(w, h, seed):
    """Integer-valued motion field: global pan, a few moving rectangles, sparse outliers."""
    rng = np.random.default_rng(seed)
    mv = np.zeros((h, w, 2), np.float32)
    mv[:] = rng.integers(-3, 4, size=2)
    for _ in range(int(rng.integers(2, 6))):
        x0, y0 = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
        x1, y1 = min(w, x0 + int(rng.integers(3, max(4, w // 3)))), min(h, y0 + int(rng.integers(3, max(4, h // 3))))
        mv[y0:y1, x0:x1] = rng.integers(-24, 25, size=2)
    noise = rng.random((h, w)) < 0.04
    mv[noise] += rng.integers(-16, 17, size=(int(noise.sum()), 2))
    jitter = rng.random((h, w)) < 0.2
    mv[jitter] += rng.integers(-1, 2, size=(int(jitter.sum()), 2))
    return mv


def main():
    out = {}
    rng = np.random.default_rng(99)
    # RANSAC: seeds x parameter sets, two consecutive calls each (engine state carries over)
    ransac_cases = []
    for ci, (w, h, seed, params) in enumerate([(120, 68, 11, (1, 7.5, 0.99, 0.5)), (60, 34, 77, (3, 2.0, 0.999, 0.4)),
                                               (20, 12, 2024, (2, 1.5, 0.9, 0.6))]):
        L = O.ref_seeded(seed)
        assert L is not None, "build oracle/_ref first"
        for call in range(2):
            mv = synth_mv_field(w, h, 100 * ci + call)
            rmse, gm, inl = O.ref_ransac(L, mv, *params)
            out[f"ransac{ci}_{call}_mv"] = mv
            out[f"ransac{ci}_{call}_params"] = np.array(params, np.float64)
            out[f"ransac{ci}_{call}_seed"] = np.array([seed], np.uint32)
            out[f"ransac{ci}_{call}_rmse"] = np.array([rmse], np.float32)
            out[f"ransac{ci}_{call}_gm"] = gm
            out[f"ransac{ci}_{call}_inliers"] = inl
        ransac_cases.append(ci)
    # morphology / connected components
    for i, (w, h, dens, rw, rh) in enumerate([(120, 68, 0.3, 3, 3), (31, 17, 0.6, 5, 3), (16, 9, 0.5, 2, 2), (9, 1, 0.5, 3, 3)]):
        m = ((rng.random((h, w)) < dens) * 255).astype(np.uint8)
        out[f"morph{i}_mask"] = m
        out[f"morph{i}_rect"] = np.array([rw, rh], np.uint32)
        el = cv2.getStructuringElement(cv2.MORPH_RECT, (rw, rh))
        for op in (0, 1, 2, 3):
            out[f"morph{i}_op{op}"] = cv2.morphologyEx(m, op, el)
        for conn in (4, 8):
            n, lab = cv2.connectedComponents(m, connectivity=conn, ltype=cv2.CV_32S)
            out[f"cc{i}_conn{conn}_n"] = np.array([n], np.uint32)
            out[f"cc{i}_conn{conn}_labels"] = lab
    # k-means
    for i, (n, k, max_iter, eps, attempts, seed) in enumerate([(700, 10, 10, 1.0, 3, 5), (40, 7, 4, 0.1, 2, 123456),
                                                               (5, 5, 10, 1.0, 3, 9), (1, 1, 10, 1.0, 3, 1)]):
        idx = np.sort(rng.choice(8160, size=n, replace=False))
        data = np.zeros((n, 4), np.float32)
        data[:, 1] = rng.integers(-8, 9, size=n)
        data[:, 2] = (idx % 120) * 16
        data[:, 3] = (idx // 120) * 16
        cv2.setRNGSeed(seed)
        comp, lab, cen = cv2.kmeans(data.reshape(n, 1, 4), k, None,
                                    (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, max_iter, eps), attempts,
                                    cv2.KMEANS_PP_CENTERS)
        out[f"kmeans{i}_data"] = data
        out[f"kmeans{i}_args"] = np.array([k, max_iter, eps, attempts, seed], np.float64)
        out[f"kmeans{i}_labels"] = lab.reshape(-1)
        out[f"kmeans{i}_centers"] = cen
        out[f"kmeans{i}_compactness"] = np.array([comp])
    # whole chain
    for i, (w, h, rseed, kseed, conn) in enumerate([(120, 68, 31, 7, 4), (60, 34, 5, 99, 8), (20, 12, 17, 3, 4)]):
        mv = synth_mv_field(w, h, 500 + i)
        L = O.ref_seeded(rseed)
        _, gm, inl = O.ref_ransac(L, mv)
        out[f"chain{i}_mv"] = mv
        out[f"chain{i}_seeds"] = np.array([rseed, kseed, conn], np.uint32)
        out[f"chain{i}_gm"] = gm
        out[f"chain{i}_types"] = O.block_types_cv2(mv, inl, kseed, connectivity=conn)
    np.savez_compressed(os.path.join(HERE, "segment.npz"), **out)
    print("wrote segment.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
