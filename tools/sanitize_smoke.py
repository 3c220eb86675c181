#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitize_smoke.py

Frames are tiny (the tools slow kernels down 10-100x); results are still compared with the oracle so
that a finding can be told from a wrong answer.  SURVEY section 5: the reference has no sanitizer coverage."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

import numpy as np  # noqa: E402


def main():
    import svc_b200 as svc
    from oracle import oracle as O
    ok = True
    w, h, n = 208, 112, 3
    frames = svc.SyntheticSequence(w, h, n, seed=11).frames()
    # (levels, range): default strip kernels (and the tile kernel they replaced, through the hook), rs kernels at r = 8 / 16 / 4 (hybrid), shared-window / pooled /
    # striped kernels, the generic kernel through the test hook
    for L, R, fam in ((4, 8, 0), (2, 16, 0), (4, 64, 0), (2, 32, 0), (1, 8, 0), (5, 64, 0), (1, 16, 0), (2, 64, 0),
                      (1, 40, 0), (3, 8, 0), (4, 8, svc.HBMA_FAMILY_GENERIC), (4, 8, svc.HBMA_FAMILY_TILE), (2, 16, svc.HBMA_FAMILY_POOL),
                      (2, 16, svc.HBMA_FAMILY_WINDOW)):
        with svc.Session(svc.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L,
                                           hbma_kernel_family=fam, max_batch=2)) as s:
            mv, mad, st = s.encode(frames)
            pw, ph = s.padded_w, s.padded_h
        pyr = [O.y_pyramid(f, pw, ph, L) for f in frames]
        for i in range(1, n):
            emv, emad = O.hbma(pyr[i - 1], pyr[i], R)
            good = np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad)
            ok &= good
            if not good:
                print("MISMATCH motion", L, R, fam, i)
        exp = O.serialize_frame(O.dct_planar(frames[1], pw, ph), None, w, h, 8, 8, pw // 16, 16, 16)
        err = np.abs(st[0].view(np.float32) - exp.view(np.float32)).max()
        ok &= bool(err <= 1e-3)
        if not err <= 1e-3:
            print("MISMATCH records", L, R, fam, float(err))
    for tb in (16, 4, 8):  # fused stream kernels of the other transform blocks + the decoder kernels
        with svc.Session(svc.SessionConfig(frame_w=w, frame_h=h, transform_block_w=tb, transform_block_h=tb, max_batch=3)) as s:
            _, _, st = s.encode(frames)
            pw, ph = s.padded_w, s.padded_h
        if ph == h:
            out = svc.decode_frame_blocks(st[0], pw, ph, 1, 24, None, tb, tb)
            exp = O.decode_frame_blocks(st[0], pw, ph, tb, tb, 1, 24, None)
            derr = float(np.abs(out.astype(np.float64) - exp.astype(np.float64)).max())
            ok &= derr <= 1e-3  # float pixels, same tolerance as tests/test_decode.py
            if not derr <= 1e-3:
                print("MISMATCH decode", tb, derr)
    ok &= svc.selftest_dequant(640, 640) == 0 if os.environ.get("SVC_SANITIZE_SELFTEST") else True
    print("sanitize_smoke:", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
