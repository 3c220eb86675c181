#!/usr/bin/env python
"""A/B timing of the search kernels at the encoder default (16x16, R=8, L=4) on ONE box: the strip
kernel (k_hbma_strip.cu, what the dispatcher picks) against the bounded-reach tile kernel it replaced
(session test hook HBMA_FAMILY_TILE).  Same pyramids, launches alternate, CUDA events on the session
stream, outputs compared bit for bit.  Writes gpurun_out/ab_hbma.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=100, help="frame pairs per launch")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_hbma.json"))
    a = ap.parse_args()
    import torch
    import svc_b200 as svc

    torch.cuda.set_device(0)
    ts = torch.cuda.Stream()
    W, H, F = a.width, a.height, a.frames
    base = svc.SyntheticSequence(W, H, 12, seed=77).frames()
    d_in = torch.from_numpy(np.concatenate([base] * ((F + 12) // 12))[:F].reshape(-1).copy()).cuda()
    res = {}
    outs = {}
    sess = {}
    for name, fam in (("strip", svc.HBMA_FAMILY_AUTO), ("tile", svc.HBMA_FAMILY_TILE)):
        s = svc.Session(svc.SessionConfig(frame_w=W, frame_h=H, max_batch=F, hbma_kernel_family=fam,
                                          cuda_stream=ts.cuda_stream))
        n = s.mv_field_w * s.mv_field_h
        mv = torch.zeros(F * n * 2, dtype=torch.float32, device="cuda")
        mad = torch.zeros(F * n, dtype=torch.float32, device="cuda")
        s.run_stage(svc.STAGE_Y_PYRAMID, d_in.data_ptr(), F)
        s.run_stage(svc.STAGE_HBMA, None, F, mv.data_ptr(), mad.data_ptr())
        torch.cuda.synchronize()
        sess[name] = (s, mv, mad)
        outs[name] = (mv.cpu().numpy().copy(), mad.cpu().numpy().copy())
        res[name] = []
    for _ in range(a.reps):
        for name in sess:
            s, mv, mad = sess[name]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            s.run_stage(svc.STAGE_HBMA, None, F, mv.data_ptr(), mad.data_ptr())
            e1.record(ts)
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1))
    out = {"frames_per_launch": F, "width": W, "height": H,
           "identical_outputs": bool(np.array_equal(outs["strip"][0], outs["tile"][0]) and
                                     np.array_equal(outs["strip"][1], outs["tile"][1]))}
    for name in res:
        t = np.array(res[name][3:] if len(res[name]) > 3 else res[name])
        out[name] = {"ms_median": float(np.median(t)), "ms_min": float(t.min()),
                     "us_per_frame": float(np.median(t)) * 1e3 / F}
    out["speedup"] = out["tile"]["ms_median"] / out["strip"]["ms_median"]
    for s, _, _ in sess.values():
        s.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
