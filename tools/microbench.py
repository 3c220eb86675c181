#!/usr/bin/env python
"""BASELINE config 5: per-kernel microbenchmarks on 4K frames (HBM GB/s vs measured peak).

K1 (stand-alone BGR->Y + pyramid), K1b (pyramid levels only), K3 (+ fused luma), and the
decoder block kernel, each >= 100 launches over inputs far larger than L2, timed with CUDA
events on the launching stream; plus a coefficient check of K3 against the oracle on a crop.
"""
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
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--frames", type=int, default=16, help="frames per launch")
    ap.add_argument("--sets", type=int, default=8, help="distinct input sets cycled through (>> L2)")
    ap.add_argument("--iters", type=int, default=104)
    ap.add_argument("--tb", type=int, default=8, help="square transform block (8: default; 16 / 4: fused variants)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench_4k.json"))
    a = ap.parse_args()
    import torch
    import svc_b200 as svc
    from oracle import oracle as O

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    W, H, F, S = a.width, a.height, a.frames, a.sets
    torch.cuda.set_device(0)
    ts = torch.cuda.Stream()
    TB = a.tb
    sess = svc.Session(svc.SessionConfig(frame_w=W, frame_h=H, max_batch=F, cuda_stream=ts.cuda_stream,
                                         transform_block_w=TB, transform_block_h=TB))
    fin, fst = sess.frame_in_bytes, sess.frame_stream_bytes
    P = sess.padded_w * sess.padded_h
    seq = svc.SyntheticSequence(W, H, F, seed=99)
    base = torch.from_numpy(seq.frames().reshape(-1)).cuda()
    d_in = torch.empty(S * F * fin, dtype=torch.uint8, device="cuda")
    for s in range(S):  # distinct sets (rolled copies) so that no launch re-reads cached input
        d_in[s * F * fin:(s + 1) * F * fin] = torch.roll(base, shifts=3 * 64 * s)
    d_st = torch.empty(S * F * fst, dtype=torch.uint8, device="cuda")
    d_px = torch.empty(2 * F * P * 3, dtype=torch.float32, device="cuda")

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for i in range(a.iters):
            fn(i)
        e1.record(ts)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    res = {}
    ms = timed(lambda i: sess.run_stage(svc.STAGE_DCT_STREAM, d_in.data_ptr() + (i % S) * F * fin, F, None, None,
                                        d_st.data_ptr() + (i % S) * F * fst))
    b = F * (fin + fst + P)
    res["K3_dct_stream_luma"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak}
    ms = timed(lambda i: sess.run_stage(svc.STAGE_Y_PYRAMID, d_in.data_ptr() + (i % S) * F * fin, F))
    b = F * (fin + sum(P >> (2 * l) for l in range(4)))
    res["K1_bgr2y_pyramid_standalone"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak}
    ms = timed(lambda i: sess.run_stage(svc.STAGE_PYR_DOWN, None, F))
    b = F * (sum(P >> (2 * l) for l in range(3)) + sum(P >> (2 * l) for l in range(1, 4)))
    res["K1b_pyr_down_only"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak,
                                "note": "F frames = %.0f MB of pyramids: L2 resident between launches" % (b / 1e6)}
    if (W, H) == (sess.padded_w, sess.padded_h):
        ms = timed(lambda i: svc.decode_frames_device(0, ts.cuda_stream, d_st.data_ptr() + (i % S) * F * fst, F,
                                                      W, H, d_px.data_ptr() + (i % 2) * F * P * 3 * 4,
                                                      fg_quant_step=1, bg_quant_step=640, tb=TB))
        b = F * (fst + P * 12)
        res["decode_idct_blocks"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak}
    # coefficient tolerance on a crop of the last K3 output
    torch.cuda.synchronize()
    fr = seq.frame(0)
    st = d_st[:fst].cpu().numpy()
    nbx = W // TB
    rec = st.view(np.uint32).reshape(-1, nbx, 1 + 3 * TB * TB)
    crop = np.ascontiguousarray(fr[512:576, 1024:1152])
    planes = O.dct_planar(crop, 128, 64, TB, TB)
    by, bx = 64 // TB, 128 // TB
    got = rec[512 // TB:512 // TB + by, 1024 // TB:1024 // TB + bx, 1:].view(np.float32).reshape(by, bx, 3, TB, TB)
    exp = planes.reshape(3, by, TB, bx, TB).transpose(1, 3, 0, 2, 4)
    res["dct_max_abs_err_vs_oracle"] = float(np.abs(got - exp).max())
    out = {"gpu": torch.cuda.get_device_name(0), "width": W, "height": H, "frames_per_launch": F,
           "iters": a.iters, "transform_block": TB, "hbm_peak_gbs": peak, "results": res}
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
