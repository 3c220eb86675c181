#!/usr/bin/env python
"""BASELINE config 3: HBMA search-range / pyramid-level sweep at 1080p on one B200.

For every (R, L) with R >= 2^(L-1): K2 time per frame (CUDA events on the session
stream), exact candidate / byte-absdiff counts (counted on the device), Gcand/s,
G absdiff/s and the fraction of the measured VABSDIFF4 peak; next to it the
reference's own CPU code (oracle/_ref: SSE2 entry for L=4, generic otherwise) on
one frame pair where that finishes in reasonable time.  Writes
gpurun_out/sweep_hbma.json and a markdown table.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=9, help="frame pairs per launch for wide ranges")
    ap.add_argument("--frames-small", type=int, default=45,
                    help="frame pairs per launch when the top-level range is <= 8 (a 9-pair launch of such a "
                         "configuration lasts 20-300 us: launch latency and the last wave would dominate)")
    ap.add_argument("--ranges", default="8,16,32,64")
    ap.add_argument("--levels", default="1,2,3,4,5")
    ap.add_argument("--cpu-budget-gabsdiff", type=float, default=4.0,
                    help="run the CPU reference only when one frame pair is below this many G absdiffs")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_hbma"))
    a = ap.parse_args()

    import torch
    import svc_b200 as svc
    from oracle import oracle as O

    torch.cuda.set_device(0)
    ts = torch.cuda.Stream()
    peak = svc.sad_peak(0)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6554.2  # measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json when present)
    W, H = a.width, a.height
    F_max = max(a.frames, a.frames_small)
    frames = svc.SyntheticSequence(W, H, F_max, seed=1234).frames()
    d_in = torch.from_numpy(frames.reshape(-1)).cuda()
    gpath = os.path.join(ROOT, "tests", "golden", "sweep_1080p.npz")
    golden = np.load(gpath) if os.path.exists(gpath) else None
    rows = []
    for L in [int(x) for x in a.levels.split(",")]:
        for R in [int(x) for x in a.ranges.split(",")]:
            if R < (1 << (L - 1)):
                continue
            F = a.frames_small if (R >> (L - 1)) <= 8 else a.frames
            sess = svc.Session(svc.SessionConfig(frame_w=W, frame_h=H, mv_search_range=R, pyr_lvl_count=L,
                                                 max_batch=F, cuda_stream=ts.cuda_stream))
            mvn = sess.mv_field_w * sess.mv_field_h
            d_mv = torch.empty(F * mvn * 2, dtype=torch.float32, device="cuda")
            d_mad = torch.empty(F * mvn, dtype=torch.float32, device="cuda")
            sess.run_stage(svc.STAGE_Y_PYRAMID, d_in.data_ptr(), F)  # slots 1..F; pairs (i, i+1), i = 1..F-1
            n_pairs = F  # K2 on slots 0..F: pair 0 has the zero slot as tracked frame
            sess.run_stage(svc.STAGE_HBMA, None, n_pairs, d_mv.data_ptr(), d_mad.data_ptr())  # warm-up
            torch.cuda.synchronize()
            reps = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for _ in range(reps):
                sess.run_stage(svc.STAGE_HBMA, None, n_pairs, d_mv.data_ptr(), d_mad.data_ptr())
            e1.record(ts)
            torch.cuda.synchronize()
            ms_frame = e0.elapsed_time(e1) / reps / n_pairs
            cand, absd = sess.hbma_work(n_pairs)
            P = sess.padded_w * sess.padded_h
            hbm_bytes = 2 * sum(P >> (2 * l) for l in range(L)) + mvn * 12  # both pyramids once + vectors, MADs
            row = {"R": R, "L": L, "r_top": R >> (L - 1), "pairs_per_launch": F, "ms_per_frame": ms_frame,
                   "cand_per_frame": cand / n_pairs, "absdiff_per_frame": absd / n_pairs,
                   "gcand_per_s": cand / n_pairs / ms_frame / 1e6,
                   "gabsdiff_per_s": absd / n_pairs / ms_frame / 1e6,
                   "frac_of_sad_peak": absd / n_pairs / (ms_frame * 1e-3) / peak,
                   "frac_of_hbm_peak": hbm_bytes / (ms_frame * 1e-3) / (hbm_peak * 1e9)}
            # CPU reference on one pair
            if absd / n_pairs / 1e9 <= a.cpu_budget_gabsdiff and O.have_ref():
                pw, ph = sess.padded_w, sess.padded_h
                p0, p1 = O.y_pyramid(frames[1], pw, ph, L), O.y_pyramid(frames[2], pw, ph, L)
                impl = "ref_sse2" if L == 4 else "ref"
                t0 = time.perf_counter()
                rmv, rmad = O.hbma(p0, p1, R, impl=impl)
                row["cpu_ref_ms_per_frame"] = (time.perf_counter() - t0) * 1e3
                row["cpu_ref_impl"] = impl
                mv = d_mv.cpu().numpy().reshape(F, sess.mv_field_h, sess.mv_field_w, 2)
                mad = d_mad.cpu().numpy().reshape(F, sess.mv_field_h, sess.mv_field_w)
                row["bit_exact_vs_reference"] = bool(np.array_equal(mv[2], rmv) and np.array_equal(mad[2], rmad))
                row["parity_source"] = "reference run live (oracle/_ref)"
            elif (golden is not None and f"mv_R{R}_L{L}" in golden and
                  (W, H, F_max) == (int(golden["width"]), int(golden["height"]), int(golden["n_frames"]))):
                # wide ranges: the scalar reference needs tens of seconds per pair; its output for this very
                # pair (frames 1, 2 of the seed-1234 sequence) is committed in tests/golden/sweep_1080p.npz
                mv = d_mv.cpu().numpy().reshape(F, sess.mv_field_h, sess.mv_field_w, 2)
                mad = d_mad.cpu().numpy().reshape(F, sess.mv_field_h, sess.mv_field_w)
                row["bit_exact_vs_reference"] = bool(
                    np.array_equal(mv[2], golden[f"mv_R{R}_L{L}"].astype(np.float32)) and
                    np.array_equal(mad[2], golden[f"mad_R{R}_L{L}"]))
                row["parity_source"] = "golden of the unmodified reference (tests/golden/sweep_1080p.npz)"
            rows.append(row)
            print(json.dumps(row), flush=True)
            sess.close()
    out = {"gpu": torch.cuda.get_device_name(0), "width": W, "height": H,
           "pairs_per_launch": {"r_top<=8": a.frames_small, "r_top>8": a.frames},
           "sad_peak_gabsdiff_per_s": peak / 1e9, "rows": rows}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".md", "w") as f:
        f.write(f"# HBMA range/level sweep, {W}x{H}, {torch.cuda.get_device_name(0)}\n\n")
        f.write(f"Measured VABSDIFF4 peak: {peak / 1e12:.2f} T byte-absdiff/s\n\n")
        f.write(f"HBM peak: {hbm_peak:.0f} GB/s; the binding roofline is the larger of the two percentages "
                "(small windows are bound by pyramid traffic and per-level latency, not by SADs)\n\n")
        f.write("| R | L | r | pairs/launch | ms/frame | Mcand/frame | Gabsdiff/frame | Gcand/s | Gabsdiff/s | % SAD peak | % HBM peak | CPU ref ms/frame | bit-exact |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['R']} | {r['L']} | {r['r_top']} | {r['pairs_per_launch']} | {r['ms_per_frame']:.4f} | {r['cand_per_frame'] / 1e6:.3f} | "
                    f"{r['absdiff_per_frame'] / 1e9:.3f} | {r['gcand_per_s']:.1f} | {r['gabsdiff_per_s']:.0f} | "
                    f"{100 * r['frac_of_sad_peak']:.1f} | {100 * r['frac_of_hbm_peak']:.1f} | {r.get('cpu_ref_ms_per_frame', float('nan')):.1f} | "
                    f"{r.get('bit_exact_vs_reference', '-')} |\n")


if __name__ == "__main__":
    main()
