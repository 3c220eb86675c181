#!/usr/bin/env python
"""bench.py -- encoder hot-path throughput (HBMA + block DCT) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

One "step" = one pass of the hot path (Y pyramid -> HBMA -> DCT + stream
records) over one synthetic sequence per GPU: BASELINE.json config 2,
1920x1080 BGR, 300 input frames (299 encoded), 16x16 blocks, search range 8,
4 pyramid levels, 8x8 DCT.  N > 1 shards a longer sequence by contiguous
frame ranges with one overlap frame (weak scaling, no data-path collective).

Prints ONE JSON line (rank 0).  `value` is device-resident frames/s
(inputs already in HBM), `e2e` the same metric through the host-buffer C-ABI
call (pinned host memory, H2D and D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

import numpy as np  # noqa: E402

METRIC = "hbma_dct_encode_fps_1080p"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=300, help="input frames per GPU per step")
    ap.add_argument("--batch", type=int, default=100, help="max frames per kernel launch")
    ap.add_argument("--search-range", type=int, default=8)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--cpu-sample-frames", type=int, default=300,
                    help="input frames in the CPU baseline sample (default: the whole 300-frame workload, about 25 core-seconds on 16 cores)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sad", action="store_true", help="skip the SAD-roofline side measurement (R=32/64, L=1)")
    return ap.parse_args()


def workload_name(a):
    cfg = {(960, 540): "C1", (1920, 1080): "C2", (3840, 2160): "C4"}.get((a.width, a.height), "custom")  # BASELINE.json configs
    return (f"{cfg} synthetic {a.width}x{a.height} 8-bit BGR, {a.frames} input frames "
            f"({a.frames - 1} encoded) per GPU, 16x16 MV blocks, R={a.search_range}, "
            f"L={a.levels}, 8x8 DCT, 772-byte stream records")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- CPU arm
def cpu_hot_path_fps(a, frames, threads):
    """The reference's CPU implementation of the path on `frames` (n,h,w,3):
    Y pyramid (C port of the OpenCV calls) -> reference HBMA (compiled unmodified
    libs/motion.cpp, SSE2 entry, when it travelled; else the C port) -> per-block
    DCT + SerializeEncodedFrame (C port).  Frame-range sharded over `threads`
    host threads (the reference itself computes on one thread)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    n, h, w, _ = frames.shape
    L, R = a.levels, a.search_range
    pw, ph = O.padded_dim(w, 16, L), O.padded_dim(h, 16, L)
    use_ref = O.have_ref() and L == 4
    mvw = pw // 16

    def work(lo, hi):  # encoded frames lo..hi-1 (anchor index), needs frame lo-1
        prev = O.y_pyramid(frames[lo - 1], pw, ph, L)
        for i in range(lo, hi):
            cur = O.y_pyramid(frames[i], pw, ph, L)
            O.hbma(prev, cur, R, impl="ref_sse2" if use_ref else "oracle")
            planes = O.dct_planar(frames[i], pw, ph)
            O.serialize_frame(planes, None, w, h, 8, 8, mvw, 16, 16)
            prev = cur

    from svc_b200.shard import shard_frame_ranges
    ranges = [r for r in shard_frame_ranges(n, threads) if r[3] > r[2]]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=len(ranges)) as ex:
        list(ex.map(lambda r: work(r[2], r[3]), ranges))
    dt = time.perf_counter() - t0
    return (n - 1) / dt, dt, ("reference libs/motion.cpp (SSE2 entry) + C port of the OpenCV stages"
                              if use_ref else "C port (oracle/svc_oracle.c)")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from svc_b200.synth import SyntheticSequence
    threads = host_cores()
    n = max(2, a.cpu_sample_frames)
    frames = SyntheticSequence(a.width, a.height, n, seed=1234).frames()
    for _ in range(a.warmup):
        cpu_hot_path_fps(a, frames[: min(n, threads + 1)], threads)
    t_tot, fps_list, what = 0.0, [], ""
    for _ in range(a.steps):
        fps, dt, what = cpu_hot_path_fps(a, frames, threads)
        t_tot += dt
        fps_list.append(fps)
    value = (n - 1) * a.steps / t_tot
    sample = f"{n} input frames ({n - 1} encoded) of the workload per step; {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_tot / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step": n - 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": 0}
        mhz = [m for m, _ in self.samples]
        return {"sm_mhz": float(np.median(mhz)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(mhz)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so the pinned host
    buffers (first touch) and the copy-issuing thread sit on the GPU's NUMA node.  The e2e
    path is PCIe/host-memory bound; without this, ranks share one node's memory bandwidth."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "first": min(cpus), "last": max(cpus)}
    except Exception as e:  # no NVML / not permitted: run unbound
        return {"error": str(e)[:80]}
    return None


# --------------------------------------------------------------------------- CUDA arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    import svc_b200 as svc
    from svc_b200.shard import shard_frame_ranges

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    if not os.path.exists(svc.lib_path()) and rank == 0:
        import subprocess  # built artefacts are git-ignored: a fresh checkout builds first
        subprocess.run(["make", "-C", os.path.join(ROOT, "scalable-video-codec_b200"), "-j4"], check=True,
                       stdout=subprocess.DEVNULL)
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)  # before any pinned allocation (first-touch placement)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    svc.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    W, H, F = a.width, a.height, a.frames
    n_enc = F - 1
    # one long sequence, sharded by contiguous encoded-frame ranges (+1 overlap frame)
    total_in = world * n_enc + 1
    in_lo, in_hi, enc_lo, enc_hi = shard_frame_ranges(total_in, world)[rank]
    assert in_hi - in_lo == F
    seq = svc.SyntheticSequence(W, H, total_in, seed=1234)

    ts = torch.cuda.Stream()
    sess = svc.Session(svc.SessionConfig(frame_w=W, frame_h=H, mv_search_range=a.search_range,
                                         pyr_lvl_count=a.levels, device=local,
                                         max_batch=a.batch, cuda_stream=ts.cuda_stream))
    mvn = sess.mv_field_w * sess.mv_field_h
    fin, fst = sess.frame_in_bytes, sess.frame_stream_bytes

    # pinned host buffers (also the e2e buffers)
    h_in = svc.PinnedBuffer(F * fin)
    h_st = svc.PinnedBuffer(n_enc * fst)
    h_mv = svc.PinnedBuffer(n_enc * mvn * 8)
    h_mad = svc.PinnedBuffer(n_enc * mvn * 4)
    frames = h_in.view(np.uint8, (F, H, W, 3))
    seq.frames(in_lo, in_hi, out=frames)

    d_in = torch.empty(F * fin, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n_enc * fst, dtype=torch.uint8, device="cuda")
    d_mv = torch.empty(n_enc * mvn * 2, dtype=torch.float32, device="cuda")
    d_mad = torch.empty(n_enc * mvn, dtype=torch.float32, device="cuda")
    svc.binding._check(svc.lib().svc_memcpy_h2d(local, d_in.data_ptr(), h_in.ptr, F * fin))
    torch.cuda.synchronize()

    def step():
        sess.reset()
        ne = sess.encode_device(d_in.data_ptr(), F, d_mv.data_ptr(), d_mad.data_ptr(), d_st.data_ptr())
        assert ne == n_enc

    clocks = ClockSampler(local)
    # ---------------- device-resident timed region --------------------------------
    for _ in range(a.warmup):
        step()
    barrier()
    clocks.start()
    l0 = sess.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ts):
        e0.record(ts)
        for _ in range(a.steps):
            step()
        e1.record(ts)
    barrier()
    launches = sess.launch_count - l0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    value = world * n_enc / (ms_step * 1e-3)

    # ---------------- per-stage timing (dominant kernel roofline) -------------------
    def time_stage(fn, reps=3):
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_launch = 0
        with torch.cuda.stream(ts):
            s0.record(ts)
            for _ in range(reps):
                n_launch += fn()
            s1.record(ts)
        torch.cuda.synchronize()
        return s0.elapsed_time(s1) / n_launch  # ms per launch(-group)

    B = a.batch
    nb = F // B  # full batches only, every launch touches different frames (>> L2)

    def stage_dct():
        for b in range(nb):
            sess.run_stage(svc.STAGE_DCT_STREAM, d_in.data_ptr() + b * B * fin, B, None, None,
                           d_st.data_ptr() + (b % max(1, (n_enc // B))) * B * fst)
        return nb

    def stage_pyr():
        for b in range(nb):
            sess.run_stage(svc.STAGE_PYR_DOWN, None, B)
        return nb

    def stage_pyr_hbma():  # HBMA right behind the pyramid build, as inside a step
        for b in range(nb):
            sess.run_stage(svc.STAGE_PYR_DOWN, None, B)
            sess.run_stage(svc.STAGE_HBMA, None, B, d_mv.data_ptr(), d_mad.data_ptr())
        return nb

    stages = {}
    if nb >= 1:
        ms_dct = time_stage(stage_dct)
        ms_pyr = time_stage(stage_pyr)
        ms_hbma = max(time_stage(stage_pyr_hbma) - ms_pyr, 1e-6)
        P = sess.padded_w * sess.padded_h
        fused_y = (W == sess.padded_w)                     # K3 also writes the level-0 luma
        dct_bytes = B * (fin + fst + (P if fused_y else 0))  # read BGR once, write records (+Y) once
        pyr_bytes = B * (sum(P >> (2 * l) for l in range(a.levels - 1)) +      # read levels 0..L-2
                         sum(P >> (2 * l) for l in range(1, a.levels)))      # write levels 1..L-1
        hbma_bytes = B * (2 * sum(P >> (2 * l) for l in range(a.levels)) + mvn * 12)
        stages = {
            "dct_stream_y": {"ms_per_launch": ms_dct, "frames_per_launch": B,
                             "algorithmic_bytes": dct_bytes, "gbs": dct_bytes / ms_dct / 1e6},
            "pyr_down": {"ms_per_launch_group": ms_pyr, "frames_per_launch": B,
                         "algorithmic_bytes": pyr_bytes, "gbs": pyr_bytes / ms_pyr / 1e6},
            "hbma": {"ms_per_launch": ms_hbma, "frames_per_launch": B,
                     "algorithmic_bytes": hbma_bytes, "gbs": hbma_bytes / ms_hbma / 1e6},
        }
        traffic = None
        try:  # DRAM bytes of this kernel from the committed ncu --set full capture (same geometry only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["dct8x8_stream_kernel"]
            if (W, H) == (1920, 1080) and fused_y:
                traffic = tj["dram_bytes_per_launch"] * B / tj["frames_per_launch"]
        except Exception:
            pass
        roofline = {"kernel": "dct8x8_stream_kernel (K3: block DCT + stream records + level-0 luma)",
                    "bound": "hbm", "achieved": dct_bytes / ms_dct / 1e6, "peak": hbm_peak,
                    "unit": "GB/s", "frac": dct_bytes / ms_dct / 1e6 / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dct_bytes}
    else:
        roofline = None

    # ---------------- end to end through the host-buffer C-ABI call -------------------
    e2e = None
    if not a.no_e2e:
        def e2e_step():
            sess.reset()
            mv, mad, st = sess.encode(frames, out_mv=h_mv.view(np.float32, (n_enc, sess.mv_field_h, sess.mv_field_w, 2)),
                                      out_mad=h_mad.view(np.float32, (n_enc, sess.mv_field_h, sess.mv_field_w)),
                                      out_stream=h_st.view(np.uint8, (n_enc, fst)))
            assert st.shape[0] == n_enc

        n_e2e = max(1, min(a.steps, 5))
        for _ in range(min(a.warmup, 2)):
            e2e_step()
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(ts)
        for _ in range(n_e2e):
            e2e_step()  # blocking: returns after the last D2H landed
        x1.record(ts)
        barrier()
        t2 = torch.tensor([x0.elapsed_time(x1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_e2e = float(t2.item()) / n_e2e
        e2e = {"value": world * n_enc / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(F * fin),
               "d2h_bytes_per_step": int(n_enc * (fst + mvn * 12)),
               "ms_per_step": ms_e2e, "steps": n_e2e,
               "api": "svc_session_encode (pinned host buffers; H2D | kernels | D2H pipelined)"}
    # ---------------- SAD-bound corner of the range sweep (BASELINE config 3), N=1 only ----------
    # The default configuration's search (r = 1) is HBM-bound; the metric also asks for the SAD
    # rate against the integer roofline, so the search kernel is timed alone at R=64, L=1 on
    # frames of this workload: exact byte-absdiff count (device counters) / CUDA-event time on the
    # session stream / measured VABSDIFF4 peak of this GPU.
    sad_roofline = None
    if rank == 0 and world == 1 and F >= 10 and W % 16 == 0 and not a.no_sad:
        try:
            nf = 9
            peak_sad = svc.sad_peak(0)
            points = []
            for R_s in (32, 64):
                s2 = svc.Session(svc.SessionConfig(frame_w=W, frame_h=H, mv_search_range=R_s, pyr_lvl_count=1,
                                                   max_batch=nf, cuda_stream=ts.cuda_stream))
                mvn2 = s2.mv_field_w * s2.mv_field_h
                d_mv2 = torch.empty(nf * mvn2 * 2, dtype=torch.float32, device="cuda")
                d_mad2 = torch.empty(nf * mvn2, dtype=torch.float32, device="cuda")
                s2.run_stage(svc.STAGE_Y_PYRAMID, d_in.data_ptr(), nf)
                s2.run_stage(svc.STAGE_HBMA, None, nf, d_mv2.data_ptr(), d_mad2.data_ptr())  # warm-up
                torch.cuda.synchronize()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                q0.record(ts)
                for _ in range(reps):
                    s2.run_stage(svc.STAGE_HBMA, None, nf, d_mv2.data_ptr(), d_mad2.data_ptr())
                q1.record(ts)
                torch.cuda.synchronize()
                ms_f = q0.elapsed_time(q1) / reps / nf
                cand, absd = s2.hbma_work(nf)
                points.append({"config": f"R={R_s}, L=1 ({2 * R_s + 1}x{2 * R_s + 1} candidates per 16x16 block)",
                               "achieved": absd / nf / ms_f / 1e6, "frac": absd / nf / (ms_f * 1e-3) / peak_sad,
                               "gcand_per_s": cand / nf / ms_f / 1e6, "ms_per_frame": ms_f})
                s2.close()
            best = max(points, key=lambda q: q["frac"])
            sad_roofline = {"kernel": "K2 search alone (hbma_ebma_tile_kernel / hbma_pool_kernel), " + best["config"],
                            "bound": "int-alu (VABSDIFF4)", "achieved": best["achieved"],
                            "peak": peak_sad / 1e9, "unit": "G byte-absdiff/s", "frac": best["frac"],
                            "gcand_per_s": best["gcand_per_s"], "ms_per_frame": best["ms_per_frame"],
                            "peak_source": "measured on this GPU (dependency-free VABSDIFF4.ACC loop)",
                            "points": points}
        except Exception as e:  # never let the side measurement take the headline line down
            sad_roofline = {"error": str(e)}
    clk = clocks.stop()

    # ---------------- CPU baseline + exact SAD work counts (rank 0, N=1) ----------------
    cpu = None
    work = None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import oracle as O
        n = max(2, min(a.cpu_sample_frames, F))
        threads = host_cores()
        fps, dt, what = cpu_hot_path_fps(a, frames[:n], threads)
        cpu = {"value": fps, "unit": UNIT, "cores": threads,
               "kind": "port",
               "sample": f"first {n} input frames ({n - 1} encoded) of the workload, {dt:.2f} s; {what}"}
        pw, ph = sess.padded_w, sess.padded_h
        nc = na = 0
        k = min(n, 5)
        pyr = [O.y_pyramid(frames[i], pw, ph, a.levels) for i in range(k)]
        for i in range(1, k):
            c, d = O.hbma_count(pyr[i - 1], pyr[i], a.search_range)
            nc += c
            na += d
        work = {"candidates_per_frame": nc / (k - 1), "absdiffs_per_frame": na / (k - 1),
                "counted_on_frames": k - 1}
        if stages:
            fps_hbma = a.batch / (stages["hbma"]["ms_per_launch"] * 1e-3)
            work["hbma_gcand_per_s"] = work["candidates_per_frame"] * fps_hbma / 1e9
            work["hbma_gabsdiff_per_s"] = work["absdiffs_per_frame"] * fps_hbma / 1e9

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "frames_per_step_per_gpu": n_enc,
                       "batch_frames_per_launch": a.batch,
                       "l2": "per-step working set (%.1f GB in + %.1f GB out per GPU) far exceeds the "
                             "126 MB L2; no explicit flush" % (F * fin / 1e9, n_enc * fst / 1e9),
                       "sharding": "contiguous frame ranges, one overlap frame, no collective",
                       "host_numa_binding": numa},
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu, "stages": stages, "sad_work": work, "sad_roofline": sad_roofline,
        }
        print(json.dumps(line), flush=True)
    sess.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
