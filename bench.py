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
    ap.add_argument("--cpu-sample-frames", type=int, default=0,
                    help="input frames in the CPU baseline sample (default 0: the whole --frames workload, "
                         "about 25 core-seconds for 300 frames of 1080p)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check of the measured outputs")
    ap.add_argument("--host-chunk", type=int, default=0, help="frames per stage of the host-path pipeline (0 = library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sad", action="store_true", help="skip the SAD-roofline side measurement (R=32/64, L=1)")
    return ap.parse_args()


def workload_name(a):
    # BASELINE.json configs, keyed on what is actually run
    if (a.width, a.height, a.frames) == (960, 540, 30):
        cfg = "C1"
    elif (a.width, a.height, a.frames) == (1920, 1080, 300):
        cfg = "C2"
    elif (a.width, a.height) == (3840, 2160) and a.frames * a.gpus == 600:
        cfg = "C4"
    else:
        cfg = "custom"
    return (f"{cfg} synthetic {a.width}x{a.height} 8-bit BGR, {a.frames} input frames "
            f"({a.frames - 1} encoded) per GPU, 16x16 MV blocks, R={a.search_range}, "
            f"L={a.levels}, 8x8 DCT, 772-byte stream records")


def config_dict(a):
    """The `config` object of the JSON line: identical in both arms (`--impl ours|reference`)."""
    fin = a.width * a.height * 3
    fst = -(-a.width // 8) * -(-a.height // 8) * 772
    return {"workload": workload_name(a), "frames_per_step_per_gpu": a.frames - 1,
            "batch_frames_per_launch": a.batch,
            "l2": "per-step working set (%.1f GB in + %.1f GB out per GPU) far exceeds the "
                  "126 MB L2; no explicit flush" % (a.frames * fin / 1e9, (a.frames - 1) * fst / 1e9),
            "sharding": "contiguous frame ranges, one overlap frame, no collective"}


ORACLE_PIN = ("MV/MAD: unmodified reference libs/motion.cpp (oracle/_ref, SSE2 entry); Y/pyramid/DCT/"
              "serializer: oracle/svc_oracle.c pinned to python cv2 4.13 fixtures (the reference names "
              "OpenCV 3.4.*, libs/encoder.cpp is not buildable here)")


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
    host threads (the reference itself computes on one thread).
    Returns (frames/s, seconds, description, per-stage busy ms per frame and core)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    n, h, w, _ = frames.shape
    L, R = a.levels, a.search_range
    pw, ph = O.padded_dim(w, 16, L), O.padded_dim(h, 16, L)
    use_ref = O.have_ref() and L == 4
    mvw = pw // 16
    pc = time.perf_counter

    def work(lo, hi):  # encoded frames lo..hi-1 (anchor index), needs frame lo-1
        t_y = t_m = t_d = 0.0
        t0 = pc()
        prev = O.y_pyramid(frames[lo - 1], pw, ph, L)
        t_y += pc() - t0
        for i in range(lo, hi):
            t0 = pc()
            cur = O.y_pyramid(frames[i], pw, ph, L)
            t1 = pc()
            O.hbma(prev, cur, R, impl="ref_sse2" if use_ref else "oracle")
            t2 = pc()
            planes = O.dct_planar(frames[i], pw, ph)
            O.serialize_frame(planes, None, w, h, 8, 8, mvw, 16, 16)
            t3 = pc()
            t_y += t1 - t0
            t_m += t2 - t1
            t_d += t3 - t2
            prev = cur
        return t_y, t_m, t_d

    from svc_b200.shard import shard_frame_ranges
    ranges = [r for r in shard_frame_ranges(n, threads) if r[3] > r[2]]
    t0 = pc()
    with ThreadPoolExecutor(max_workers=len(ranges)) as ex:
        parts = list(ex.map(lambda r: work(r[2], r[3]), ranges))
    dt = pc() - t0
    ne = n - 1
    split = {"ypyr_ms": 1e3 * sum(p[0] for p in parts) / ne, "hbma_ms": 1e3 * sum(p[1] for p in parts) / ne,
             "dct_serialize_ms": 1e3 * sum(p[2] for p in parts) / ne,
             "note": "busy milliseconds per encoded frame on one core, summed over the shard threads: "
                     "ypyr = C port of copyMakeBorder+cvtColor+extractChannel+buildPyramid, hbma = "
                     + ("reference EstimateMotionHierarchical16x16Sse2" if use_ref else "C port of the HBMA")
                     + ", dct_serialize = C port of the per-block DCT + SerializeEncodedFrame"}
    return ne / dt, dt, ("reference libs/motion.cpp (SSE2 entry) + C port of the OpenCV stages"
                         if use_ref else "C port (oracle/svc_oracle.c)"), split


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from svc_b200.synth import SyntheticSequence
    threads = host_cores()
    n = max(2, a.cpu_sample_frames or a.frames)
    frames = SyntheticSequence(a.width, a.height, n, seed=1234).frames()
    for _ in range(a.warmup):
        cpu_hot_path_fps(a, frames[: min(n, threads + 1)], threads)
    t_tot, what, splits = 0.0, "", []
    for _ in range(a.steps):
        fps, dt, what, split = cpu_hot_path_fps(a, frames, threads)
        t_tot += dt
        splits.append(split)
    value = (n - 1) * a.steps / t_tot
    sample = f"{n} input frames ({n - 1} encoded) of the workload per step; {what}"
    stage_split = {k: float(np.mean([sp[k] for sp in splits])) for k in ("ypyr_ms", "hbma_ms", "dct_serialize_ms")}
    stage_split["note"] = splits[0]["note"] if splits else ""
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_tot / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": config_dict(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads,
                         "kind": "port", "sample": sample, "stage_split": stage_split},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "oracle_pin": ORACLE_PIN,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- parity
def check_parity(a, frames, outputs, pw, ph, threads, dct_frames=8):
    """Compare the outputs of the measured workload with the checkers: EVERY motion field / MAD
    field against the compiled reference (EstimateMotionHierarchical16x16Sse2, libs/motion.cpp:691-749;
    the generic entry or the C port for other level counts), the records of `dct_frames` evenly
    spaced frames against the oracle DCT + serializer (libs/encoder.cpp:323-339, 222-269).
    frames: (n,h,w,3) input.  outputs: {name: (mv (n-1,mh,mw,2), mad (n-1,mh,mw), stream_of)} with
    stream_of(k) -> record bytes of encoded frame k (0-based)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    n, h, w, _ = frames.shape
    L, R = a.levels, a.search_range
    impl = "ref_sse2" if (O.have_ref() and L == 4) else ("ref" if O.have_ref() else "oracle")
    mvw = pw // 16
    ne = n - 1
    bad = {name: [] for name in outputs}

    def work(lo, hi):
        prev = O.y_pyramid(frames[lo - 1], pw, ph, L)
        for i in range(lo, hi):
            cur = O.y_pyramid(frames[i], pw, ph, L)
            emv, emad = O.hbma(prev, cur, R, impl=impl)
            for name, (mv, mad, _) in outputs.items():
                if not (np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad)):
                    bad[name].append(i)
            prev = cur

    from svc_b200.shard import shard_frame_ranges
    ranges = [r for r in shard_frame_ranges(n, max(1, threads)) if r[3] > r[2]]
    with ThreadPoolExecutor(max_workers=len(ranges)) as ex:
        list(ex.map(lambda r: work(r[2], r[3]), ranges))

    picks = sorted({int(round(x)) for x in np.linspace(1, ne, min(dct_frames, ne))})

    def dct_err(i):
        exp = O.serialize_frame(O.dct_planar(frames[i], pw, ph), None, w, h, 8, 8, mvw, 16, 16).reshape(-1, 772)
        ef = exp[:, 4:].copy().view(np.float32)
        err, bt_ok = 0.0, True
        for name, (_, _, stream_of) in outputs.items():
            g = np.asarray(stream_of(i - 1)).reshape(-1, 772)
            bt_ok &= bool(np.array_equal(exp[:, :4], g[:, :4]))
            err = max(err, float(np.abs(ef - g[:, 4:].copy().view(np.float32)).max()))
        return err, bt_ok

    with ThreadPoolExecutor(max_workers=max(1, min(threads, len(picks)))) as ex:
        errs = list(ex.map(dct_err, picks))
    dct_max = max(e for e, _ in errs)
    mv_ok = not any(bad.values())
    res = {"mv_frames": ne, "mv_bit_exact": mv_ok, "mad_bit_exact": mv_ok,
           "outputs_checked": sorted(outputs),
           "mv_checker": {"ref_sse2": "reference EstimateMotionHierarchical16x16Sse2 (oracle/_ref), every frame",
                          "ref": "reference EstimateMotionHierarchical (oracle/_ref), every frame",
                          "oracle": "C port (oracle/svc_oracle.c), every frame"}[impl],
           "dct_frames": len(picks), "dct_max_abs_err": dct_max, "dct_tolerance_abs": 1e-3,
           "block_type_words_exact": all(ok for _, ok in errs), "oracle_pin": ORACLE_PIN}
    if not mv_ok:
        res["first_bad_frames"] = {k: sorted(v)[:8] for k, v in bad.items() if v}
    res["ok"] = bool(mv_ok and dct_max <= 1e-3 and res["block_type_words_exact"])
    return res


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.segments = {}  # name -> [t0, t1] (perf_counter)
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
                rs = [name for bit, name in names.items() if r & bit]
                self.samples.append((mhz, util, time.perf_counter(), rs))
                self.reasons.update(rs)
            except Exception:
                pass
            time.sleep(0.002)

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
        mhz = [m[0] for m in self.samples]
        out = {"sm_mhz": float(np.median(mhz)), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(mhz)}
        per = {}
        for name, (t0, t1) in self.segments.items():
            seg = [m for m in self.samples if t0 <= m[2] <= t1]
            if seg:
                per[name] = {"sm_mhz": float(np.median([m[0] for m in seg])), "samples": len(seg),
                             "reasons": sorted({r for m in seg for r in m[3]})}
        if per:
            out["per_stage"] = per
        return out

    def segment(self, name):
        sampler = self

        class _Seg:
            def __enter__(self):
                sampler.segments[name] = [time.perf_counter(), float("inf")]

            def __exit__(self, *exc):
                sampler.segments[name][1] = time.perf_counter()
        return _Seg()


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
            os.environ["NCCL_DEBUG"] = "WARN"
        # NCCL writes its version banner to stdout when the first communicator is created: keep stdout to
        # the one JSON line by pointing fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
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
                                         max_batch=a.batch, cuda_stream=ts.cuda_stream,
                                         host_chunk_frames=a.host_chunk))
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
    d_st = torch.empty(F * fst, dtype=torch.uint8, device="cuda")  # (F, not n_enc: the stage timings run whole batches)
    # (F fields, not n_enc: the stage timings search whole batches of F // batch * batch pairs, slot 0 included)
    d_mv = torch.empty(F * mvn * 2, dtype=torch.float32, device="cuda")
    d_mad = torch.empty(F * mvn, dtype=torch.float32, device="cuda")
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
    with clocks.segment("step"), torch.cuda.stream(ts):
        e0.record(ts)
        for _ in range(a.steps):
            step()
        e1.record(ts)
        torch.cuda.synchronize()
    barrier()
    launches = sess.launch_count - l0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    value = world * n_enc / (ms_step * 1e-3)

    # outputs of the LAST timed step, kept for the parity check (the stage timings below overwrite them)
    par_out = {}
    if not a.no_parity:
        mh_, mw_ = sess.mv_field_h, sess.mv_field_w
        ne_ = n_enc
        picks_ = sorted({int(round(x)) - 1 for x in np.linspace(1, ne_, min(8, ne_))})
        st_dev = {k: d_st[k * fst:(k + 1) * fst].cpu().numpy() for k in picks_}
        par_out["device_resident"] = (d_mv[:ne_ * mvn * 2].cpu().numpy().reshape(ne_, mh_, mw_, 2),
                                      d_mad[:ne_ * mvn].cpu().numpy().reshape(ne_, mh_, mw_), st_dev.__getitem__)

    # ---------------- per-stage timing (dominant kernel roofline) -------------------
    def time_stage(fn, name, reps=40):
        fn()  # untimed warm-up of this stage
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_launch = 0
        with clocks.segment(name), torch.cuda.stream(ts):
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
        ms_dct = time_stage(stage_dct, "dct_stream_y")
        ms_pyr = time_stage(stage_pyr, "pyr_down")
        ms_hbma = max(time_stage(stage_pyr_hbma, "hbma") - ms_pyr, 1e-6)
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
        # DRAM bytes per launch from the committed `ncu --set full` captures of these kernels
        # (profiles/traffic.json; same geometry only -- a profiler number, not measured in this run)
        traffic = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if (W, H, a.levels, a.search_range) == (1920, 1080, 4, 8) and fused_y:
                for k, v in tj["stages"].items():
                    traffic[k] = v["dram_bytes_per_launch"] * B / v["frames_per_launch"]
        except Exception:
            pass
        kernels = {"dct_stream_y": "dct8x8_stream_kernel<withY> (K3: block DCT + stream records + level-0 luma)",
                   "pyr_down": "pyr_down kernels (K1b: pyramid levels 1..L-1 of the batch)",
                   "hbma": ("hbma_strip_coarse_kernel + hbma_strip_fine_kernel (K2: 4-level search, r = 1; against the HBM "
                            "roofline at this range, SURVEY 8d)") if (a.levels, a.search_range) == (4, 8) else
                           f"K2 search kernels the dispatcher picks for L={a.levels}, R={a.search_range}"}
        rooflines = {}
        for k, st_ in stages.items():
            rooflines[k] = {"kernel": kernels[k], "bound": "hbm", "achieved": st_["gbs"], "peak": hbm_peak,
                            "unit": "GB/s", "frac": st_["gbs"] / hbm_peak, "traffic": traffic.get(k),
                            "algorithmic_bytes_per_launch": st_["algorithmic_bytes"]}
        roofline = dict(rooflines["dct_stream_y"], peak_source=peak_src)
    else:
        roofline = None
        rooflines = {}

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
        if not a.no_parity:
            # (copies of the sampled frames' records: the copy-only probe below reuses the landing zone)
            st_view = h_st.view(np.uint8, (n_enc, fst))
            st_host = {k: st_view[k].copy() for k in picks_}
            par_out["e2e_host_buffers"] = (h_mv.view(np.float32, (n_enc, sess.mv_field_h, sess.mv_field_w, 2)).copy(),
                                           h_mad.view(np.float32, (n_enc, sess.mv_field_h, sess.mv_field_w)).copy(),
                                           st_host.__getitem__)
    # ---------------- copy-only ceiling of the e2e figure, measured in this run ----------------------
    # The same pinned buffers and byte volumes moved by plain cudaMemcpyAsync on two streams, no kernels
    # (tools/pcie_probe.py, mode `both`), all ranks at once: what the host <-> device path of this box
    # delivers at this N.  e2e.frac = e2e / that.
    if e2e is not None:
        try:
            t_in = torch.from_numpy(h_in.array)
            t_out = torch.from_numpy(h_st.array)
            C = 16
            d_ci = torch.empty(2 * C * fin, dtype=torch.uint8, device="cuda")
            d_co = torch.empty(2 * C * fst, dtype=torch.uint8, device="cuda")
            s_i, s_o = torch.cuda.Stream(), torch.cuda.Stream()

            def copy_only():
                k = 0
                for lo in range(0, F, C):
                    n = min(C, F - lo)
                    with torch.cuda.stream(s_i):
                        d_ci[(k & 1) * C * fin:(k & 1) * C * fin + n * fin].copy_(t_in[lo * fin:(lo + n) * fin], non_blocking=True)
                    ne = max(0, min(n, n_enc - lo))
                    if ne:
                        with torch.cuda.stream(s_o):
                            t_out[lo * fst:(lo + ne) * fst].copy_(d_co[(k & 1) * C * fst:(k & 1) * C * fst + ne * fst], non_blocking=True)
                    k += 1

            copy_only()
            tot, reps_c = 0.0, n_e2e  # the same number of repetitions as the e2e leg, mean like it
            for _ in range(reps_c):
                barrier()
                t0 = time.perf_counter()
                copy_only()
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                tot += float(dt.item())
            best = tot / reps_c
            ceil_fps = world * n_enc / best
            e2e["copy_only_ceiling"] = {"value": ceil_fps, "unit": UNIT, "seconds": best,
                                        "aggregate_gbs": world * (F * fin + n_enc * fst) / best / 1e9,
                                        "what": "same pinned buffers and bytes, cudaMemcpyAsync H2D || D2H in 16-frame "
                                                "chunks, no kernels, all ranks at once (max over ranks, mean over as many "
                                                "repetitions as the e2e leg)"}
            e2e["frac"] = e2e["value"] / ceil_fps
            del d_ci, d_co
        except Exception as ex:  # never let the side measurement take the headline line down
            e2e["copy_only_ceiling"] = {"error": str(ex)[:200]}
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
        n = max(2, min(a.cpu_sample_frames or F, F))
        threads = host_cores()
        fps, dt, what, split = cpu_hot_path_fps(a, frames[:n], threads)
        cpu = {"value": fps, "unit": UNIT, "cores": threads,
               "kind": "port",
               "sample": f"first {n} input frames ({n - 1} encoded) of the workload, {dt:.2f} s; {what}",
               "stage_split": split}
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

    # ---------------- parity of the measured workload (every rank checks its own shard) ----------------
    parity = None
    if par_out:
        thr = max(1, host_cores() // world)
        parity = check_parity(a, frames, par_out, sess.padded_w, sess.padded_h, thr)
        if world > 1:
            flags = torch.tensor([1.0 if parity["ok"] else 0.0, -parity["dct_max_abs_err"]],
                                 dtype=torch.float64, device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            all_ok = bool(flags[0].item() > 0.5)
            parity["dct_max_abs_err"] = float(-flags[1].item())
            parity["mv_frames"] = world * n_enc
            if not all_ok:
                parity["ok"] = parity["mv_bit_exact"] = parity["mad_bit_exact"] = False
                parity["note"] = "a rank other than 0 reported a mismatch"
        if not parity["ok"]:
            print(f"bench.py: rank {rank}: PARITY MISMATCH {json.dumps(parity)}", file=sys.stderr, flush=True)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": config_dict(a),
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu, "parity": parity, "rooflines": rooflines, "stages": stages,
            "sad_work": work, "sad_roofline": sad_roofline, "host": {"numa_binding": numa, "cores": host_cores()},
        }
        print(json.dumps(line), flush=True)
    sess.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
