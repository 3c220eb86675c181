#!/usr/bin/env python
"""Copy-only probe of the host <-> device path the end-to-end number rides on.

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/pcie_probe.py                  # N GPUs of one box

Every rank moves the byte volume of one bench.py step of BASELINE config 2 (300 frames of
1920x1080 BGR in = 1.87 GB host->device, 299 frames of records + motion fields out = 7.51 GB
device->host) between pinned host memory and its GPU with plain cudaMemcpyAsync in 16-frame
chunks -- no kernels, no library code of this repository on the data path -- in four modes:

  h2d       host->device alone
  d2h       device->host alone (one 7.5 GB pinned landing zone, as bench.py's e2e leg uses)
  both      the two directions concurrently on two streams (what svc_session_encode does)
  both_ring as `both`, but device->host lands in a small pinned ring of 4 chunks that is
            overwritten round robin (what an application draining records to a file would use)

Time = max over ranks (barrier + device synchronize on both sides).  The aggregate GB/s of
`both` is the ceiling of bench.py's e2e figure at that N: e2e.frac = e2e bytes/s / this.
Prints one JSON line (rank 0) and writes --out.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--chunk", type=int, default=16, help="frames per cudaMemcpyAsync")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "pcie_probe.json"))
    a = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, H, F = a.width, a.height, a.frames
    fin = W * H * 3
    pw, ph = (W + 15) // 16 * 16, (H + 15) // 16 * 16
    fout = -(-W // 8) * -(-H // 8) * 772 + (pw // 16) * (ph // 16) * 12  # records + vectors + MADs per frame
    n_enc = F - 1
    C = a.chunk
    h_in = torch.empty(F * fin, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n_enc * fout, dtype=torch.uint8, pin_memory=True)
    h_ring = torch.empty(4 * C * fout, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(0)  # first touch before timing
    h_ring.fill_(0)
    d_in = torch.empty(2 * C * fin, dtype=torch.uint8, device="cuda")   # double-buffered device staging
    d_out = torch.empty(2 * C * fout, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def run(mode):
        k = 0
        for lo in range(0, F, C):
            n = min(C, F - lo)
            if mode in ("h2d", "both", "both_ring"):
                with torch.cuda.stream(s_in):
                    d_in[(k & 1) * C * fin:(k & 1) * C * fin + n * fin].copy_(h_in[lo * fin:(lo + n) * fin], non_blocking=True)
            if mode in ("d2h", "both", "both_ring"):
                ne = min(n, n_enc - lo) if lo < n_enc else 0
                if ne > 0:
                    src = d_out[(k & 1) * C * fout:(k & 1) * C * fout + ne * fout]
                    with torch.cuda.stream(s_out):
                        if mode == "both_ring":
                            h_ring[(k & 3) * C * fout:(k & 3) * C * fout + ne * fout].copy_(src, non_blocking=True)
                        else:
                            h_out[lo * fout:(lo + ne) * fout].copy_(src, non_blocking=True)
            k += 1

    res = {}
    for mode in ("h2d", "d2h", "both", "both_ring"):
        run(mode)  # warm-up
        barrier()
        best = None
        for _ in range(a.reps):
            barrier()
            t0 = time.perf_counter()
            run(mode)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.item())
            best = dt if best is None else min(best, dt)
        nb_in = F * fin if mode != "d2h" else 0
        nb_out = n_enc * fout if mode != "h2d" else 0
        res[mode] = {"seconds": best, "h2d_gbs_aggregate": world * nb_in / best / 1e9,
                     "d2h_gbs_aggregate": world * nb_out / best / 1e9,
                     "total_gbs_aggregate": world * (nb_in + nb_out) / best / 1e9,
                     "frames_per_s_equivalent": world * n_enc / best}
    if rank == 0:
        out = {"n_gpus": world, "gpu": torch.cuda.get_device_name(local), "host_cores": len(os.sched_getaffinity(0)),
               "frame": f"{W}x{H}", "frames_per_rank": F, "chunk_frames": C,
               "h2d_bytes_per_rank": F * fin, "d2h_bytes_per_rank": n_enc * fout, "modes": res,
               "note": "time = max over ranks of the wall time between a barrier + device synchronize and the "
                       "device synchronize after the last copy; best of %d repetitions" % a.reps}
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
