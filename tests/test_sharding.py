"""CPU: host-side frame-range sharding -- ranges, overlap frame,
gather order -- including a world_size-2 gloo run where each rank encodes its
shard (with the oracle standing in for the GPU) and rank 0 checks that the
gathered stream equals the single-rank stream."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from svc_b200.shard import gather_streams, shard_frame_ranges
from svc_b200.synth import SyntheticSequence


@pytest.mark.parametrize("n,world", [(300, 1), (300, 2), (300, 8), (601, 8), (5, 8), (2, 2), (1, 4), (0, 2)])
def test_ranges_partition_encoded_frames(n, world):
    r = shard_frame_ranges(n, world)
    assert len(r) == world
    enc = []
    for in_lo, in_hi, e_lo, e_hi in r:
        assert e_hi >= e_lo
        if e_hi > e_lo:
            assert in_lo == e_lo - 1 and in_hi == e_hi  # exactly one overlap frame in front
        enc += list(range(e_lo, e_hi))
    assert enc == list(range(1, max(n, 1)))
    sizes = [e[3] - e[2] for e in r]
    assert max(sizes) - min(sizes) <= 1


def test_synthetic_sequence_is_frame_addressable():
    a = SyntheticSequence(96, 64, 12, seed=3)
    b = SyntheticSequence(96, 64, 12, seed=3)
    assert np.array_equal(a.frames(4, 9), np.stack([b.frame(i) for i in range(4, 9)]))
    assert not np.array_equal(a.frame(1), a.frame(2))
    f = a.frame(5)
    py, px, ps, col = a.flat
    assert (f[py:py + ps, px:px + ps] == np.array(col, np.uint8)).all()  # the exact flat patch


def _encode_shard(oracle, seq, in_lo, in_hi, pw, ph):
    """Oracle stand-in for one rank's Session.encode on input frames [in_lo, in_hi)."""
    out_mv, out_st = [], []
    prev = None
    for i in range(in_lo, in_hi):
        f = seq.frame(i)
        cur = oracle.y_pyramid(f, pw, ph, 4)
        if prev is not None:
            mv, mad = oracle.hbma(prev, cur, 8)
            out_mv.append(mv)
            h, w, _ = f.shape
            out_st.append(oracle.serialize_frame(oracle.dct_planar(f, pw, ph), None, w, h, 8, 8,
                                                 pw // 16, 16, 16))
        prev = cur
    return out_mv, out_st


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, n = 64, 48, 9
    pw, ph = O.padded_dim(w, 16, 4), O.padded_dim(h, 16, 4)
    seq = SyntheticSequence(w, h, n, seed=17)
    in_lo, in_hi, e_lo, e_hi = shard_frame_ranges(n, world)[rank]
    mv, st = _encode_shard(O, seq, in_lo, in_hi, pw, ph)
    gathered = [None] * world
    dist.all_gather_object(gathered, (e_lo, e_hi, mv, st))  # host gather, rank order = frame order
    if rank == 0:
        full_mv, full_st = _encode_shard(O, seq, 0, n, pw, ph)
        g_mv = [m for part in gathered for m in part[2]]
        g_st = [s for part in gathered for s in part[3]]
        ok = len(g_mv) == n - 1 and all(np.array_equal(a, b) for a, b in zip(g_mv, full_mv))
        hdr = O.header(n, w, h, pw, ph)
        ok = ok and np.array_equal(gather_streams(hdr, g_st), gather_streams(hdr, full_st))
        ok = ok and [p[0] for p in gathered] == [r[2] for r in shard_frame_ranges(n, world)]
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_stream_equals_single_rank():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
