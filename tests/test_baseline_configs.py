"""Parity on the literal BASELINE.json configurations:

  C1  960x540, 30 frames, defaults: EVERY encoded frame (vectors, MADs, records, header);
  C3  the SAD-bound rows of the 1080p range / level sweep against golden fields generated once by
      the unmodified reference (tests/golden/make_golden_sweep.py);
  C4  3840x2160, 600 frames sharded 2- and 8-way: both frame pairs at EVERY shard boundary, encoded
      the way the shards encode them (overlap frame first, tracked-only) and in one piece.

(C2, 1920x1080 x 300 frames, is checked on all 299 frames inside bench.py itself -- the `parity`
object of its JSON line -- and on a 12-frame prefix in test_gpu_parity.py.)
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from svc_b200.shard import shard_frame_ranges
from svc_b200.synth import SyntheticSequence


def _records_err(got, exp, rec=772):
    g, e = got.reshape(-1, rec), exp.reshape(-1, rec)
    assert np.array_equal(g[:, :4], e[:, :4])  # block-type words
    return float(np.abs(g[:, 4:].copy().view(np.float32) - e[:, 4:].copy().view(np.float32)).max())


@pytest.mark.gpu
def test_c1_960x540x30_every_frame(gpu, oracle):
    """BASELINE config 1: the workload the reference's CPU encoder is quoted on."""
    w, h, n = 960, 540, 30
    frames = SyntheticSequence(w, h, n, seed=1234).frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=16)) as s:
        mv, mad, st = s.encode(frames)
        pw, ph = s.padded_w, s.padded_h
        hdr = s.header(n)
    assert (pw, ph) == (960, 544) and mv.shape[0] == n - 1
    assert np.array_equal(hdr, oracle.header(n, w, h, pw, ph))
    impl = "ref_sse2" if oracle.have_ref() else "oracle"
    pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
    worst = 0.0
    for i in range(1, n):
        emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 8, impl=impl)
        assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad), i
        exp = oracle.serialize_frame(oracle.dct_planar(frames[i], pw, ph), None, w, h, 8, 8, pw // 16, 16, 16)
        worst = max(worst, _records_err(st[i - 1], exp))
    assert worst <= 1e-3


def _c4_boundaries():
    out = set()
    for world in (2, 8):
        for (in_lo, in_hi, enc_lo, enc_hi) in shard_frame_ranges(600, world)[1:]:
            out.add(enc_lo)
    return sorted(out)


@pytest.mark.gpu
def test_c4_4k_600_frames_shard_boundaries(gpu, oracle):
    """BASELINE config 4.  Shard k of a 2- or 8-way split starts with the overlap frame enc_lo - 1
    (tracked only); the pair before the boundary belongs to shard k - 1.  Only frame t-1 crosses a
    boundary (libs/encoder.cpp:472-476, 661-663), so both pairs must equal the reference and the
    one-piece encode of the same three frames byte for byte."""
    w, h, n = 3840, 2160, 600
    seq = SyntheticSequence(w, h, n, seed=1234)
    bounds = _c4_boundaries()
    assert len(bounds) == 7 and 301 in bounds  # 599 encoded frames: 300 + 299, and 7 x 75 + 74
    impl = "ref_sse2" if oracle.have_ref() else "oracle"
    worst = 0.0
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=4)) as s:
        pw, ph = s.padded_w, s.padded_h
        assert (pw, ph) == (3840, 2160)
        for b in bounds:
            fr = np.stack([seq.frame(b - 2), seq.frame(b - 1), seq.frame(b)])
            pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in fr]
            s.reset()
            mv_all, mad_all, st_all = s.encode(fr)                 # one piece: pairs (b-2,b-1), (b-1,b)
            s.reset()
            mv_a, mad_a, st_a = s.encode(fr[:2])                   # tail of shard k-1
            s.reset()
            mv_b, mad_b, st_b = s.encode(fr[1:])                   # head of shard k (overlap frame b-1)
            assert np.array_equal(mv_all[0], mv_a[0]) and np.array_equal(mv_all[1], mv_b[0])
            assert np.array_equal(mad_all[0], mad_a[0]) and np.array_equal(mad_all[1], mad_b[0])
            assert np.array_equal(st_all[0], st_a[0]) and np.array_equal(st_all[1], st_b[0])
            for k in (0, 1):
                emv, emad = oracle.hbma(pyr[k], pyr[k + 1], 8, impl=impl)
                assert np.array_equal(mv_all[k], emv) and np.array_equal(mad_all[k], emad), (b, k)
            exp = oracle.serialize_frame(oracle.dct_planar(fr[2], pw, ph), None, w, h, 8, 8, pw // 16, 16, 16)
            worst = max(worst, _records_err(st_b[0], exp))
    assert worst <= 1e-3


SWEEP_GOLDEN = os.path.join(GOLDEN, "sweep_1080p.npz")


@pytest.mark.gpu
@pytest.mark.parametrize("R,L", [(32, 1), (64, 1), (64, 2), (64, 3), (32, 2)])
def test_c3_sweep_wide_range_rows_1080p_vs_reference_golden(gpu, R, L):
    """BASELINE config 3: the wide-range rows of the 1080p sweep (tens of seconds each on the scalar
    reference) against fields the unmodified reference produced once (make_golden_sweep.py)."""
    g = np.load(SWEEP_GOLDEN)
    w, h = int(g["width"]), int(g["height"])
    seq = SyntheticSequence(w, h, int(g["n_frames"]), seed=int(g["seed"]))  # the sweep tool's sequence
    fr = np.stack([seq.frame(1), seq.frame(2)])
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L, max_batch=2)) as s:
        mv, mad, _ = s.encode(fr, want_stream=False)
    assert np.array_equal(mv[0], g[f"mv_R{R}_L{L}"].astype(np.float32))
    assert np.array_equal(mad[0], g[f"mad_R{R}_L{L}"])


def test_sweep_golden_is_committed_and_well_formed():
    g = np.load(SWEEP_GOLDEN)
    assert (int(g["width"]), int(g["height"]), int(g["seed"]), int(g["n_frames"])) == (1920, 1080, 1234, 45)
    for R, L in [(32, 1), (64, 1), (64, 2), (64, 3), (32, 2)]:
        mv, mad = g[f"mv_R{R}_L{L}"], g[f"mad_R{R}_L{L}"]
        assert mv.shape == (68, 120, 2) and mad.shape == (68, 120) and mad.dtype == np.float32
        reach = (R >> (L - 1)) * ((1 << L) - 1)
        assert np.abs(mv).max() <= reach
