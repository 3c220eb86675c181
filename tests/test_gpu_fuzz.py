"""GPU: seeded random configurations through the session API against the oracle --
frame sizes that pad on either/both axes, tiny frames, every level count, non-default MV
and transform blocks (generic kernels), search ranges on all three HBMA kernels,
uneven pushes, block types."""
import numpy as np
import pytest

from svc_b200.synth import SyntheticSequence

pytestmark = pytest.mark.gpu
DCT_TOL = 1e-3


def _configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        L = int(rng.integers(1, 6))
        mb = int(rng.choice([8, 16, 16, 16, 32]))
        if mb % (1 << (L - 1)):
            continue
        tb = int(rng.choice([t for t in (2, 4, 8, 8, 8, 16) if t <= mb and mb % t == 0]))
        r_top = int(rng.choice([1, 1, 2, 3, 4, 6, 9]))
        R = r_top * (1 << (L - 1)) + int(rng.integers(0, 1 << (L - 1)))
        w = int(rng.integers(8, 200))
        h = int(rng.integers(8, 140))
        n_frames = int(rng.integers(2, 5))
        batch = int(rng.integers(1, 4))
        out.append((w, h, L, mb, tb, R, n_frames, batch))
    return out


@pytest.mark.parametrize("cfg", _configs(36, 2026))
def test_random_session_config(gpu, oracle, cfg):
    w, h, L, mb, tb, R, n, batch = cfg
    frames = SyntheticSequence(w, h, n, seed=w * 131 + h, n_rects=2).frames()
    sc = gpu.SessionConfig(frame_w=w, frame_h=h, mv_block_w=mb, mv_block_h=mb, mv_search_range=R,
                           pyr_lvl_count=L, transform_block_w=tb, transform_block_h=tb, max_batch=batch)
    with gpu.Session(sc) as s:
        pw, ph = s.padded_w, s.padded_h
        assert (pw, ph) == (oracle.padded_dim(w, mb, L), oracle.padded_dim(h, mb, L))
        rng = np.random.default_rng(7)
        bt = rng.integers(0, 4, size=(n - 1, s.mv_field_h * s.mv_field_w)).astype(np.uint32)
        mv, mad, st = s.encode(frames, block_types=bt)
        assert mv.shape[0] == n - 1
        rec = 1 + 3 * tb * tb
        pyr = [oracle.y_pyramid(f, pw, ph, L) for f in frames]
        for i in range(1, n):
            emv, emad = oracle.hbma(pyr[i - 1], pyr[i], R, mb, mb)
            assert np.array_equal(mv[i - 1], emv), (cfg, i)
            assert np.array_equal(mad[i - 1], emad), (cfg, i)
            planes = oracle.dct_planar(frames[i], pw, ph, tb, tb)
            exp = oracle.serialize_frame(planes, bt[i - 1], w, h, tb, tb, pw // mb, mb, mb)
            g = st[i - 1].view(np.uint32).reshape(-1, rec)
            e = exp.view(np.uint32).reshape(-1, rec)
            assert np.array_equal(g[:, 0], e[:, 0]), (cfg, i)
            assert np.abs(g[:, 1:].view(np.float32) - e[:, 1:].view(np.float32)).max() <= DCT_TOL, (cfg, i)


def _wide_configs(n, seed):
    """16x16 blocks with top-level ranges 5..70: the pooled / shared-window / column-striped kernels
    of k_hbma_pool.cu in session mode (several frame pairs per launch, frames smaller than the window)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        L = int(rng.choice([1, 1, 1, 2, 2, 3]))
        r_top = int(rng.choice([5, 7, 8, 11, 16, 19, 32, 33, 40, 64, 70]))
        if L > 1 and r_top > 64:
            continue
        R = r_top * (1 << (L - 1)) + int(rng.integers(0, 1 << (L - 1)))
        w = int(rng.integers(16, 280))
        h = int(rng.integers(16, 200))
        out.append((w, h, L, R, int(rng.integers(2, 5)), int(rng.integers(1, 4))))
    return out


@pytest.mark.parametrize("cfg", _wide_configs(16, 77))
def test_random_wide_range_session(gpu, oracle, cfg):
    w, h, L, R, n, batch = cfg
    frames = SyntheticSequence(w, h, n, seed=w * 17 + h, n_rects=3).frames()
    sc = gpu.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L, max_batch=batch)
    with gpu.Session(sc) as s:
        pw, ph = s.padded_w, s.padded_h
        mv, mad, _ = s.encode(frames)
        pyr = [oracle.y_pyramid(f, pw, ph, L) for f in frames]
        for i in range(1, n):
            emv, emad = oracle.hbma(pyr[i - 1], pyr[i], R)
            assert np.array_equal(mv[i - 1], emv), (cfg, i)
            assert np.array_equal(mad[i - 1], emad), (cfg, i)


@pytest.mark.parametrize("w,h,mb,L,tb,n,batch", [(320, 180, 16, 4, 16, 5, 3), (320, 176, 16, 4, 4, 4, 2),
                                                  (64, 40, 16, 3, 16, 3, 1), (200, 90, 8, 3, 4, 4, 3),
                                                  (1920, 1080, 16, 4, 16, 3, 2), (960, 540, 16, 4, 4, 3, 2),
                                                  # padded block rows that carry luma but no record (32x32 MV blocks)
                                                  (64, 40, 32, 3, 16, 3, 2), (192, 72, 32, 2, 16, 4, 3),
                                                  (96, 36, 32, 3, 4, 3, 1)])
def test_session_fused_square_transform_blocks(gpu, oracle, w, h, mb, L, tb, n, batch):
    """Frames without horizontal padding and 16x16 / 4x4 transform blocks: the fused stream kernels
    (records + level-0 luma in one pass) inside a session -- the motion field checks their luma."""
    frames = SyntheticSequence(w, h, n, seed=w + 3 * h + tb, n_rects=3).frames()
    sc = gpu.SessionConfig(frame_w=w, frame_h=h, mv_block_w=mb, mv_block_h=mb, pyr_lvl_count=L,
                           mv_search_range=1 << (L - 1), transform_block_w=tb, transform_block_h=tb,
                           max_batch=batch)
    with gpu.Session(sc) as s:
        pw, ph = s.padded_w, s.padded_h
        assert pw == w
        rng = np.random.default_rng(11)
        bt = rng.integers(0, 5, size=(n - 1, s.mv_field_h * s.mv_field_w)).astype(np.uint32)
        mv, mad, st = s.encode(frames, block_types=bt)
        rec = 1 + 3 * tb * tb
        pyr = [oracle.y_pyramid(f, pw, ph, L) for f in frames]
        for i in range(1, n):
            emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 1 << (L - 1), mb, mb)
            assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad), i
            planes = oracle.dct_planar(frames[i], pw, ph, tb, tb)
            exp = oracle.serialize_frame(planes, bt[i - 1], w, h, tb, tb, pw // mb, mb, mb)
            g = st[i - 1].view(np.uint32).reshape(-1, rec)
            e = exp.view(np.uint32).reshape(-1, rec)
            assert np.array_equal(g[:, 0], e[:, 0]), i
            assert np.abs(g[:, 1:].view(np.float32) - e[:, 1:].view(np.float32)).max() <= DCT_TOL, i


def _default_cfgs(n, seed):
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(16, 420)), int(rng.integers(16, 300)), int(rng.integers(0, 4)), int(rng.integers(1, 1 << 30)))
            for _ in range(n)]


@pytest.mark.parametrize("cfg", _default_cfgs(24, 77))
def test_random_default_config_strip_kernels(gpu, oracle, cfg):
    """Encoder default (16x16, R=8, L=4: hbma_strip_coarse/fine_kernel) on random frame sizes and four
    kinds of content: the synthetic sequence with pans up to the full reach of the search (15 pixels at
    level 0), unrelated frames (vectors all over the bounded window), white noise, and frames that differ
    by a constant (every candidate ties somewhere).  Against the oracle and the tile kernel."""
    w, h, kind, seed = cfg
    rng = np.random.default_rng(seed)
    if kind == 0:
        frames = SyntheticSequence(w, h, 4, seed=seed % 9973, n_rects=3, max_pan=15).frames()
    elif kind == 1:
        frames = np.stack([SyntheticSequence(w, h, 1, seed=seed % 9973 + k).frame(0) for k in range(3)])
    elif kind == 2:
        frames = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    else:
        base = SyntheticSequence(w, h, 1, seed=seed % 9973).frame(0).astype(np.int32)
        frames = np.stack([np.clip(base + d, 0, 255).astype(np.uint8) for d in (0, 9, -7)])
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=2)) as s:
        mv, mad, _ = s.encode(frames, want_stream=False)
        pw, ph = s.padded_w, s.padded_h
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, hbma_kernel_family=gpu.HBMA_FAMILY_TILE)) as s:
        mv_t, mad_t, _ = s.encode(frames, want_stream=False)
    assert np.array_equal(mv, mv_t) and np.array_equal(mad, mad_t), cfg
    pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
    for i in range(1, len(frames)):
        emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 8)
        assert np.array_equal(mv[i - 1], emv), (cfg, i)
        assert np.array_equal(mad[i - 1], emad), (cfg, i)
