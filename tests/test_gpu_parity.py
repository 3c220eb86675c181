"""GPU: the CUDA path, called through the C ABI, against the oracle, the
golden fixtures (compiled reference + cv2) and, where it travelled, the
compiled reference itself.  Integer work is bit-exact; DCT coefficients are
within DCT_TOL absolute of OpenCV / the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from svc_b200.synth import SyntheticSequence

pytestmark = pytest.mark.gpu

DCT_TOL = 1e-3  # absolute, per coefficient (coefficients reach 2040)


# ---------------------------------------------------------------- K1: Y pyramid
@pytest.mark.parametrize("name,n", [("small_default.npz", 3), ("aligned_default.npz", 2)])
def test_y_pyramid_golden(gpu, name, n):
    g = load_golden(name)
    for i in range(n):
        pyr = gpu.y_pyramid(g["frames"][i], int(g["pw"]), int(g["ph"]), 4)
        for l in range(4):
            assert np.array_equal(pyr[l], g[f"pyr{i}_{l}"]), (name, i, l)


@pytest.mark.parametrize("w,h,L,B", [(1920, 1080, 4, 16), (960, 540, 4, 16), (333, 77, 3, 8),
                                     (16, 16, 4, 16), (50, 34, 2, 2), (8, 8, 1, 8), (130, 70, 5, 16)])
def test_y_pyramid_vs_oracle(gpu, oracle, w, h, L, B):
    rng = np.random.default_rng(w * 7 + h)
    f = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    pw, ph = gpu.padded_dim(w, B, L), gpu.padded_dim(h, B, L)
    got, exp = gpu.y_pyramid(f, pw, ph, L), oracle.y_pyramid(f, pw, ph, L)
    for l in range(L):
        assert np.array_equal(got[l], exp[l]), l


@pytest.mark.parametrize("w,h,L", [(1920, 1080, 4), (3840, 2160, 4), (416, 240, 4), (400, 232, 4), (208, 120, 5),
                                   (1000, 600, 5), (520, 264, 4), (64, 48, 4), (56, 40, 4), (2048, 1024, 6),
                                   (776, 392, 4), (32, 32, 4), (48, 16, 3)])
def test_y_pyramid_fused_upper_levels_vs_oracle(gpu, oracle, w, h, L):
    """Levels above 1 come two per launch (pyr_down2_kernel): full and partial tiles in both directions,
    REFLECT_101 at the border of BOTH levels, regions larger than the level, 5 and 6 levels (two fused
    launches / a fused and a single one), frames too small for the fused kernel."""
    rng = np.random.default_rng(w * 31 + h + L)
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    bgr[: h // 3, : w // 3] = 255  # saturated corner: the rounding of each level matters
    pw, ph = oracle.padded_dim(w, 16, L), oracle.padded_dim(h, 16, L)
    got = gpu.y_pyramid(bgr, pw, ph, L)
    exp = oracle.y_pyramid(bgr, pw, ph, L)
    for l in range(L):
        assert np.array_equal(got[l], exp[l]), l


def test_y_pyramid_fused_levels_many_tiles_per_cta(gpu, oracle):
    """A 24-frame batch of random 1080p frames (2 160 tiles of pyr_down2_kernel, 3 672 of the level-0
    kernel): every level of every pyramid matters for the vectors of random content."""
    w, h, n = 1920, 1080, 24
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=n)) as s:
        mv, mad, _ = s.encode(frames, want_stream=False)
        pw, ph = s.padded_w, s.padded_h
    pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
    for i in range(1, n):  # random frames: every level of both pyramids matters for the vectors
        emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 8)
        assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad), i


def test_y_pyramid_fused_levels_in_a_batch(gpu, oracle):
    """The session path: pyramids of a whole batch (slot stride, first_slot = 1) feed the search."""
    w, h, n = 432, 248, 7
    frames = SyntheticSequence(w, h, n, seed=77).frames()
    for L, R in ((4, 8), (5, 16)):
        with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L, max_batch=4)) as s:
            mv, mad, _ = s.encode(frames, want_stream=False)
            pw, ph = s.padded_w, s.padded_h
        pyr = [oracle.y_pyramid(f, pw, ph, L) for f in frames]
        for i in range(1, n):
            emv, emad = oracle.hbma(pyr[i - 1], pyr[i], R)
            assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad), (L, i)


# ---------------------------------------------------------------- K2: motion
def test_hbma_golden_default(gpu):
    g = load_golden("small_default.npz")
    for i in (1, 2):
        t = [g[f"pyr{i-1}_{l}"] for l in range(4)]
        a = [g[f"pyr{i}_{l}"] for l in range(4)]
        mv, mad = gpu.EstimateMotionHierarchical16x16Sse2(t, a, int(g["pw"]), int(g["ph"]), 8)
        assert np.array_equal(mv, g[f"mv{i}"]) and np.array_equal(mad, g[f"mad{i}"])
        mv, mad = gpu.EstimateMotionHierarchical(t, a, 4, int(g["pw"]), int(g["ph"]), 8, 16, 16)
        assert np.array_equal(mv, g[f"mv{i}"]) and np.array_equal(mad, g[f"mad{i}"])


@pytest.mark.parametrize("R", [8, 16, 32])
def test_hbma_golden_ranges(gpu, R):
    g = load_golden("aligned_default.npz")
    t = [g[f"pyr0_{l}"] for l in range(4)]
    a = [g[f"pyr1_{l}"] for l in range(4)]
    mv, mad = gpu.EstimateMotionHierarchical16x16Sse2(t, a, int(g["pw"]), int(g["ph"]), R)
    assert np.array_equal(mv, g[f"mv_R{R}"]) and np.array_equal(mad, g[f"mad_R{R}"])


def test_hbma_golden_generic(gpu):
    g = load_golden("generic_motion.npz")
    for ci, (lv, bw, bh, rr) in enumerate(g["cases"]):
        t = [g[f"t{ci}_{l}"] for l in range(lv)]
        a = [g[f"a{ci}_{l}"] for l in range(lv)]
        fh, fw = t[0].shape
        mv, mad = gpu.EstimateMotionHierarchical(t, a, int(lv), fw, fh, int(rr), int(bw), int(bh))
        assert np.array_equal(mv, g[f"mv{ci}"]), ci
        assert np.array_equal(mad, g[f"mad{ci}"]), ci
    fh, fw = g["t3"].shape
    mv, mad = gpu.EstimateMotionExhaustiveSearch(g["t3"], g["a3"], fw, fh, 3, 8, 8)
    assert np.array_equal(mv, g["ebma_mv"]) and np.array_equal(mad, g["ebma_mad"])


@pytest.mark.parametrize("w,h,R,seed", [(960, 540, 8, 1), (960, 540, 32, 2), (320, 180, 16, 3),
                                        (64, 48, 8, 4), (16, 16, 8, 5), (1920, 1080, 8, 6)])
def test_hbma_vs_oracle_synthetic(gpu, oracle, w, h, R, seed):
    pw, ph = gpu.padded_dim(w, 16, 4), gpu.padded_dim(h, 16, 4)
    seq = SyntheticSequence(w, h, 2, seed=seed)
    p0, p1 = oracle.y_pyramid(seq.frame(0), pw, ph, 4), oracle.y_pyramid(seq.frame(1), pw, ph, 4)
    mv, mad = gpu.EstimateMotionHierarchical16x16Sse2(p0, p1, pw, ph, R)
    emv, emad = oracle.hbma(p0, p1, R)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)
    if oracle.have_ref():
        rmv, rmad = oracle.hbma(p0, p1, R, impl="ref_sse2")
        assert np.array_equal(mv, rmv) and np.array_equal(mad, rmad)


@pytest.mark.parametrize("L,bw,bh,R", [(1, 16, 16, 5), (2, 8, 8, 9), (3, 32, 16, 12),
                                       (5, 16, 16, 16), (2, 6, 10, 7), (1, 3, 7, 1), (4, 64, 64, 8)])
def test_hbma_vs_oracle_generic(gpu, oracle, L, bw, bh, R):
    w, h = bw * 7, bh * 5
    pw, ph = gpu.padded_dim(w, bw, L), gpu.padded_dim(h, bh, L)
    seq = SyntheticSequence(w, h, 2, seed=bw + bh + L, n_rects=2)
    p0, p1 = oracle.y_pyramid(seq.frame(0), pw, ph, L), oracle.y_pyramid(seq.frame(1), pw, ph, L)
    mv, mad = gpu.EstimateMotionHierarchical(p0, p1, L, pw, ph, R, bw, bh)
    emv, emad = oracle.hbma(p0, p1, R, bw, bh)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


@pytest.mark.parametrize("L,R", [(4, 8), (4, 16), (4, 24), (4, 32), (3, 4), (3, 8), (3, 12), (3, 16),
                                 (5, 16), (5, 32), (2, 2), (2, 4), (2, 6), (2, 8),
                                 # 5 levels, r = 3, 4: tile kernel over levels 4..2 + refinement launches
                                 (5, 48), (5, 55), (5, 64), (5, 79)])
@pytest.mark.parametrize("w,h", [(352, 208), (176, 80)])
def test_hbma_tiled_path_vs_oracle(gpu, oracle, L, R, w, h):
    """16x16 blocks with top-level range r = R >> (L-1) <= 4: the TMA-staged tiled kernel
    (partial tiles on both axes, frame-border clamping, flat patch ties)."""
    pw, ph = gpu.padded_dim(w, 16, L), gpu.padded_dim(h, 16, L)
    seq = SyntheticSequence(w, h, 2, seed=L * 10 + R)
    p0, p1 = oracle.y_pyramid(seq.frame(0), pw, ph, L), oracle.y_pyramid(seq.frame(1), pw, ph, L)
    mv, mad = gpu.EstimateMotionHierarchical(p0, p1, L, pw, ph, R, 16, 16)
    emv, emad = oracle.hbma(p0, p1, R)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


@pytest.mark.parametrize("L,R", [(4, 8), (4, 16), (4, 23), (3, 4), (3, 8), (3, 16), (2, 2), (2, 5), (2, 8), (1, 1),
                                 (1, 2), (1, 4)])
@pytest.mark.parametrize("w,h", [(328, 200), (88, 40), (8, 64)])
def test_hbma_tiled_path_8x8_blocks(gpu, oracle, L, R, w, h):
    """8x8 motion blocks on the TMA-staged tiled kernel with base block 8: the top
    level of L = 4 compares single pixels; partial tiles, one-block-wide frames, clamping."""
    pw, ph = gpu.padded_dim(w, 8, L), gpu.padded_dim(h, 8, L)
    seq = SyntheticSequence(w, h, 2, seed=L * 11 + R, n_rects=2)
    p0, p1 = oracle.y_pyramid(seq.frame(0), pw, ph, L), oracle.y_pyramid(seq.frame(1), pw, ph, L)
    mv, mad = gpu.EstimateMotionHierarchical(p0, p1, L, pw, ph, R, 8, 8)
    emv, emad = oracle.hbma(p0, p1, R, 8, 8)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)
    z = [np.zeros_like(a) for a in p0]  # flat frames: ties everywhere, zero-vector rule
    mv, mad = gpu.EstimateMotionHierarchical(z, z, L, pw, ph, R, 8, 8)
    assert not mv.any() and not mad.any()


@pytest.mark.parametrize("L,R", [(4, 64), (4, 100), (3, 32), (3, 60), (2, 32), (2, 10), (1, 3), (1, 8),
                                 (1, 16), (1, 40), (5, 80), (5, 128)])
@pytest.mark.parametrize("w,h", [(176, 112), (64, 48)])
def test_hbma_window_path_vs_oracle(gpu, oracle, L, R, w, h):
    """16x16 blocks with a large top-level range: the per-block TMA window kernel (windows
    larger than the frame, clamping on every side, flat-patch ties, zero padding rows)."""
    pw, ph = gpu.padded_dim(w, 16, L), gpu.padded_dim(h, 16, L)
    seq = SyntheticSequence(w, h, 2, seed=L * 7 + R)
    p0, p1 = oracle.y_pyramid(seq.frame(0), pw, ph, L), oracle.y_pyramid(seq.frame(1), pw, ph, L)
    mv, mad = gpu.EstimateMotionHierarchical(p0, p1, L, pw, ph, R, 16, 16)
    emv, emad = oracle.hbma(p0, p1, R)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


@pytest.mark.parametrize("L,R", [(1, 5), (1, 8), (1, 12), (1, 16), (1, 27), (1, 32), (1, 50), (1, 64),
                                 (2, 16), (2, 34), (2, 64), (2, 128), (3, 32), (3, 64), (3, 128),
                                 (4, 64), (4, 128), (5, 128)])
@pytest.mark.parametrize("w,h", [(416, 240), (48, 176)])
def test_hbma_pooled_window_path_vs_oracle(gpu, oracle, L, R, w, h):
    """16x16 blocks, top-level range 5..64: the pooled kernel with pre-shifted window copies
    (every range class, ranges inside a class, interior blocks with unclamped windows as well as
    windows clamped on every side, a frame narrower than the window, flat-patch ties).  The session
    test hook hbma_kernel_family selects it also where the dispatcher prefers another kernel."""
    seq = SyntheticSequence(w, h, 2, seed=L * 13 + R)
    frames = np.stack([seq.frame(0), seq.frame(1)])
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L,
                                       hbma_kernel_family=gpu.HBMA_FAMILY_POOL)) as s:
        mv, mad, _ = s.encode(frames, want_stream=False)
        pw, ph = s.padded_w, s.padded_h
    p0, p1 = oracle.y_pyramid(frames[0], pw, ph, L), oracle.y_pyramid(frames[1], pw, ph, L)
    emv, emad = oracle.hbma(p0, p1, R)
    assert np.array_equal(mv[0], emv) and np.array_equal(mad[0], emad)


@pytest.mark.parametrize("family", ["GENERIC", "WINDOW", "TILE"])
@pytest.mark.parametrize("L,R", [(1, 8), (2, 16), (4, 8), (4, 64), (5, 64)])
def test_hbma_kernel_family_hook_vs_oracle(gpu, oracle, family, L, R):
    """svc_session_config.hbma_kernel_family (the library's one test hook): the universal kernel and
    the per-block window kernels on configurations the dispatcher gives to faster kernels."""
    w, h = 272, 144
    seq = SyntheticSequence(w, h, 3, seed=L * 5 + R)
    frames = seq.frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L,
                                       hbma_kernel_family=getattr(gpu, "HBMA_FAMILY_" + family))) as s:
        mv, mad, _ = s.encode(frames, want_stream=False)
        pw, ph = s.padded_w, s.padded_h
    pyr = [oracle.y_pyramid(f, pw, ph, L) for f in frames]
    for i in (1, 2):
        emv, emad = oracle.hbma(pyr[i - 1], pyr[i], R)
        assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad)


@pytest.mark.parametrize("w,h,n", [(1920, 1080, 3), (960, 540, 3), (130, 70, 4), (16, 16, 3), (24, 200, 3), (400, 24, 3),
                                   (272, 144, 5), (128, 128, 3), (144, 272, 3)])
def test_hbma_strip_kernel_default_config(gpu, oracle, w, h, n):
    """Encoder default (16x16, R=8, L=4): hbma_strip_kernel (k_hbma_strip.cu: a lane owns a strip of a
    block and all nine candidates) against the oracle and against the bounded-reach tile kernel it
    replaced (test hook family TILE) -- whole tiles, partial tiles on both edges, motion fields of one
    block row / column, a single block."""
    frames = SyntheticSequence(w, h, n, seed=w + h).frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h)) as s:
        mv, mad, _ = s.encode(frames, want_stream=False)
        pw, ph = s.padded_w, s.padded_h
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, hbma_kernel_family=gpu.HBMA_FAMILY_TILE)) as s:
        mv_t, mad_t, _ = s.encode(frames, want_stream=False)
    assert np.array_equal(mv, mv_t) and np.array_equal(mad, mad_t)
    pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
    for i in range(1, n):
        emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 8)
        assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad)


def test_hbma_strip_kernel_special_content(gpu, oracle):
    """Flat, saturated and ramp frames through the strip kernel: ties everywhere (first minimum at the
    refinement levels, last minimum and the zero-vector rule at the top level)."""
    w, h = 208, 144
    ramp = (np.arange(w, dtype=np.int64)[None, :, None] * 3 + np.arange(h)[:, None, None] * 5 + np.zeros((1, 1, 3), np.int64))
    seqs = [np.stack([np.full((h, w, 3), v, np.uint8) for v in (0, 255, 255, 17)]),
            np.stack([(ramp % 256).astype(np.uint8), ((ramp + 7) % 256).astype(np.uint8), ((ramp // 2) % 256).astype(np.uint8)])]
    for frames in seqs:
        with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h)) as s:
            mv, mad, _ = s.encode(frames, want_stream=False)
            pw, ph = s.padded_w, s.padded_h
        pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
        for i in range(1, len(frames)):
            emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 8)
            assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad)


@pytest.mark.parametrize("L,R", [(2, 10), (2, 16), (2, 27), (2, 32), (2, 50), (2, 64), (3, 20), (3, 32), (3, 45),
                                 (3, 64), (3, 100), (3, 128), (4, 40), (4, 64), (4, 100), (4, 128), (4, 200), (4, 256),
                                 (5, 48), (5, 64), (5, 80), (5, 128), (5, 300), (5, 512)])
@pytest.mark.parametrize("w,h", [(416, 240), (48, 176)])
def test_hbma_level_synchronous_path_vs_oracle(gpu, oracle, L, R, w, h):
    """2..5 levels, top-level range 5..32 (3..32 for 5 levels): one launch per level (shared-window tile kernel at the top
    level, single-level refinement kernels below), the coarser level's vector and MAD carried through
    the output arrays -- partial tiles, frames narrower than a window, flat-patch ties, the
    zero-vector rule of the top level and the strict '<' of the refinement levels."""
    pw, ph = gpu.padded_dim(w, 16, L), gpu.padded_dim(h, 16, L)
    seq = SyntheticSequence(w, h, 2, seed=L * 19 + R)
    p0, p1 = oracle.y_pyramid(seq.frame(0), pw, ph, L), oracle.y_pyramid(seq.frame(1), pw, ph, L)
    mv, mad = gpu.EstimateMotionHierarchical(p0, p1, L, pw, ph, R, 16, 16)
    emv, emad = oracle.hbma(p0, p1, R)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)
    z = [np.zeros_like(a) for a in p0]
    f = [np.full_like(a, 255) for a in p0]
    for t, a in ((z, z), (z, f)):
        mv, mad = gpu.EstimateMotionHierarchical(t, a, L, pw, ph, R, 16, 16)
        emv, emad = oracle.hbma(t, a, R)
        assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


@pytest.mark.parametrize("R,L", [(8, 1), (16, 1), (32, 1), (40, 1), (64, 1), (100, 1), (64, 2)])
def test_hbma_pooled_window_monotone_sequences(gpu, oracle, R, L):
    """Top-level zero-vector rule (libs/motion.cpp:333-337) on the pooled kernel: horizontal /
    vertical ramps and constant frames make the SAD sequence non-increasing over whole windows
    for some blocks and break it by a single candidate for others."""
    w, h = 224, 160
    yy, xx = np.mgrid[0:h, 0:w]
    bases = [np.clip(xx, 0, 255), np.clip(yy, 0, 255), np.clip(255 - xx - yy // 2, 0, 255),
             np.full((h, w), 7), np.clip((xx // 16) * 9 + (yy // 16) * 5, 0, 255)]
    def pyr(img):
        out = [img.astype(np.uint8)]
        for _ in range(L - 1):
            out.append(oracle.pyr_down(out[-1]))
        return out
    for i, a in enumerate(bases):
        for t in (bases[(i + 1) % len(bases)], a):
            tp, ap = pyr(t), pyr(a)
            mv, mad = gpu.EstimateMotionHierarchical(tp, ap, L, w, h, R, 16, 16)
            emv, emad = oracle.hbma(tp, ap, R)
            assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


@pytest.mark.parametrize("r", [5, 8, 13, 16, 21, 32, 33, 47, 64, 80, 112])
@pytest.mark.parametrize("w,h", [(432, 96), (208, 64), (16, 160)])
def test_ebma_16x16_shared_window_tiles(gpu, oracle, r, w, h):
    """EstimateMotionExhaustiveSearch with 16x16 blocks (= HBMA with one level): the tile kernel that
    shares one search window between horizontally adjacent blocks (r <= 32) and the kernel that
    searches a wide window in column stripes (r > 32) -- partial tiles at the right frame edge, a
    single-column frame, 1..5 stripes, windows clamped on every side, flat-patch ties."""
    seq = SyntheticSequence(w, h, 2, seed=r * 5 + w)
    t = oracle.y_pyramid(seq.frame(0), w, h, 1)[0]
    a = oracle.y_pyramid(seq.frame(1), w, h, 1)[0]
    mv, mad = gpu.EstimateMotionExhaustiveSearch(t, a, w, h, r, 16, 16)
    emv, emad = oracle.ebma(t, a, r, 16, 16)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


def test_hbma_flat_frames_tie_break(gpu, oracle):
    z = [np.zeros((64 >> l, 96 >> l), np.uint8) for l in range(4)]
    c = [np.full((64 >> l, 96 >> l), 9, np.uint8) for l in range(4)]
    for t, a in ((z, z), (z, c), (c, z)):
        mv, mad = gpu.EstimateMotionHierarchical16x16Sse2(t, a, 96, 64, 8)
        emv, emad = oracle.hbma(t, a, 8)
        assert np.array_equal(mv, emv) and np.array_equal(mad, emad)
        assert not mv.any()


@pytest.mark.parametrize("R,L", [(8, 4), (64, 4), (20, 1), (24, 2)])
def test_hbma_saturated_sad_ties(gpu, oracle, R, L):
    """Tracked all 0, anchor all 255: every candidate has the maximum SAD (MAD 255) -- the widest
    packed keys, ties everywhere -- on the tiled, warp-window and block-window kernels."""
    w, h = 176, 112
    z = [np.zeros((h >> l, w >> l), np.uint8) for l in range(L)]
    f = [np.full((h >> l, w >> l), 255, np.uint8) for l in range(L)]
    for t, a in ((z, f), (f, z)):
        mv, mad = gpu.EstimateMotionHierarchical(t, a, L, w, h, R, 16, 16)
        emv, emad = oracle.hbma(t, a, R)
        assert np.array_equal(mv, emv) and np.array_equal(mad, emad)
        assert not mv.any() and (mad == 255.0).all()


def test_ebma_zero_range(gpu, oracle):
    rng = np.random.default_rng(3)
    t = rng.integers(0, 256, size=(32, 48)).astype(np.uint8)
    a = rng.integers(0, 256, size=(32, 48)).astype(np.uint8)
    mv, mad = gpu.EstimateMotionExhaustiveSearch(t, a, 48, 32, 0, 8, 8)
    emv, emad = oracle.ebma(t, a, 0, 8, 8)
    assert np.array_equal(mv, emv) and np.array_equal(mad, emad)


# ---------------------------------------------------------------- K3: DCT + stream
@pytest.mark.parametrize("name", ["small_default.npz", "aligned_default.npz"])
def test_dct_planar_golden(gpu, name):
    g = load_golden(name)
    d = gpu.dct_planar(g["frames"][1], int(g["pw"]), int(g["ph"]))
    assert np.abs(d - g["dct1"]).max() <= DCT_TOL


def test_dct_known_answers(gpu):
    g = load_golden("dct_kat.npz")
    for key, tbw, tbh in (("8", 8, 8), ("4", 4, 4), ("16", 16, 16), ("48", 8, 4)):
        for b, o in zip(g["b" + key], g["o" + key]):
            bgr = np.repeat(b.astype(np.uint8)[..., None], 3, axis=2)
            d = gpu.dct_planar(bgr, tbw, tbh, tbw, tbh)
            for c in range(3):
                assert np.abs(d[c] - o).max() <= DCT_TOL, (key, c)


def test_dct_known_answers_through_the_fused_stream_kernels(gpu):
    """The cv2.dct known-answer blocks (tests/golden/dct_kat.npz) laid side by side in one frame and
    sent through the fused 8x8 / 4x4 / 16x16 stream kernels: record k must hold cv2's coefficients of
    block k in all three channels."""
    g = load_golden("dct_kat.npz")
    for key, tb in (("8", 8), ("4", 4), ("16", 16)):
        blocks, outs = g["b" + key], g["o" + key]
        row = np.concatenate([b.astype(np.uint8) for b in blocks], axis=1)
        while row.shape[1] % 16:  # frame width: a multiple of 16 (the fused 16x16 kernel loads 16-byte aligned rows)
            row = np.concatenate([row, row], axis=1)
        frame = np.tile(row, (16 // tb, 1))  # 16 rows: one row of motion blocks
        bgr = np.ascontiguousarray(np.repeat(frame[..., None], 3, axis=2))
        nbx = bgr.shape[1] // tb
        st = gpu.encode_frame_stream(bgr, bgr.shape[1], 16, tb, tb)
        rec = st.view(np.uint32).reshape(-1, 1 + 3 * tb * tb)
        assert rec.shape[0] == nbx * (16 // tb)
        for k in range(rec.shape[0]):
            exp = outs[(k % nbx) % len(blocks)]
            for c in range(3):
                got = rec[k, 1 + c * tb * tb: 1 + (c + 1) * tb * tb].view(np.float32).reshape(tb, tb)
                assert np.abs(got - exp).max() <= DCT_TOL, (key, k, c)


@pytest.mark.parametrize("w,h,tbw,tbh", [(1920, 1080, 8, 8), (104, 56, 8, 8), (96, 48, 4, 4),
                                         (64, 64, 16, 16), (80, 48, 8, 4), (48, 80, 2, 16)])
def test_dct_planar_vs_oracle(gpu, oracle, w, h, tbw, tbh):
    rng = np.random.default_rng(w + h + tbw)
    f = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    pw, ph = gpu.padded_dim(w, 16, 4), gpu.padded_dim(h, 16, 4)
    d = gpu.dct_planar(f, pw, ph, tbw, tbh)
    e = oracle.dct_planar(f, pw, ph, tbw, tbh)
    assert np.abs(d - e).max() <= DCT_TOL


def _stream_check(gpu, oracle, f, pw, ph, tbw, tbh, bt):
    h, w, _ = f.shape
    st = gpu.encode_frame_stream(f, pw, ph, tbw, tbh, 16, 16, bt)
    planes = oracle.dct_planar(f, pw, ph, tbw, tbh)
    exp = oracle.serialize_frame(planes, bt, w, h, tbw, tbh, pw // 16, 16, 16)
    assert st.size == exp.size
    rec = 1 + 3 * tbw * tbh
    got_w = st.view(np.uint32).reshape(-1, rec)
    exp_w = exp.view(np.uint32).reshape(-1, rec)
    assert np.array_equal(got_w[:, 0], exp_w[:, 0])  # block types: exact
    assert np.abs(got_w[:, 1:].view(np.float32) - exp_w[:, 1:].view(np.float32)).max() <= DCT_TOL


@pytest.mark.parametrize("w,h,tbw,tbh", [(160, 92, 8, 8), (104, 56, 8, 8), (1920, 1080, 8, 8),
                                         (960, 540, 8, 8), (48, 40, 8, 8), (96, 48, 4, 4),
                                         (80, 48, 8, 4), (100, 36, 16, 16),
                                         # fused 16x16 / 4x4 stream kernels (w == padded w)
                                         (160, 92, 16, 16), (64, 64, 16, 16), (1920, 1080, 16, 16),
                                         (16, 16, 16, 16), (48, 40, 4, 4), (176, 130, 4, 4), (1920, 1080, 4, 4),
                                         (2000, 48, 4, 4)])
def test_stream_layout_vs_oracle(gpu, oracle, w, h, tbw, tbh):
    rng = np.random.default_rng(w * 3 + h + tbw)
    f = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    pw, ph = gpu.padded_dim(w, 16, 4), gpu.padded_dim(h, 16, 4)
    bt = rng.integers(0, 11, size=(ph // 16) * (pw // 16)).astype(np.uint32)
    _stream_check(gpu, oracle, f, pw, ph, tbw, tbh, bt)
    _stream_check(gpu, oracle, f, pw, ph, tbw, tbh, None)


# ---------------------------------------------------------------- session
@pytest.mark.parametrize("w,h,n,batch", [(960, 540, 6, 4), (104, 56, 5, 2), (160, 92, 4, 32)])
def test_session_host_path_matches_oracle(gpu, oracle, w, h, n, batch):
    seq = SyntheticSequence(w, h, n, seed=21)
    frames = seq.frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=batch)) as s:
        pw, ph = s.padded_w, s.padded_h
        rng = np.random.default_rng(9)
        bt = rng.integers(0, 6, size=(n - 1, s.mv_field_h * s.mv_field_w)).astype(np.uint32)
        # feed in two uneven pushes: exercises the resident previous pyramid
        mv1, mad1, st1 = s.encode(frames[:2], block_types=bt[:1])
        mv2, mad2, st2 = s.encode(frames[2:], block_types=bt[1:])
        mv = np.concatenate([mv1, mv2]); mad = np.concatenate([mad1, mad2]); st = np.concatenate([st1, st2])
        assert mv.shape[0] == n - 1
        pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
        for i in range(1, n):
            emv, emad = oracle.hbma(pyr[i - 1], pyr[i], 8)
            assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad), i
            planes = oracle.dct_planar(frames[i], pw, ph)
            exp = oracle.serialize_frame(planes, bt[i - 1], w, h, 8, 8, pw // 16, 16, 16)
            g = st[i - 1].view(np.uint32).reshape(-1, 193)
            e = exp.view(np.uint32).reshape(-1, 193)
            assert np.array_equal(g[:, 0], e[:, 0])
            assert np.abs(g[:, 1:].view(np.float32) - e[:, 1:].view(np.float32)).max() <= DCT_TOL
        assert s.launch_count > 0


def test_session_device_path_full_1080p_properties(gpu, oracle):
    """BASELINE config 2 geometry on the device-resident path: a few frames are
    checked exactly against the oracle, the rest through properties that do
    not need the oracle (static repeat -> zero motion and zero MAD; DC
    coefficient = 8 * block mean)."""
    w, h, n = 1920, 1080, 12
    seq = SyntheticSequence(w, h, n, seed=5)
    frames = seq.frames()
    frames[7] = frames[6]  # a repeated frame: exact zero-motion / zero-MAD property
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=5)) as s:
        mvn = s.mv_field_w * s.mv_field_h
        d_in = gpu.DeviceBuffer(0, frames.nbytes)
        d_mv = gpu.DeviceBuffer(0, (n - 1) * mvn * 8)
        d_mad = gpu.DeviceBuffer(0, (n - 1) * mvn * 4)
        d_st = gpu.DeviceBuffer(0, (n - 1) * s.frame_stream_bytes)
        d_in.upload(frames)
        ne = s.encode_device(d_in, n, d_mv, d_mad, d_st)
        s.synchronize()
        assert ne == n - 1
        mv = d_mv.download(np.float32, (n - 1, s.mv_field_h, s.mv_field_w, 2))
        mad = d_mad.download(np.float32, (n - 1, s.mv_field_h, s.mv_field_w))
        st = d_st.download(np.uint8, (n - 1, s.frame_stream_bytes))
        pw, ph = s.padded_w, s.padded_h
        for i in (1, 6, 11):
            p0, p1 = oracle.y_pyramid(frames[i - 1], pw, ph, 4), oracle.y_pyramid(frames[i], pw, ph, 4)
            emv, emad = oracle.hbma(p0, p1, 8)
            assert np.array_equal(mv[i - 1], emv) and np.array_equal(mad[i - 1], emad)
        # frame 7 == frame 6: every block matches itself exactly
        assert not mad[6].any()
        # DC term of every 8x8 block and channel = sum / 8
        rec = st.view(np.uint32).reshape(n - 1, 135, 240, 193)
        assert not rec[..., 0].any()
        for i in (2, 9):
            dc = rec[i - 1][..., 1::64].view(np.float32)  # (135, 240, 3)
            exp = frames[i].reshape(135, 8, 240, 8, 3).astype(np.float64).sum(axis=(1, 3)) / 8.0
            assert np.abs(dc - exp).max() <= DCT_TOL
        planes = oracle.dct_planar(frames[4], pw, ph)
        exp = oracle.serialize_frame(planes, None, w, h, 8, 8, pw // 16, 16, 16)
        assert np.abs(st[3].view(np.float32) - exp.view(np.float32)).max() <= DCT_TOL
        for b in (d_in, d_mv, d_mad, d_st):
            b.free()


def test_session_4k_geometry(gpu, oracle):
    """BASELINE config 4 geometry (3840x2160, no padding): exact motion on one pair, DC property
    and a sampled coefficient check on the stream."""
    w, h, n = 3840, 2160, 3
    seq = SyntheticSequence(w, h, n, seed=44)
    frames = seq.frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=2)) as s:
        assert (s.padded_w, s.padded_h, s.mv_field_w, s.mv_field_h) == (3840, 2160, 240, 135)
        assert s.frame_stream_bytes == 480 * 270 * 772
        mv, mad, st = s.encode(frames)
        p0, p1 = oracle.y_pyramid(frames[1], w, h, 4), oracle.y_pyramid(frames[2], w, h, 4)
        emv, emad = oracle.hbma(p0, p1, 8)
        assert np.array_equal(mv[1], emv) and np.array_equal(mad[1], emad)
        rec = st[0].view(np.uint32).reshape(270, 480, 193)
        dc = rec[..., 1::64].view(np.float32)
        exp = frames[1].reshape(270, 8, 480, 8, 3).astype(np.float64).sum(axis=(1, 3)) / 8.0
        assert np.abs(dc - exp).max() <= DCT_TOL
        crop = np.ascontiguousarray(frames[1][1040:1104, 1920:2048])  # 8 x 16 blocks
        planes = oracle.dct_planar(crop, 128, 64)
        got = rec[130:138, 240:256, 1:].view(np.float32).reshape(8, 16, 3, 8, 8)
        exp_blocks = planes.reshape(3, 8, 8, 16, 8).transpose(1, 3, 0, 2, 4)
        assert np.abs(got - exp_blocks).max() <= DCT_TOL


@pytest.mark.parametrize("R,L", [(8, 4), (24, 3), (40, 1), (64, 5)])
def test_device_work_counters_match_oracle(gpu, oracle, R, L):
    w, h, n = 176, 112, 3
    frames = SyntheticSequence(w, h, n, seed=R).frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, mv_search_range=R, pyr_lvl_count=L,
                                       max_batch=4)) as s:
        d_in = gpu.DeviceBuffer(0, frames.nbytes)
        d_in.upload(frames)
        s.run_stage(gpu.STAGE_Y_PYRAMID, d_in, n)
        cand, absd = s.hbma_work(n)
        pw, ph = s.padded_w, s.padded_h
        pyr = [[np.zeros((ph >> l, pw >> l), np.uint8) for l in range(L)]]
        pyr += [oracle.y_pyramid(f, pw, ph, L) for f in frames]
        ec = ea = 0
        for i in range(n):
            c, a = oracle.hbma_count(pyr[i], pyr[i + 1], R)
            ec += c
            ea += a
        assert (cand, absd) == (ec, ea)
        d_in.free()
    assert gpu.sad_peak(0) > 1e12  # > 1 T byte-absdiff/s on any B200


def test_session_output_subsets_reset_and_concurrent_sessions(gpu, oracle):
    """NULL outputs skip their stage, reset() restarts the sequence, and two sessions driven
    from two host threads on one GPU do not disturb each other."""
    import threading
    w, h, n = 320, 180, 7
    frames = SyntheticSequence(w, h, n, seed=77).frames()
    pw, ph = gpu.padded_dim(w, 16, 4), gpu.padded_dim(h, 16, 4)
    pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
    emv = [oracle.hbma(pyr[i - 1], pyr[i], 8)[0] for i in range(1, n)]
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=3)) as s:
        mv, mad, st = s.encode(frames, want_mad=False, want_stream=False)
        assert mad is None and st is None and all(np.array_equal(mv[i], emv[i]) for i in range(n - 1))
        s.reset()
        mv, mad, st = s.encode(frames, want_mv=False, want_mad=False)
        assert mv is None and st.shape == (n - 1, s.frame_stream_bytes)
        exp = oracle.serialize_frame(oracle.dct_planar(frames[3], pw, ph), None, w, h, 8, 8, pw // 16, 16, 16)
        assert np.abs(st[2].view(np.float32) - exp.view(np.float32)).max() <= DCT_TOL
        s.reset()
        mv2, _, _ = s.encode(frames[2:])  # a fresh sequence starting at frame 2
        assert mv2.shape[0] == n - 3 and np.array_equal(mv2[0], emv[2])
    results = {}

    def run(tag, seed):
        fr = SyntheticSequence(w, h, n, seed=seed).frames()
        with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=2)) as ss:
            for _ in range(3):
                ss.reset()
                results[tag] = (fr, ss.encode(fr)[0])

    th = [threading.Thread(target=run, args=(k, 100 + k)) for k in range(2)]
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    for k in range(2):
        fr, mv = results[k]
        p = [oracle.y_pyramid(f, pw, ph, 4) for f in fr]
        for i in range(1, n):
            assert np.array_equal(mv[i - 1], oracle.hbma(p[i - 1], p[i], 8)[0])


def test_stage_entry_points(gpu, oracle):
    w, h, n = 320, 180, 3
    seq = SyntheticSequence(w, h, n, seed=8)
    frames = seq.frames()
    with gpu.Session(gpu.SessionConfig(frame_w=w, frame_h=h, max_batch=4)) as s:
        mvn = s.mv_field_w * s.mv_field_h
        d_in = gpu.DeviceBuffer(0, frames.nbytes)
        d_mv = gpu.DeviceBuffer(0, n * mvn * 8)
        d_mad = gpu.DeviceBuffer(0, n * mvn * 4)
        d_st = gpu.DeviceBuffer(0, n * s.frame_stream_bytes)
        d_in.upload(frames)
        s.run_stage(gpu.STAGE_Y_PYRAMID, d_in, n)
        s.run_stage(gpu.STAGE_HBMA, None, n, d_mv, d_mad)
        s.run_stage(gpu.STAGE_DCT_STREAM, d_in, n, None, None, d_st)
        s.synchronize()
        mv = d_mv.download(np.float32, (n, s.mv_field_h, s.mv_field_w, 2))
        pw, ph = s.padded_w, s.padded_h
        pyr = [oracle.y_pyramid(f, pw, ph, 4) for f in frames]
        for i in (1, 2):  # slot i+1 vs slot i  (slot 0 is the zero-initialised previous frame)
            emv, _ = oracle.hbma(pyr[i - 1], pyr[i], 8)
            assert np.array_equal(mv[i], emv)
